"""HBM-bound kernels at C3 size: separate_scores (five selections, one streaming pass) and the
fused means (no selections materialised).  Prints achieved GB/s against the algorithmic bytes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from munk_b200.device import Device
from munk_b200.scores_plan import plan_separate_scores

dev = Device(0)
n1, n2, k, H = 20000, 16000, 400, 4000
rng = np.random.default_rng(0)
hs, ht = rng.permutation(n1)[:H], rng.permutation(n2)[:H]
hom = list(zip(hs.tolist(), ht.tolist()))
S = torch.rand(n1, n2, dtype=torch.float64, device="cuda")
t0 = time.time(); plan = plan_separate_scores(n1, n2, hom[:k], hom); t_plan = time.time() - t0


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best


ms = timed(lambda: plan.run(dev, S, to_host=False))
out_elems = plan.n_lloff + plan.n_lnonl + plan.n_other + plan.n_hh + k
gb = (n1 * n2 + out_elems) * 8 / 1e9
print("separate_scores  %d x %d: host plan %.1f ms, device %.3f ms, %.2f GB moved -> %.0f GB/s" % (n1, n2, t_plan * 1e3, ms, gb, gb / ms * 1e3))
ms = timed(lambda: plan.means(dev, S))
gb = n1 * n2 * 8 / 1e9
print("separate_scores_means: %.3f ms, %.2f GB read -> %.0f GB/s" % (ms, gb, gb / ms * 1e3))
