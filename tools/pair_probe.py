"""Time munk_potrf_pair at C3 sizes for the panel-width / lead knobs of the environment
(MUNK_PANEL_BIG_UNTIL, MUNK_PAIR_LEAD_PCT, ...; read once per process, so one process per setting).
Random sparse graphs (no networkx), same density as the bench's.  usage: python tools/pair_probe.py"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from munk_b200.device import Device, side_device  # noqa: E402

n1, n2 = 20000, 16000
dev = Device(0)
dev2, hi, lo = side_device(dev)
rng = np.random.default_rng(0)


def edges(n):
    u = rng.integers(0, n, 10 * n).astype(np.int32)
    v = rng.integers(0, n, 10 * n).astype(np.int32)
    keep = u != v
    lo_, hi_ = np.minimum(u, v)[keep], np.maximum(u, v)[keep]
    key = np.unique(lo_.astype(np.int64) * n + hi_)
    return (key // n).astype(np.int32), (key % n).astype(np.int32)


e1, e2 = edges(n1), edges(n2)
l1, l2 = dev.new_linv(n1), dev2.new_linv(n2)
best = 1e9
A1, A2 = dev.empty(n1, n1), dev2.empty(n2, n2)
with torch.cuda.stream(hi):
    for rep in range(5):
        dev.assemble(n1, e1[0], e1[1], 0.05, out=A1)
        dev2.assemble(n2, e2[0], e2[1], 0.05, out=A2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dev.potrf_pair(A1, n1, l1, dev2, A2, n2, l2)
        b.record()
        torch.cuda.synchronize()
        if rep >= 2:
            best = min(best, a.elapsed_time(b))
    assert dev.potrf_info() == 0 and dev2.potrf_info() == 0
knobs = {k: v for k, v in os.environ.items() if k.startswith("MUNK_")}
print("potrf_pair %d + %d: %.2f ms  %s" % (n1, n2, best, knobs), flush=True)
