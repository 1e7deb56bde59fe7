"""Times the C3 score product S = W1^T P (and a plain 8192^3 product) for the current tile order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from munk_b200.device import Device, KC, MC
dev = Device(0)
n1, n2 = 20000, 16000
W = torch.tril(torch.randn(n1, n1, dtype=torch.float64, device="cuda"))
P = torch.randn(n1, n2, dtype=torch.float64, device="cuda")
S = dev.empty(n1, n2)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
ms = timed(lambda: dev.score(W, n1, P, n2, out=S))
print("strip=%s score %.2f ms %.2f TF/s" % (os.environ.get("MUNK_GEMM_STRIP", "12"), ms, n1 * n1 * n2 / ms / 1e9))
a = torch.randn(8192, 8192, dtype=torch.float64, device="cuda"); c = torch.empty_like(a)
ms = timed(lambda: dev.gemm(KC, MC, 8192, 8192, 8192, 1.0, a, a, 0.0, c))
print("strip=%s gemm8192 %.2f ms %.2f TF/s" % (os.environ.get("MUNK_GEMM_STRIP", "12"), ms, 2 * 8192 ** 3 / ms / 1e9))
