"""Times the trailing-update shape of the Cholesky driver (M = N large, small K) in isolation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from munk_b200.device import Device, KC, TILE_LOWER

dev = Device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16928
C = torch.zeros(n, n, dtype=torch.float64, device="cuda")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for K in (256, 512, 1024, 2048):
    X = torch.randn(n, K, dtype=torch.float64, device="cuda")
    for name, beta, flags in (("lower beta=1", 1.0, TILE_LOWER), ("lower beta=0", 0.0, TILE_LOWER),
                              ("full  beta=1", 1.0, 0), ("full  beta=0", 0.0, 0)):
        ms = timed(lambda: dev.gemm(KC, KC, n, n, K, -1.0, X, X, beta, C, flags))
        tiles = (n // 128) * (n // 128 + 1) / 2 if flags else (n / 128) ** 2
        fl = 2.0 * tiles * 128 * 128 * K
        print("K=%4d %s: %7.3f ms  %5.2f TF/s  (%.1f us per wave of 148 tiles)" % (
            K, name, ms, fl / ms / 1e9, ms * 1e3 / (tiles / 148)), flush=True)
