"""Times munk_potrf and munk_trtri in isolation (CUDA events, best of 3) and checks the factor:
usage: python tools/factor_probe.py [n] [--trace].  Environment knobs of csrc/factor.cu
(MUNK_PANEL_BIG, MUNK_PANEL_BIG_UNTIL, MUNK_PANEL, MUNK_PANEL_SMALL, MUNK_PANEL_SWITCH,
MUNK_TRTRI_FORK_DEPTH) are read by the library at first use."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from munk_b200.device import Device  # noqa: E402

n = int(next((a for a in sys.argv[1:] if a.isdigit()), "20000"))
dev = Device(0)
g = torch.Generator(device="cpu").manual_seed(1)
X = torch.randn(n, 512, dtype=torch.float64, generator=g).cuda()
A0 = X @ X.t() / 512 + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")
del X
A = torch.empty_like(A0)
linv = dev.new_linv(n)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
best_p = best_t = 1e9
for rep in range(4):
    A.copy_(A0)
    torch.cuda.synchronize()
    a, b, c = ev(), ev(), ev()
    a.record()
    dev.potrf(A, n, linv=linv, check_info=False)
    b.record()
    if rep == 0:
        L = torch.tril(A).clone()
    dev.trtri(A, n, linv)
    c.record()
    torch.cuda.synchronize()
    if rep:
        best_p, best_t = min(best_p, a.elapsed_time(b)), min(best_t, b.elapsed_time(c))
assert dev.potrf_info() == 0
x = torch.randn(n, 8, dtype=torch.float64, device="cuda")
r1 = (L @ (L.t() @ x) - A0 @ x).abs().max().item() / (A0 @ x).abs().max().item()
W = torch.tril(A)
r2 = (W @ (L @ x) - x).abs().max().item() / x.abs().max().item()
print("n=%d potrf %.2f ms (%.1f TF/s)  trtri %.2f ms (%.1f TF/s)  |LL^T x - A x| %.1e  |W L x - x| %.1e  env %s" % (
    n, best_p, n ** 3 / 3 / best_p / 1e9, best_t, n ** 3 / 3 / best_t / 1e9, r1, r2,
    {k: v for k, v in os.environ.items() if k.startswith("MUNK_")}), flush=True)
if "--trace" in sys.argv:
    os.environ["MUNK_TRACE_POTRF"] = "1"
    A.copy_(A0)
    torch.cuda.synchronize()
    dev.potrf(A, n, linv=linv)
