"""ncu target for the one-launch diagonal-block kernel (potrf_diag_kernel, 512 columns) inside a real
factorisation: Cholesky of a 3072-column matrix with 512-column panels, direct launches."""
import os
import sys

os.environ["MUNK_NO_GRAPH"] = "1"
os.environ["MUNK_PANEL_SWITCH"] = "0"      # 512-column panels down to the end
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from munk_b200.device import Device

n = 3072
dev = Device(0)
torch.cuda.set_stream(torch.cuda.Stream())
X = torch.randn(n, 64, dtype=torch.float64, device="cuda")
A0 = X @ X.t() / 64 + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")
A = dev.empty(n, n)
A[:, :n] = A0
linv = dev.potrf(A, n)
L = torch.tril(A[:, :n])
torch.cuda.synchronize()
print("ok", float((L @ L.t() - A0).abs().max()))
