import os, sys, ctypes
sys.path.insert(0, "/root/repo")
import torch
from munk_b200.device import Device
dev = Device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = torch.Generator(device="cpu").manual_seed(1)
X = torch.randn(n, n, dtype=torch.float64, generator=g).cuda()
A0 = X @ X.t() / n + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")
del X
A = A0.clone()
linv = dev.new_linv(n)
dev.potrf(A, n, linv=linv)            # warm-up
os.environ["MUNK_TRACE_POTRF"] = "1"
A.copy_(A0)
torch.cuda.synchronize()
dev.potrf(A, n, linv=linv)
