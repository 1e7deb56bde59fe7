"""Host enqueue time vs device time of munk_potrf / munk_trtri (is the host blocked by the launch
queue?).  usage: python tools/host_probe.py [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

from munk_b200.device import Device  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
dev = Device(0)
X = torch.randn(n, 512, dtype=torch.float64, device="cuda")
A0 = X @ X.t() / 512 + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")
del X
A = torch.empty_like(A0)
linv = dev.new_linv(n)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
stream = torch.cuda.Stream()       # not the legacy default stream: the factorisations replay CUDA graphs
torch.cuda.set_stream(stream)
for rep in range(4):
    A.copy_(A0)
    torch.cuda.synchronize()
    l0 = dev.stats()[0]
    a, b, c = ev(), ev(), ev()
    t0 = time.perf_counter()
    a.record()
    dev.potrf(A, n, linv=linv, check_info=False)
    b.record()
    t1 = time.perf_counter()
    l1 = dev.stats()[0]
    dev.trtri(A, n, linv)
    c.record()
    t2 = time.perf_counter()
    l2 = dev.stats()[0]
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("n=%d rep %d: potrf host enqueue %.1f ms (%d launches), device %.1f ms | trtri host %.1f ms (%d launches), device %.1f ms | "
          "total wall %.1f ms" % (n, rep, (t1 - t0) * 1e3, l1 - l0, a.elapsed_time(b), (t2 - t1) * 1e3, l2 - l1,
                                  b.elapsed_time(c), (t3 - t0) * 1e3), flush=True)
