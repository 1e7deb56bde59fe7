"""Does interleaving the block columns of two independent Cholesky factorisations on ONE trailing-
update stream hide their latency-bound chains behind each other's dense work?  Uses the stepwise
driver of csrc/dist.cu with one rank.  usage: python tools/dual_probe.py [n1 n2]"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import networkx as nx  # noqa: E402
import torch  # noqa: E402

from munk_b200 import util  # noqa: E402
from munk_b200.device import Device  # noqa: E402
from munk_b200.dist import DistFactor  # noqa: E402

n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
n2 = int(sys.argv[2]) if len(sys.argv) > 2 else 16000
dev, dev2 = Device(0), Device(0)
e1 = util.edge_index_arrays(nx.barabasi_albert_graph(n1, 10, seed=1), list(range(n1)))
e2 = util.edge_index_arrays(nx.barabasi_albert_graph(n2, 10, seed=2), list(range(n2)))
e1 = tuple(dev.ints(x) for x in e1)
e2 = tuple(dev.ints(x) for x in e2)
f1, f2 = DistFactor(dev, None, n1), DistFactor(dev2, None, n2)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731


def timed(fn, reps=3):
    best = 1e9
    for _ in range(reps):
        f1.assemble(e1, 0.05)
        f2.assemble(e2, 0.05)
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    assert dev.potrf_info() == 0 and dev2.potrf_info() == 0
    return best


def sequential():
    f1.factor()
    f2.factor()


def interleaved(ratio):
    """one target column after every `ratio` source columns (both on the same caller stream)"""
    f1.begin()
    f2.begin()
    l1 = l2 = 1
    i = 0
    while l1 or l2:
        if l1:
            l1 = f1.step()
        i += 1
        if l2 and (i % ratio == 0 or not l1):
            l2 = f2.step()
    f1.end()
    f2.end()


def only(f):
    return lambda: f.factor()


t1 = timed(only(f1))
t2 = timed(only(f2))
ts = timed(sequential)
print("alone: n1=%d %.1f ms, n2=%d %.1f ms; back to back %.1f ms" % (n1, t1, n2, t2, ts), flush=True)
for ratio in (1, 2):
    print("interleaved, 1 target column per %d source columns: %.1f ms" % (ratio, timed(lambda: interleaved(ratio))),
          flush=True)
