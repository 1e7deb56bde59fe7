"""A small pass over every kernel family of libmunk_b200.so, sized for compute-sanitizer
(memcheck / racecheck / synccheck run 10-100x slower than native):

    compute-sanitizer --tool memcheck  python tools/sanitize_target.py
    compute-sanitizer --tool racecheck python tools/sanitize_target.py
    compute-sanitizer --tool synccheck python tools/sanitize_target.py

Covers: the TMA/mbarrier GEMM ring in all four operand layouts incl. split-K and the beta != 0
epilogue, the in-place TRSM leaf products, the one-CTA Cholesky leaf, the three-stream block-column
driver, the forked triangular inverse, left solves with a staircase, assembly / gather /
transpose / separate_scores kernels, the row-Jacobi pseudo-inverse, and the panel-major
distributed driver with one rank.  Results are checked against torch float64."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import networkx as nx  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

import munk_b200 as munk  # noqa: E402
from munk_b200 import util  # noqa: E402
from munk_b200.device import KC, MC, Device  # noqa: E402
from munk_b200.dist import DistFactor, staircase_rhs  # noqa: E402

os.environ.setdefault("MUNK_NO_GRAPH", "1")
dev = Device(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
rel = lambda a, b: float((a - b).abs().max() / b.abs().max())  # noqa: E731
g = torch.Generator(device="cpu").manual_seed(0)


def rnd(*shape):
    return torch.randn(*shape, dtype=torch.float64, generator=g).cuda()


# GEMM ring, four layouts, ragged sizes, beta != 0, split-K (few tiles, long K)
for (al, bl) in ((KC, KC), (KC, MC), (MC, KC), (MC, MC)):
    M, N, K = 200, 136, 1000
    A, B = rnd(M, K), rnd(N, K)
    Ad = dev.upload((A if al == KC else A.t().contiguous()).cpu().numpy())
    Bd = dev.upload((B if bl == KC else B.t().contiguous()).cpu().numpy())
    C0 = rnd(M, N)
    Cd = dev.upload(C0.cpu().numpy())
    dev.gemm(al, bl, M, N, K, 0.5, Ad, Bd, 2.0, Cd)
    assert rel(Cd[:M, :N], 0.5 * A @ B.t() + 2.0 * C0) < 1e-12
# Cholesky (leaf, recursion, three-stream driver), triangular inverse (forked), staircase solve
n = 1500
X = rnd(n, 64)
A0 = X @ X.t() / 64 + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")
A = dev.empty(n, n)
A[:, :n] = A0
linv = dev.potrf(A, n)
L = torch.tril(A[:, :n]).clone()
assert rel(L @ L.t(), A0) < 1e-12
Xs, stair, pad = staircase_rhs(n, list(range(512, 1024)), [0, 7, 900], dev.torch_device)
dev.trsm_left(A, n, linv, Xs, Xs.shape[1], transposed=False, stair=stair)
W = dev.trtri(A, n, linv)
Wl = torch.tril(W[:, :n])
assert rel(Wl @ L, torch.eye(n, dtype=torch.float64, device="cuda")) < 1e-11
assert rel(Xs[:, :512], Wl[:, 512:1024]) < 1e-11
# public API on a small graph pair (assembly, gathers, projection, scores, separate_scores)
G1, G2 = nx.barabasi_albert_graph(300, 4, seed=1), nx.barabasi_albert_graph(260, 4, seed=2)
G1 = nx.relabel_nodes(G1, {i: "s%03d" % i for i in G1})
G2 = nx.relabel_nodes(G2, {i: "t%03d" % i for i in G2})
homs = [("s%03d" % i, "t%03d" % (259 - i)) for i in range(0, 120, 3)]
(C1, sn), (C2, tn), l_idx, h_idx = munk.embed_networks(G1, G2, homs, 12, return_idxs=True)
S = munk.scores(C1, C2)
assert np.abs(C1 @ C2.T - S).max() < 1e-12
parts = munk.separate_scores(S, l_idx, h_idx)
assert len(parts) == 5 and len(parts[0]) == 12 and len(parts[4]) > 0
# robust pinv path (dependent landmark rows)
Cx = np.random.default_rng(1).standard_normal((40, 40))
Cx[9] = Cx[2] - Cx[5]
_, got = munk.embed_matrices(Cx, np.random.default_rng(2).standard_normal((30, 30)), [(2, 1), (5, 2), (9, 3), (11, 4)])
from munk_b200.device import default_device  # noqa: E402
assert default_device().last_rank == 3 and np.isfinite(got).all()
# panel-major distributed driver, one rank, fused forward substitution
Gd = nx.barabasi_albert_graph(1100, 5, seed=3)
u, v = util.edge_index_arrays(Gd, list(range(1100)))
fac = DistFactor(dev, None, 1100)
fac.assemble((u, v), 0.05)
Xd, std, padd = staircase_rhs(1100, list(range(512, 1024)), [3, 1000], dev.torch_device)
fac.set_rhs(Xd, Xd.shape[1], std)
fac.factor()
torch.cuda.synchronize()
assert dev.potrf_info() == 0
print("sanitize target: all checks passed", flush=True)
