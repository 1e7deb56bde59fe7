"""Small fixed workload for `ncu --set full`: one launch of every kernel the bench's roofline talks
about, at C3 sizes, in this order (so `-k regex:... -c N` captures them by position):

  1. dgemm_dmma_tma_kernel<MC,MC>  the score product S = W1^T P, 20000 x 16000 x 20000, k >= row only
  2. dgemm_dmma_tma_kernel<KC,KC>  a K = 512 trailing update of the Cholesky driver (11 000 rows, lower tiles)
  3. dgemm_dmma_tma_kernel<KC,KC>  a K = 1024 trailing update (16 000 rows, lower tiles)
  4. dgemm_dmma_tma_kernel<KC,MC>  a trtri product T = L21 W11 (5000 x 5000 x 5000, k >= column)
  5. trsm_strip_kernel             the chain's panel solve, 512 rows x 512
  6. potrf_leaf_kernel             one 128 x 128 Cholesky + inverse
  7. fill_reglap_kernel            Laplacian assembly of the C3 source (20000^2)
  8. separate_scores_kernel, masked_row_sum_kernel   the two passes over S of separate_scores / its means
Run plain first (must exit 0), then under ncu (tools/gpu_ncu.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from munk_b200.device import Device, KC, MC, KLO_N, TILE_LOWER
from munk_b200.scores_plan import plan_separate_scores

os.environ["MUNK_NO_GRAPH"] = "1"
n1 = int(os.environ.get("N1", 20000))
n2 = int(os.environ.get("N2", 16000))
dev = Device(0)
torch.manual_seed(0)
W = torch.tril(torch.randn(n1, n1, dtype=torch.float64, device="cuda")) / n1 ** 0.5
P = torch.randn(n1, n2, dtype=torch.float64, device="cuda")
S = dev.empty(n1, n2)
dev.score(W, n1, P, n2, out=S)                                                         # 1
m = 11000
C = dev.empty(m, m)
dev.gemm(KC, KC, m, m, 512, -1.0, W[n1 - m:, :512], W[n1 - m:, :512], 1.0, C, TILE_LOWER)       # 2
m = 16000
C = dev.empty(m, m)
dev.gemm(KC, KC, m, m, 1024, -1.0, W[n1 - m:, :1024], W[n1 - m:, :1024], 1.0, C, TILE_LOWER)    # 3
h = 5000
T = dev.empty(h, h)
dev.gemm(KC, MC, h, h, h, 1.0, W[h:2 * h, :h], W[:h, :h], 0.0, T, KLO_N)                        # 4
del C, T
X = torch.randn(640, 64, dtype=torch.float64, device="cuda")
A0 = X @ X.t() / 64 + 2.0 * torch.eye(640, dtype=torch.float64, device="cuda")
A = dev.empty(640, 640)
A[:, :640] = A0
torch.cuda.synchronize()
sub = A[:128, :128].clone()
lin128 = dev.new_linv(128)
blk = A[:512, :512].clone()
lin = dev.potrf(blk, 512)                                                              # (leaves inside: not the captured one)
B = torch.randn(512, 512, dtype=torch.float64, device="cuda")
dev.trsm_right(B, 512, blk, 512, lin)                                                  # 5
dev.potrf(sub, 128, linv=lin128)                                                       # 6
import networkx as nx
from munk_b200 import util
G = nx.barabasi_albert_graph(n1, 10, seed=1)
u, v = util.edge_index_arrays(G, list(range(n1)))
A = dev.assemble(n1, u, v, 0.05)                                                       # 7
rng = np.random.default_rng(0)
hs, ht = rng.permutation(n1)[:4000], rng.permutation(n2)[:4000]
hom = list(zip(hs.tolist(), ht.tolist()))
plan = plan_separate_scores(n1, n2, hom[:400], hom)
out = plan.run(dev, S)                                                                 # 8a
means = plan.means(dev, S)                                                             # 8b
torch.cuda.synchronize()
print("ok", float(means[0]))
