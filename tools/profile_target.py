"""Small fixed workload for `ncu --set full`: the dominant launches of one C3 pipeline step.

  1. the score product S = W1^T P at C3 size (dgemm_dmma_tma_kernel<MC,MC>, k >= row only)
  2. one trailing SYRK update of the Cholesky recursion (dgemm_dmma_tma_kernel<KC,KC>, lower tiles)
  3. the Laplacian fill (fill_reglap_kernel) and one potrf leaf (potrf_leaf_kernel)
Run plain first, then under ncu with -k regex:... (see profiles/README.md).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from munk_b200.device import Device, KC, MC, KLO_M, TILE_LOWER

n1 = int(os.environ.get("N1", 20000))
n2 = int(os.environ.get("N2", 16000))
dev = Device(0)
torch.manual_seed(0)
W = torch.tril(torch.randn(n1, n1, dtype=torch.float64, device="cuda")) / n1 ** 0.5
P = torch.randn(n1, n2, dtype=torch.float64, device="cuda")
S = dev.empty(n1, n2)
for _ in range(2):
    dev.score(W, n1, P, n2, out=S)
h = n1 // 2
dev.gemm(KC, KC, h, h, h, -1.0, W[h:, :h], W[h:, :h], 1.0, S[:h, :h], TILE_LOWER)
import networkx as nx
from munk_b200 import util
G = nx.barabasi_albert_graph(n1, 10, seed=1)
u, v = util.edge_index_arrays(G, list(range(n1)))
A = dev.assemble(n1, u, v, 0.05)
sub = A[:128, :128].clone()
lin = dev.new_linv(128)
dev.potrf(sub, 128, linv=lin)
torch.cuda.synchronize()
print("ok")
