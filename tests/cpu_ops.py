"""torch-CPU implementation of the `ops` protocol of munk_b200.dist.dist_potrf_steps (the schedule specification) — TEST INFRASTRUCTURE ONLY.

Lets the multi-rank schedule of munk_b200.dist.dist_potrf (ownership map, look-ahead order,
packing, broadcasts) run under gloo on CPU tensors.  It is never imported by the product."""
import torch


class TorchCpuOps:
    def __init__(self):
        self.device = torch.device("cpu")
        self.bad = 0

    def syrk_lower(self, C, X, Y, m, nc, k):
        C[:m, :nc] -= X[:m, :k] @ Y[:nc, :k].t()

    def factor_panel(self, A, n, c0, w, linv):
        blk = torch.tril(A[c0:c0 + w, c0:c0 + w])
        sym = blk + blk.t() - torch.diag(torch.diagonal(blk))
        L, info = torch.linalg.cholesky_ex(sym)
        if int(info) and not self.bad:
            self.bad = c0 + int(info)
        A[c0:c0 + w, c0:c0 + w] = L
        if n > c0 + w:
            B = A[c0 + w:, c0:c0 + w]
            A[c0 + w:, c0:c0 + w] = torch.linalg.solve_triangular(L, B.t(), upper=False).t()

    def info(self):
        return self.bad
