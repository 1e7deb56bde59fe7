"""Device layer: the MUNK hot path as calls into libmunk_b200.so on CUDA tensors.

torch is used only as plumbing (device memory, streams, pinned host buffers); every kernel is
ours and is reached through the C ABI in include/munk_b200.h.  All device matrices are row-major
float64 with a padded leading dimension (multiple of 16 elements = 128 bytes, so TMA box rows
are line-aligned).

Factor convention: for an SPD matrix A = L L^T we keep W = L^-1 (lower triangular).  Then
D = A^-1 = W^T W, and C = W^T is the RKHS factor (C C^T = D) that replaces the reference's
eigendecomposition (munk/munk/munk.py:85-88).
"""
import ctypes

import numpy as np
import torch

from ._lib import MunkError, check, lib  # noqa: F401  (MunkError re-exported)

# structure flags of munk_dgemm (include/munk_b200.h)
KC, MC = 0, 1
TILE_LOWER, KLO_M, KLO_N, KHI_M, KHI_N, SYM_MIRROR = 1, 2, 4, 8, 16, 32


class NotPositiveDefinite(np.linalg.LinAlgError):
    pass


def pad16(n):
    return (int(n) + 15) // 16 * 16


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Device:
    """One CUDA device + the library workspace bound to torch's current stream on it."""

    def __init__(self, index=None):
        if not torch.cuda.is_available():
            raise RuntimeError("munk_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.index = torch.cuda.current_device() if index is None else int(index)
        self.torch_device = torch.device("cuda", self.index)
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.index):
            check(lib.munk_workspace_create(self.index, ctypes.byref(handle)), "munk_workspace_create")
        self._ws = handle

    def __del__(self):
        ws = getattr(self, "_ws", None)
        if ws:
            try:
                lib.munk_workspace_destroy(ws)
            except Exception:
                pass
            self._ws = None

    # ---- plumbing -------------------------------------------------------------------------
    def stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.index).cuda_stream)

    def empty(self, rows, cols, ld=None):
        """(rows x ld) buffer; the logical matrix is [:, :cols]."""
        ld = pad16(cols) if ld is None else ld
        return torch.empty((int(rows), int(ld)), dtype=torch.float64, device=self.torch_device)

    def ints(self, values):
        """int32 CUDA tensor from a list / numpy array / (pinned) host tensor / CUDA tensor."""
        if torch.is_tensor(values):
            if values.is_cuda:
                return values
            return values.to(self.torch_device, non_blocking=True)
        arr = np.ascontiguousarray(np.asarray(values, dtype=np.int32))
        return torch.from_numpy(arr).to(self.torch_device, non_blocking=False)

    def stats(self, reset=False):
        launches, flops = ctypes.c_long(), ctypes.c_double()
        check(lib.munk_workspace_stats(self._ws, ctypes.byref(launches), ctypes.byref(flops),
                                       1 if reset else 0), "munk_workspace_stats")
        return launches.value, flops.value

    def upload(self, array, cols=None):
        """Host (rows x cols) float64 array -> padded device buffer."""
        a = np.asarray(array, dtype=np.float64)
        rows, cols = a.shape
        buf = self.empty(rows, cols)
        if buf.shape[1] != cols:
            buf[:, cols:].zero_()
        buf[:, :cols].copy_(torch.from_numpy(np.ascontiguousarray(a)))
        return buf

    def download(self, buf, rows, cols, out=None):
        """Device [:rows, :cols] -> fresh host ndarray, or into `out` (e.g. a reusable pinned
        tensor).  Fresh results use pageable memory: pinning 3.2 GB takes ~2.5 s on the B200 hosts
        (measured), more than the copy it would speed up; callers that repeat use a HostSink."""
        host = out if out is not None else torch.empty((rows, cols), dtype=torch.float64)
        host.copy_(buf[:rows, :cols], non_blocking=True)
        torch.cuda.current_stream(self.index).synchronize()
        return host.numpy()

    def download_fresh(self, buf, rows, cols, upper=False, threads=None):
        """Device [:rows, :cols] -> a fresh (pageable) host ndarray through the multi-threaded
        pinned staging path (munk_download_pageable): ~20-30 GB/s instead of the ~3 GB/s a plain
        copy into untouched pageable memory reaches.  upper=True: `buf` is upper triangular (the
        factor C1), the zero part is filled on the host instead of being copied.  Blocking."""
        import os
        out = np.empty((int(rows), int(cols)), dtype=np.float64)
        if rows and cols:
            if threads is None:
                threads = max(2, min(12, (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else 8) - 2))
            check(lib.munk_download_pageable(self.index, int(rows), int(cols), _ptr(buf), buf.stride(0),
                                             out.ctypes.data_as(ctypes.c_void_p), int(cols), 1 if upper else 0,
                                             int(threads), self.stream()), "munk_download_pageable")
        return out

    # ---- kernels ----------------------------------------------------------------------------
    def gemm(self, a_layout, b_layout, M, N, K, alpha, A, B, beta, C, flags=0):
        check(lib.munk_dgemm(self._ws, a_layout, b_layout, M, N, K, alpha, _ptr(A), A.stride(0),
                             _ptr(B), B.stride(0), beta, _ptr(C), C.stride(0), flags, self.stream()),
              "munk_dgemm")

    def assemble(self, n, edges_u, edges_v, lam, out=None, lval=None, ldiag=None):
        """A = I + lam*L from int32 edge index arrays (host or device).  lval/ldiag (float64, host
        or device): the Laplacian's off-diagonal / diagonal entries for weighted graphs and
        multigraphs (util.edge_index_arrays returns them as a 4-tuple)."""
        u, v = self.ints(edges_u), self.ints(edges_v)
        A = self.empty(n, n) if out is None else out
        if lval is None:
            check(lib.munk_assemble_reglap(self._ws, n, int(u.numel()), _ptr(u), _ptr(v), float(lam),
                                           _ptr(A), A.stride(0), self.stream()), "munk_assemble_reglap")
        else:
            w, d = self.doubles(lval), self.doubles(ldiag)
            assert int(w.numel()) == int(u.numel()) and int(d.numel()) == n
            check(lib.munk_assemble_reglap_entries(self._ws, n, int(u.numel()), _ptr(u), _ptr(v), _ptr(w),
                                                   _ptr(d), float(lam), _ptr(A), A.stride(0), self.stream()),
                  "munk_assemble_reglap_entries")
        return A

    def assemble_edges(self, n, edges, lam, out=None):
        """assemble() for an `edges` tuple: (u, v) or (u, v, lval, ldiag)."""
        if len(edges) == 2:
            return self.assemble(n, edges[0], edges[1], lam, out=out)
        return self.assemble(n, edges[0], edges[1], lam, out=out, lval=edges[2], ldiag=edges[3])

    def assemble_info(self):
        """1 + index of the first edge the assembly skipped as out of range (0 = none); synchronises."""
        bad = ctypes.c_int(0)
        check(lib.munk_assemble_info(self._ws, ctypes.byref(bad), self.stream()), "munk_assemble_info")
        return bad.value

    def doubles(self, values):
        if torch.is_tensor(values):
            return values if values.is_cuda else values.to(self.torch_device, non_blocking=True)
        arr = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        return torch.from_numpy(arr).to(self.torch_device)

    def new_linv(self, n):
        return torch.empty(lib.munk_linv_bytes(int(n)) // 8, dtype=torch.float64, device=self.torch_device)

    def potrf(self, A, n, linv=None, check_info=True):
        """In-place lower Cholesky; returns the leaf-inverse buffer trtri needs."""
        linv = self.new_linv(n) if linv is None else linv
        check(lib.munk_potrf(self._ws, n, _ptr(A), A.stride(0), _ptr(linv), self.stream()), "munk_potrf")
        if check_info:
            self.raise_if_not_spd()
        return linv

    def potrf_pair(self, A, n, linv, other, B, nb, linvb):
        """In-place Cholesky of A (this workspace) and B (`other`'s workspace) with their block
        columns interleaved on the current stream (munk_potrf_pair); flags are checked separately."""
        check(lib.munk_potrf_pair(self._ws, n, _ptr(A), A.stride(0), _ptr(linv), other._ws, nb, _ptr(B),
                                  B.stride(0), _ptr(linvb), self.stream()), "munk_potrf_pair")

    def potrf_info(self):
        info = ctypes.c_int(0)
        check(lib.munk_potrf_info(self._ws, ctypes.byref(info), self.stream()), "munk_potrf_info")
        return info.value

    def potrf_info_async(self):
        """The status flag moved into a fresh 1-element int32 CUDA tensor (flag cleared); no sync."""
        out = torch.empty(1, dtype=torch.int32, device=self.torch_device)
        check(lib.munk_potrf_info_async(self._ws, _ptr(out), self.stream()), "munk_potrf_info_async")
        return out

    def raise_if_not_spd(self):
        info = self.potrf_info()
        if info:
            raise NotPositiveDefinite("matrix is not positive definite (pivot %d)" % (info - 1))

    def trtri(self, L, n, linv):
        check(lib.munk_trtri(self._ws, n, _ptr(L), L.stride(0), _ptr(linv), self.stream()), "munk_trtri")
        return L

    def inverse_factor(self, A, n, check_info=True):
        """A (SPD, overwritten) -> W = chol(A)^-1, lower triangular, in place."""
        linv = self.potrf(A, n, check_info=check_info)
        return self.trtri(A, n, linv)

    def gram_lower(self, W, n, out=None, ld=None):
        """D = W^T W (full symmetric)."""
        D = self.empty(n, n, ld) if out is None else out
        check(lib.munk_gram_lower(self._ws, n, _ptr(W), W.stride(0), _ptr(D), D.stride(0), self.stream()),
              "munk_gram_lower")
        return D

    def gather_cols_lower(self, W, n, idx):
        """Q (n x pad16(k)) with Q[:, l] = column idx[l] of the lower-triangular W."""
        idx_d = idx if torch.is_tensor(idx) else self.ints(idx)
        k = int(idx_d.numel())
        Q = self.empty(n, k)
        check(lib.munk_gather_cols_lower(n, _ptr(W), W.stride(0), k, _ptr(idx_d), _ptr(Q), Q.stride(0),
                                         self.stream()), "munk_gather_cols_lower")
        return Q

    def transpose_lower(self, W, n, ld=None):
        C = self.empty(n, n, ld)
        check(lib.munk_transpose_lower(n, _ptr(W), W.stride(0), _ptr(C), C.stride(0), self.stream()),
              "munk_transpose_lower")
        return C

    def copy_lower(self, L, n, ld=None):
        out = self.empty(n, n, ld)
        check(lib.munk_copy_lower(n, _ptr(L), L.stride(0), _ptr(out), out.stride(0), self.stream()),
              "munk_copy_lower")
        return out

    def download_upper(self, C, n, host, row_block=1024):
        """host[i, j] <- C[i, j] for the upper-triangular C (device), in row blocks that skip the
        zero part left of each block's first column; `host` (a CPU tensor whose strictly lower
        part is already zero) receives half the bytes of a full copy.  Enqueued on the current
        stream; asynchronous if `host` is pinned."""
        assert host.device.type == "cpu" and host.dtype == torch.float64 and host.stride(1) == 1
        check(lib.munk_download_upper(n, _ptr(C), C.stride(0), host.data_ptr(), host.stride(0),
                                      row_block, self.stream()), "munk_download_upper")

    def download_block(self, src, host):
        """host (CPU view, unit column stride) <- src (CUDA view of the same shape, unit column
        stride): one strided DMA on the current stream, asynchronous if `host` is pinned."""
        assert host.device.type == "cpu" and host.dtype == torch.float64 and src.shape == host.shape
        assert host.dim() == 2 and (host.shape[1] <= 1 or (host.stride(1) == 1 and src.stride(1) == 1))
        check(lib.munk_download_block(host.shape[0], host.shape[1], _ptr(src), src.stride(0),
                                      host.data_ptr(), host.stride(0), self.stream()), "munk_download_block")

    def group_mean_rows(self, B, ncols, gptr, members):
        g = len(gptr) - 1
        out = self.empty(g, ncols)
        gp, mem = self.ints(gptr), self.ints(members)
        check(lib.munk_group_mean_rows(g, _ptr(gp), _ptr(mem), ncols, _ptr(B), B.stride(0), _ptr(out),
                                       out.stride(0), self.stream()), "munk_group_mean_rows")
        return out

    def gather_pairs(self, S, rows, cols):
        r, c = self.ints(rows), self.ints(cols)
        out = torch.empty(int(r.numel()), dtype=torch.float64, device=self.torch_device)
        check(lib.munk_gather_pairs(int(r.numel()), _ptr(r), _ptr(c), _ptr(S), S.stride(0), _ptr(out),
                                    self.stream()), "munk_gather_pairs")
        return out

    # ---- the hot path, device resident -------------------------------------------------------
    def kernel_factor(self, n, edges_u, edges_v, lam):
        """W with D = (I + lam L)^-1 = W^T W  (assembly + Cholesky + triangular inverse)."""
        A = self.assemble(n, edges_u, edges_v, lam)
        return self.inverse_factor(A, n)

    def kernel_factor_edges(self, n, edges, lam):
        return self.inverse_factor(self.assemble_edges(n, edges, lam), n)

    # ---- landmark projection: np.linalg.pinv(source_C[source_idxs,:]).dot(target_D[target_idxs,:]) --
    COND_LIMIT = 1e6      # largest cond(M M^T) the Gram/Cholesky path may be used for (error ~ cond * eps)

    def project(self, Q1, n1, B, n2, groups, defer=False):
        """P = C2hat^T = pinv(M) B with M^T = Q1 (n1 x k'), B = D2[Lt,:] (k x n2)  (munk.py:108-110).

        `groups` = (gptr, members) folds rows of B that belong to the same (duplicated) source
        landmark: np.linalg.pinv's least-squares / minimum-norm answer for duplicate rows of M
        equals the answer for the de-duplicated M with the rows of B averaged.

        Fast path, always enqueued: G = M M^T, Cholesky, P = M^T G^-1 B.  With distinct rows of a
        kernel factor G = D1[Ls,Ls] is a principal block of an SPD matrix, cond(G) <= cond(D1), and
        this IS the pseudo-inverse.  Whether it was safe is decided on the device, without a host
        synchronisation: `bad` = Cholesky failed, or the rigorous bound
        |G|_F |G^-1|_F >= cond_2(G) exceeds COND_LIMIT.  If it was not, the robust path
        (munk_pinv_rows_solve: one-sided Jacobi on the rows of M, numpy's rcond * sigma_max cut)
        recomputes P.  defer=False checks the flag now (one synchronisation) and returns P;
        defer=True returns (P, check) and the caller calls check.resolve() at its own
        synchronisation point (it returns the final P)."""
        Z, kq, check = self.landmark_solve(Q1, n1, B, n2, groups)
        P = self.empty(n1, n2)
        self.gemm(KC, MC, n1, n2, kq, 1.0, Q1, Z, 0.0, P, 0)           # P = M^T Z
        check.P = P
        if defer:
            return P, check
        return check.resolve()

    def landmark_solve(self, Q1, n1, B, n2, groups):
        """Fast path only: Z = (M M^T)^-1 mean_groups(B) (k' x n2), k' and the ProjectionCheck that
        says whether Z may be used (shared by `project` and the rank-k statistic S = D1[:,Ls] Z)."""
        gptr, members = groups
        kq = len(gptr) - 1
        mult = [gptr[g + 1] - gptr[g] for g in range(kq)]
        if max(mult) > 1 or list(members) != list(range(kq)):
            Bm = self.group_mean_rows(B, n2, gptr, members)
        else:
            Bm = B
        G = self.empty(kq, kq)
        self.gemm(MC, MC, kq, kq, n1, 1.0, Q1, Q1, 0.0, G, 0)          # G = M M^T
        gnorm = torch.linalg.norm(G[:kq, :kq])
        linv = self.potrf(G, kq, check_info=False)
        Wg = self.trtri(G, kq, linv)
        pot = self.potrf_info_async()
        winv = torch.tril(Wg[:kq, :kq])
        bound = gnorm * (winv * winv).sum()                           # |G|_F |G^-1|_F >= cond_2(G)
        bad = torch.logical_or(torch.logical_not(bound < self.COND_LIMIT), pot[0] != 0)
        Y = self.empty(kq, n2)
        self.gemm(KC, MC, kq, n2, kq, 1.0, Wg, Bm, 0.0, Y, KHI_M)      # Y = Wg B
        Z = self.empty(kq, n2)
        self.gemm(MC, MC, kq, n2, kq, 1.0, Wg, Y, 0.0, Z, KLO_M)       # Z = Wg^T Y
        self.last_rank = kq
        return Z, kq, ProjectionCheck(self, bad, Q1, n1, Bm, n2, kq, mult if max(mult) > 1 else None)

    def pinv_rows_solve(self, R, k, m, B, ncols, rcond=1e-15):
        """Robust path: R (k x m, landmark rows, overwritten with the rotated rows J M) and
        Z = diag(sigma^+2) J B such that pinv(M) B = R^T Z.  Returns (Z, rank, sigma)."""
        Z = self.empty(k, ncols)
        work = torch.empty(lib.munk_pinv_rows_work_bytes(k) // 8 + 2, dtype=torch.float64,
                           device=self.torch_device)
        rank = ctypes.c_int(0)
        sigma = np.zeros(max(k, 1), dtype=np.float64)
        check(lib.munk_pinv_rows_solve(self._ws, k, m, _ptr(R), R.stride(0), ncols, _ptr(B), B.stride(0),
                                       _ptr(Z), Z.stride(0), _ptr(work), float(rcond), ctypes.byref(rank),
                                       sigma.ctypes.data_as(ctypes.c_void_p), self.stream()),
              "munk_pinv_rows_solve")
        self.last_rank = rank.value
        return Z, rank.value, sigma[:k]

    def project_robust(self, Q1, n1, Bm, n2, kq, weights=None, rcond=1e-15):
        """P = pinv(M) B through munk_pinv_rows_solve (numpy semantics; see `project`)."""
        R = self.transpose(Q1, n1, kq)                                 # rows of M, contiguous
        Bw = Bm
        if weights is not None:
            # A source landmark listed c times counts c-fold in pinv's least-squares objective:
            # pinv(M_all) B_all = pinv(D M) (D mean(B)) with D = diag(sqrt(c)).  (With full row rank
            # the residual is zero and D drops out, which is why the fast path ignores it.)
            d = torch.sqrt(torch.as_tensor(np.asarray(weights, dtype=np.float64), device=self.torch_device))
            R[:kq] *= d[:, None]
            Bw = Bm.clone()
            Bw[:kq] *= d[:, None]
        Z, _, _ = self.pinv_rows_solve(R, kq, n1, Bw, n2, rcond)
        P = self.empty(n1, n2)
        self.gemm(MC, MC, n1, n2, kq, 1.0, R, Z, 0.0, P, 0)            # P = R^T Z
        return P

    def pair_dots(self, X, Z, k, rows, cols):
        """out[p] = X[rows[p], :k] . Z[:k, cols[p]]"""
        r, c = self.ints(rows), self.ints(cols)
        out = torch.empty(int(r.numel()), dtype=torch.float64, device=self.torch_device)
        check(lib.munk_pair_dots(int(r.numel()), _ptr(r), _ptr(c), k, _ptr(X), X.stride(0), _ptr(Z),
                                 Z.stride(0), _ptr(out), self.stream()), "munk_pair_dots")
        return out

    def trsm_left(self, L, n, linv, B, nrhs, transposed, stair=None):
        """B[:n, :nrhs] <- L^-1 B or L^-T B in place (factor + leaf inverses from potrf).
        `stair`: int32 CUDA tensor, first non-zero row of every 128-column tile of B (forward only)."""
        check(lib.munk_trsm_left(self._ws, n, _ptr(L), L.stride(0), _ptr(linv), 1 if transposed else 0,
                                 nrhs, _ptr(B), B.stride(0), _ptr(stair) if stair is not None else None,
                                 self.stream()), "munk_trsm_left")
        return B

    def trsm_right(self, B, m, L, s, linv):
        """B[:m, :s] <- B L^-T in place (panel solve)."""
        check(lib.munk_trsm_right(self._ws, m, _ptr(B), B.stride(0), _ptr(L), L.stride(0), s, _ptr(linv),
                                  self.stream()), "munk_trsm_right")
        return B

    def gemm_stair(self, a_layout, b_layout, M, N, K, alpha, A, B, beta, C, flags=0, stair_m=None,
                   stair_n=None, row0=0, k0=0):
        check(lib.munk_dgemm_stair(self._ws, a_layout, b_layout, M, N, K, alpha, _ptr(A), A.stride(0),
                                   _ptr(B), B.stride(0), beta, _ptr(C), C.stride(0), flags,
                                   _ptr(stair_m) if stair_m is not None else None,
                                   _ptr(stair_n) if stair_n is not None else None, row0, k0,
                                   self.stream()), "munk_dgemm_stair")

    def transpose(self, X, rows, cols, out=None):
        """out (cols x rows) = X[:rows, :cols]^T."""
        out = self.empty(cols, rows) if out is None else out
        check(lib.munk_transpose(rows, cols, _ptr(X), X.stride(0), _ptr(out), out.stride(0),
                                 self.stream()), "munk_transpose")
        return out

    def target_landmark_rows_from_factor(self, L2, linv2, n2, target_idx):
        """B = D2[Lt, :] = (A2^-1 E_Lt)^T from the Cholesky factor alone: two triangular solves
        with k right-hand sides (2 n2^2 k flops) instead of the full inverse (2 n2^3 / 3)."""
        if torch.is_tensor(target_idx):
            k, rows = int(target_idx.numel()), target_idx.to(self.torch_device).long()
        else:
            k = len(target_idx)
            rows = torch.as_tensor(np.asarray(target_idx, dtype=np.int64), device=self.torch_device)
        X = torch.zeros((n2, pad16(k)), dtype=torch.float64, device=self.torch_device)
        cols = torch.arange(k, device=self.torch_device)
        X[rows, cols] = 1.0
        self.trsm_left(L2, n2, linv2, X, k, transposed=False)
        self.trsm_left(L2, n2, linv2, X, k, transposed=True)
        return self.transpose(X, n2, k)

    def target_landmark_rows(self, W2, n2, target_idx):
        """B = D2[Lt, :] = (W2[:, Lt])^T W2 without forming D2 (k x n2)."""
        Q2 = self.gather_cols_lower(W2, n2, target_idx)
        k = len(target_idx)
        B = self.empty(k, n2)
        self.gemm(MC, MC, k, n2, n2, 1.0, Q2, W2, 0.0, B, KLO_N)
        return B

    def score(self, W1, n1, P, n2, out=None, ld=None):
        """S = C1 C2hat^T = W1^T P; only k >= row of the triangular W1 is contracted."""
        S = self.empty(n1, n2, ld) if out is None else out
        self.gemm(MC, MC, n1, n2, n1, 1.0, W1, P, 0.0, S, KLO_M)
        return S


class ProjectionCheck:
    """Device-side verdict on the fast landmark solve (Device.project).  resolve() reads the flag
    (synchronising the stream once) and returns the projection to use: the fast one, or the
    robust recomputation written into the same buffer."""

    def __init__(self, dev, bad, Q1, n1, Bm, n2, kq, weights):
        self.dev, self.bad, self.P = dev, bad, None
        self.args = (Q1, n1, Bm, n2, kq, weights)

    def fallback_needed(self):
        return bool(self.bad.item())

    def resolve(self):
        if self.fallback_needed():
            P = self.dev.project_robust(*self.args)
            if self.P is not None:
                self.P.copy_(P)
                return self.P
            return P
        return self.P


def landmark_groups(source_idx):
    """First-occurrence de-duplication of the source landmark list.

    Returns (unique_source_idx, gptr, members): group g collects the positions l of
    source_idx with source_idx[l] == unique_source_idx[g], in list order.
    """
    first = {}
    buckets = []
    for pos, s in enumerate(source_idx):
        g = first.get(s)
        if g is None:
            first[s] = len(buckets)
            buckets.append([pos])
        else:
            buckets[g].append(pos)
    uniq = [source_idx[b[0]] for b in buckets]
    gptr = [0]
    members = []
    for b in buckets:
        members.extend(b)
        gptr.append(len(members))
    return uniq, gptr, members


_side = {}


def side_device(dev):
    """(second Device, high-priority CUDA stream, normal-priority CUDA stream) on the same GPU, for
    running two independent chains of work concurrently (own workspace, so split-K scratch,
    look-ahead streams, graph cache and status flag are not shared).  The block scheduler serves
    streams by priority, then age: the chain on the high-priority stream keeps its speed, the
    other one fills the SMs it leaves idle.  Both are explicit streams (not torch's default
    stream, which is the legacy NULL stream and cannot be captured into a CUDA graph).
    Created once per device."""
    res = _side.get(dev.index)
    if res is None:
        import os
        res = _side[dev.index] = (Device(dev.index),
                                  torch.cuda.Stream(device=dev.index, priority=int(os.environ.get("MUNK_SRC_PRIORITY", "-1"))),
                                  torch.cuda.Stream(device=dev.index, priority=int(os.environ.get("MUNK_TGT_PRIORITY", "0"))))
    return res


_default = {}


def default_device(index=None):
    """Process-wide Device per CUDA ordinal (created on first use)."""
    idx = torch.cuda.current_device() if index is None else int(index)
    dev = _default.get(idx)
    if dev is None:
        dev = _default[idx] = Device(idx)
    return dev
