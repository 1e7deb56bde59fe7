/*
 * munk_b200 - C ABI of the B200 (sm_100a) implementation of MUNK's dense FP64 hot path.
 *
 * The reference (lrgr/munk) is pure Python and has no FFI layer; its boundary for this path is
 * the `munk` package API (munk/munk/__init__.py:1-11).  Each entry point below names the
 * reference call it replaces.  A binding (ctypes, cffi, cgo ...) needs nothing but this header:
 * plain pointers and sizes, no C++/torch types.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - matrices are row-major FP64 with a leading dimension `ld` in elements; ld must be even
 *     and base pointers 16-byte aligned (TMA requirement);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls only enqueue
 *     work unless stated otherwise;
 *   - return value: 0 = ok, <0 = bad argument, >=1000 CUDA error (1000 + cudaError_t),
 *     >=2000 matrix not positive definite, >=3000 tensor-map creation failed.
 *     munk_last_error_string() describes the last failure of the calling thread.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef MUNK_B200_H
#define MUNK_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct munk_workspace munk_workspace_t; /* per-stream scratch + statistics */
typedef struct munk_comm munk_comm_t;           /* one NCCL communicator (multi-GPU entry points) */
typedef struct munk_dist munk_dist_t;           /* one matrix being factored over the ranks of a communicator */

/* Operand storage for munk_dgemm: the contraction index k contiguous, or the free index. */
#define MUNK_OP_KC 0 /* X[i*ld + k] */
#define MUNK_OP_MC 1 /* X[k*ld + i] */

/* Structure flags for munk_dgemm (tile granularity 128; see munk_b200/csrc/dgemm.cuh). */
#define MUNK_GF_TILE_LOWER 1
#define MUNK_GF_KLO_M 2
#define MUNK_GF_KLO_N 4
#define MUNK_GF_KHI_M 8
#define MUNK_GF_KHI_N 16
#define MUNK_GF_SYM_MIRROR 32

const char* munk_last_error_string(void);
int munk_version(void);

/* Workspace: one per (device, stream) in use.  Not thread-safe; re-entrant across workspaces. */
int munk_workspace_create(int device, munk_workspace_t** ws);
int munk_workspace_destroy(munk_workspace_t* ws);
/* kernels enqueued and useful flops enqueued through `ws` since the last reset */
int munk_workspace_stats(munk_workspace_t* ws, long* launches, double* flops, int reset);

/* Bytes of leaf-inverse storage (`linv`) munk_potrf/munk_trtri need for order n. */
size_t munk_linv_bytes(int n);

/*
 * A = I + lam * L(G), dense n x ld, from the undirected edge list (u[e], v[e]), e < nnz, each
 * edge once, no self loops, indices in sorted-node order.
 * Replaces: munk/munk/munk.py:82-83 (nx.laplacian_matrix(...).todense(); np.eye + lam*L).
 * Bit-exact with numpy (diagonal fl(1 + fl(lam*deg)), off-diagonal -lam).
 */
int munk_assemble_reglap(munk_workspace_t* ws, int n, int nnz, const int* u, const int* v,
                         double lam, double* A, long ld, void* stream);
/* An edge with an endpoint outside [0, n) is skipped (never written) and remembered:
 * munk_assemble_info synchronises `stream` and returns 1 + the index of the first such edge in
 * *bad_edge_host (0 = every edge was in range); the device flag is cleared. */
int munk_assemble_info(munk_workspace_t* ws, int* bad_edge_host, void* stream);

/*
 * The same from explicit Laplacian entries, for graphs whose Laplacian is not 0/-1/degree
 * (edge weights, parallel edges of a MultiGraph - nx.laplacian_matrix sums them,
 * munk/munk/munk.py:82): lval[e] = L[u[e]][v[e]] (each unordered pair once), ldiag[i] = L[i][i],
 * both float64 exactly as networkx produced them.  A[u][v] = A[v][u] = fl(lam * lval),
 * A[i][i] = fl(1 + fl(lam * ldiag[i])): bit-exact with np.eye(n) + lam * L.
 */
int munk_assemble_reglap_entries(munk_workspace_t* ws, int n, int nnz, const int* u, const int* v,
                                 const double* lval, const double* ldiag, double lam, double* A,
                                 long ld, void* stream);

/*
 * In-place lower Cholesky A = L L^T (row-major; lower triangle read and written; the 128x128
 * diagonal blocks get zeros above the diagonal, other upper blocks are left untouched).
 * First half of the replacement for np.linalg.inv at munk/munk/munk.py:83.
 */
int munk_potrf(munk_workspace_t* ws, int n, double* A, long ld, double* linv, void* stream);

/* munk_potrf of two independent matrices (two workspaces on one device), e.g. the source and the
 * target Laplacian of munk/munk/munk.py:135,145: the block columns of both are interleaved on one
 * trailing-update stream, so that each factorisation's latency-bound diagonal-block chain hides
 * behind the other's tensor-core work.  Flags and leaf inverses as for two munk_potrf calls. */
int munk_potrf_pair(munk_workspace_t* wa, int na, double* A, long lda, double* linva, munk_workspace_t* wb,
                    int nb, double* B, long ldb, double* linvb, void* stream);

/* Synchronises `stream`; *info_host = 0 if every pivot so far was positive, else 1 + index of
 * the first bad column; the device flag is cleared. */
int munk_potrf_info(munk_workspace_t* ws, int* info_host, void* stream);

/* Same without synchronising: the flag is moved into the device int *info_dev (and cleared), so a
 * caller can keep enqueuing work and inspect all its flags once at the end of a step. */
int munk_potrf_info_async(munk_workspace_t* ws, int* info_dev, void* stream);

/*
 * In-place W = L^-1 of the factor left by munk_potrf (same linv).  C = W^T is an RKHS factor of
 * D = A^-1 (C C^T = D); replaces scipy.linalg.eigh + v.dot(diag(sqrt(e))) at
 * munk/munk/munk.py:87-88 (the factor differs from the reference's by an orthogonal matrix on
 * the right, which leaves D and the scores C * C2hat^T unchanged).
 */
int munk_trtri(munk_workspace_t* ws, int n, double* L, long ld, double* linv, void* stream);

/*
 * B (n x nrhs, in place) <- L^-1 B (transposed = 0) or L^-T B (transposed = 1) with the factor
 * left by munk_potrf.  Two calls give A^-1 B; with B = the landmark columns of the identity this
 * is target_D[:, target_idxs] (munk/munk/munk.py:110) without forming the full inverse.
 */
int munk_trsm_left(munk_workspace_t* ws, int n, const double* L, long ld, const double* linv,
                   int transposed, int nrhs, double* B, long ldb, const int* stair, void* stream);
/* `stair` (device int array, one entry per 128 columns of B, or NULL; forward solve only): the
 * first row at which that column tile of B is non-zero.  Work above it is skipped, so solving
 * for a set of columns of L^-1 (B = columns of the identity) costs only their non-zero part. */

/* B (m x s, in place) <- B L^-T: the panel solve of the blocked Cholesky (multi-GPU driver). */
int munk_trsm_right(munk_workspace_t* ws, int m, double* B, long ldb, const double* L, long ldl,
                    int s, const double* linv, void* stream);

/* munk_dgemm with staircase operands: stair_m[t] / stair_n[t] (device, per 128-wide tile, or NULL)
 * give the first global k at which the A-operand rows of m-tile t / the B-operand rows of n-tile
 * t are non-zero; row0 and k0 are the global indices of C's row 0 and of k = 0. */
int munk_dgemm_stair(munk_workspace_t* ws, int a_layout, int b_layout, int M, int N, int K,
                     double alpha, const double* A, long lda, const double* B, long ldb, double beta,
                     double* C, long ldc, int flags, const int* stair_m, const int* stair_n,
                     int row0, int k0, void* stream);

/* out (cols x rows) = in (rows x cols)^T, row-major. */
int munk_transpose(int rows, int cols, const double* in, long ldi, double* out, long ldo,
                   void* stream);

/* D = W^T W, full symmetric n x n (the value np.linalg.inv returns at munk/munk/munk.py:83). */
int munk_gram_lower(munk_workspace_t* ws, int n, const double* W, long ldw, double* D, long ldd,
                    void* stream);

/*
 * C(M x N) = alpha * A(M x K) * B(N x K)^T + beta * C, DMMA tensor cores, TMA-staged operands.
 * Replaces np.dot / @ at scripts/compute_embeddings.py:81,
 * experiments/homology-scores-permutation-test/permtest.py:55 and the products inside
 * munk/munk/munk.py:109-110.
 */
int munk_dgemm(munk_workspace_t* ws, int a_layout, int b_layout, int M, int N, int K, double alpha,
               const double* A, long lda, const double* B, long ldb, double beta, double* C,
               long ldc, int flags, void* stream);

/* Q[i][l] = W[i][idx[l]] for i >= idx[l], else 0 (n x ldq, columns l >= k are zeroed):
 * the rows C[idx,:] of C = W^T, transposed (munk/munk/munk.py:109 source_C[source_idxs,:]). */
int munk_gather_cols_lower(int n, const double* W, long ldw, int k, const int* idx, double* Q,
                           long ldq, void* stream);

/* C[i][j] = W[j][i] for j >= i, else 0: materialises the upper-triangular factor C = W^T. */
int munk_transpose_lower(int n, const double* W, long ldw, double* C, long ldc, void* stream);

/*
 * Device -> host copy of the upper-triangular factor C (n x n on the device, row-major) into the
 * host array `host` (row-major, ldh elements per row): row block b of `row_block` rows copies the
 * columns from its first row index onwards, i.e. every non-zero of C and at most row_block - 1
 * structural zeros per row; the strictly lower part of `host` is NOT written (the caller zeroes it
 * once).  Halves the bytes moved for the array the reference pickles as the source embedding
 * (scripts/compute_embeddings.py:72, munk/munk/munk.py:178-180).  Asynchronous on `stream`
 * when `host` is page-locked.
 */
int munk_download_upper(int n, const double* C, long ldc, double* host, long ldh, int row_block,
                        void* stream);

/* Device -> host copy of a rows x cols block (both row-major, leading dimensions in elements):
 * one strided DMA, asynchronous on `stream` when `host` is page-locked.  The multi-GPU result
 * path uses it to move only the non-zero part of each rank's rows of the triangular factor
 * (the array saved at scripts/compute_embeddings.py:72). */
int munk_download_block(int rows, int cols, const double* C, long ldc, double* host, long ldh,
                        void* stream);

/*
 * Device -> PAGEABLE host memory (a fresh numpy array the caller owns: the return values of
 * munk.embed_networks / the score matrix of scripts/compute_embeddings.py:81).  Blocking.  `threads`
 * worker threads (1..16), each with a pinned staging buffer and a stream of its own, move row chunks:
 * DMA into the pinned buffer, then memcpy (and first touch) into `host`.  Work enqueued on `stream`
 * before the call is waited for.  upper != 0 (square matrices): the source is upper triangular;
 * columns left of each chunk's first row are zero-filled on the host instead of being moved.
 */
int munk_download_pageable(int device, long rows, long cols, const double* src, long lds, double* host,
                           long ldh, int upper, int threads, void* stream);

/*
 * 128-bit content digest of a HOST buffer (no device work): out[0] = position-dependent polynomial
 * hash of the 64-bit words (any one-word change always changes it), out[1] = sum of the words.
 * `threads` host threads read the buffer at memory-bus speed; the result does not depend on the
 * thread count.  munk.scores() uses it to decide whether the arrays it is handed are, byte for byte,
 * the ones munk.embed_networks returned (so that np.dot(source_X, target_X.T),
 * scripts/compute_embeddings.py:72-81, can be served from the factors still on the device).
 */
int munk_host_digest(const void* data, size_t bytes, int threads, unsigned long long* out);

/* out[g][:] = mean of in[members[gptr[g] .. gptr[g+1])][:]; used to fold duplicated source
 * landmarks (np.linalg.pinv's minimum-norm least-squares answer, munk/munk/munk.py:109). */
int munk_group_mean_rows(int ngroups, const int* gptr, const int* members, int ncols,
                         const double* in, long ldi, double* out, long ldo, void* stream);

/* out[l] = S[r[l]][c[l]] (munk/munk/munk.py:42 scores[source_landmark_idxs, target_landmark_idxs]) */
int munk_gather_pairs(int npairs, const int* r, const int* c, const double* S, long lds,
                      double* out, void* stream);

/* out = tril(in): the lower factor with exact zeros above the diagonal (n x n). */
int munk_copy_lower(int n, const double* in, long ldi, double* out, long ldo, void* stream);

/*
 * separate_scores (munk/munk/munk.py:20-68) as one streaming pass over S (n1 x n2).  The host
 * derives from the integer index lists: rowL/colL (landmark row / column flags), colrank[j] =
 * number of landmark columns before j, the landmark-pair cells and the H_H cells (homolog
 * pairs minus landmark pairs) as CSR over rows with ascending columns, and for each row the
 * offset of its first element in each output.  Outputs are in the reference's row-major mask
 * order.  (L_L_diag and H_H themselves are munk_gather_pairs over the index lists.)
 */
int munk_separate_scores(int n1, int n2, const double* S, long lds, const unsigned char* rowL,
                         const unsigned char* colL, const int* colrank, const int* p_ptr,
                         const int* p_col, const int* h_ptr, const int* h_col,
                         const long* off_lloff, const long* off_lnonl, const long* off_other,
                         double* out_lloff, double* out_lnonl, double* out_other, void* stream);

/*
 * out5 = {mean(H_H), mean(other), mean(H_H) - mean(other), |H_H|, |other|} without materialising
 * the selections (permtest.py:58-66 dissim_diff).  h_row/h_col list the nh H_H cells;
 * n_other = |other| (host integer arithmetic); rowsum_scratch holds n1 doubles.
 */
int munk_separate_scores_means(int n1, int n2, const double* S, long lds,
                               const unsigned char* rowL, const unsigned char* colL, int nh,
                               const int* h_row, const int* h_col, double n_other,
                               double* rowsum_scratch, double* out5, void* stream);

/*
 * out[p] = X[r[p], 0:k] . Z[0:k, c[p]]: single scores S[r,c] of the factored form S = X Z
 * (X = D1[:,Ls] n1 x k, Z = D1[Ls,Ls]^-1 D2[Lt,:] k x n2) without forming S.  Used by the opt-in
 * rank-k permutation statistic, never by the dense path.
 */
int munk_pair_dots(int npairs, const int* r, const int* c, int k, const double* X, long ldx,
                   const double* Z, long ldz, double* out, void* stream);

/*
 * Rank-revealing pseudo-inverse solve on the landmark rows themselves, numpy semantics
 * (np.linalg.pinv(source_C[source_idxs,:]) at munk/munk/munk.py:109: singular values
 * <= rcond * sigma_max dropped, rcond = 1e-15).  R (k x m, row-major) holds the landmark rows M on
 * entry; one-sided Jacobi rotations from the left make its rows orthogonal, R <- J M.  On return
 *     Z (k x n) = diag(1 / sigma_i^2 if sigma_i > rcond * sigma_max else 0) (J B),
 * so that pinv(M) B = R^T Z (one munk_dgemm with R as an MC-layout operand).  *rank_host receives
 * the retained rank, sigma_host (k doubles, may be NULL) the singular values.  Synchronises
 * `stream` once per sweep.  `work` holds munk_pinv_rows_work_bytes(k) bytes.  This is the robust
 * path; the hot path solves with the Cholesky factor of M M^T when a device-side condition bound
 * allows it (munk_b200/device.py).
 */
int munk_pinv_rows_solve(munk_workspace_t* ws, int k, int m, double* R, long ldr, int n,
                         const double* B, long ldb, double* Z, long ldz, double* work, double rcond,
                         int* rank_host, double* sigma_host, void* stream);
size_t munk_pinv_rows_work_bytes(int k);

/*
 * ---- multi-GPU (one process per GPU, NCCL over NVLink; munk_b200/csrc/dist.cu) -------------------
 * Replaces np.linalg.inv (munk/munk/munk.py:83) and the eigh factor (munk.py:87-88) for a matrix
 * whose block columns (width munk_dist_block() = 512) are dealt to the ranks in serpentine rounds
 * (rank of block j: r = j mod P in even rounds, P-1-r in odd rounds).  NCCL is taken from the
 * process at run time (dlopen of libnccl.so.2); status codes >= 4000 are 4000 + ncclResult_t.
 *
 * Communicator: rank 0 calls munk_comm_unique_id and ships the 128 bytes to the other ranks by
 * any means (torch.distributed in munk_b200/dist.py); every rank then calls munk_comm_create.
 * world = 1 needs no id and no NCCL.
 */
int munk_dist_block(void);
int munk_comm_unique_id(void* id128);
/* max_ctas > 0 caps the SMs one collective of this communicator may occupy (ncclConfig_t.maxCTAs);
 * 0 leaves NCCL's default. */
int munk_comm_create(int device, int rank, int world, const void* id128, int max_ctas, munk_comm_t** comm);
int munk_comm_destroy(munk_comm_t* comm);

/*
 * Panel-major factor storage: block column c is one contiguous buffer [leaf inverses | rows
 * s_c..n, leading dimension round_up(w_c, 16)], the owner's working storage and the other ranks'
 * broadcast buffer at once.  `storage` holds munk_dist_storage_elems(n) doubles (about n^2/2),
 * 16-byte aligned, owned by the caller, and must outlive the handle.  comm = NULL: one rank.
 */
size_t munk_dist_storage_elems(int n);
/* `bulk` (may be NULL = use `comm`): a second communicator of the same ranks for the large part of
 * every panel, so that the next panel's small critical piece never queues behind 70 MB. */
int munk_dist_create(munk_workspace_t* ws, munk_comm_t* comm, munk_comm_t* bulk, int n, double* storage,
                     munk_dist_t** dist);
int munk_dist_destroy(munk_dist_t* dist);
int munk_dist_blocks(munk_dist_t* dist);

/* I + lam L into the block columns this rank owns (lval/ldiag NULL: unweighted, as
 * munk_assemble_reglap; otherwise as munk_assemble_reglap_entries).  Bit-exact with numpy. */
int munk_dist_assemble(munk_dist_t* dist, int nnz, const int* u, const int* v, const double* lval,
                       const double* ldiag, double lam, void* stream);

/* Right-hand side of the fused forward substitution: X (n x ncols, row-major, ldx) <- L^-1 X,
 * applied panel by panel while the factorisation runs.  `stair` (device, one int per 128 columns,
 * or NULL): first non-zero row of each column tile (identity columns: only the non-zero part of
 * every column is computed).  X = NULL: factorisation only. */
int munk_dist_set_rhs(munk_dist_t* dist, double* X, long ldx, int ncols, const int* stair);

/* The factorisation, one block column per step so that a caller can interleave several
 * matrices from one host thread (every rank must make the same sequence of calls): begin forks
 * the library's streams from `stream`, step enqueues the next block column (*remaining = block
 * columns still to go), end enqueues whatever is left and joins everything back into `stream`.
 * The non-SPD flag is the workspace's (munk_potrf_info / munk_potrf_info_async). */
int munk_dist_potrf_begin(munk_dist_t* dist, void* stream);
int munk_dist_potrf_step(munk_dist_t* dist, int* remaining);
int munk_dist_potrf_end(munk_dist_t* dist);

/* X (n x ncols, in place) <- L^-T X with the finished factor (every rank holds all of it):
 * with munk_dist_set_rhs this gives A^-1 X, i.e. target_D[:, target_idxs] (munk/munk/munk.py:110). */
int munk_dist_solve_backward(munk_dist_t* dist, double* X, long ldx, int ncols, void* stream);

/* L (n x ld, row-major, lower triangle incl. zeros above the diagonal inside the diagonal blocks)
 * <- the panel-major factor (tests, tools). */
int munk_dist_unpack(munk_dist_t* dist, double* L, long ld, void* stream);

/* Micro-benchmarks used by bench.py / tools to measure the FP64 ceilings of the box.
 * Each runs on `device`, blocks, and returns TFLOP/s in *tflops. */
int munk_microbench_dmma(int device, int iters, double* tflops);
int munk_microbench_dfma(int device, int iters, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* MUNK_B200_H */
