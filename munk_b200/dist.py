"""Multi-GPU MUNK pipeline: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

Process grid.  The factorisations use a 1 x P block-column-cyclic layout (block width 512) dealt
in serpentine rounds (ranks 0..P-1, then P-1..0, ...; ColumnCyclic.owner).  With NVSwitch every rank sees every peer at full bandwidth, so
the per-rank panel traffic of a P_r x P_c grid (the whole panel, n^2/2 doubles in total) is the
same for any grid shape; 1 x P keeps each panel's factorisation on one GPU (no intra-panel
exchange) and leaves every rank with the complete factor L, which is what the later stages want:

  potrf   right-looking; the owner of panel j+1 applies panel j to it first (look-ahead), factors
          it (diagonal block + panel solve, munk_potrf / munk_trsm_right) and broadcasts it
          (ncclBroadcast, asynchronous) while all ranks apply panel j to the block columns they
          own (DMMA syrk).  At the end every rank holds all of L and the leaf inverses.
  W, Q1   columns of W = L^-1 are independent given L: each rank solves L X = E for the identity
          columns of its own blocks and (redundantly, k columns) for the landmark columns, in ONE
          forward solve whose right-hand side is a staircase (munk_trsm_left with `stair`), so
          only the non-zero part of every column is computed: n^3/(3P) flops per rank, no exchange.
  C1, S   rank r owns the rows of C1 = W^T and of S = C1 C2hat^T that belong to its blocks
          (row-block sharding): S_r = X_r^T P with the same staircase on the k-range.
          P = C2hat^T is small work (2 n1 n2 k flops) and is formed on every rank from the
          replicated Q1 and D2[Lt,:], so the score product needs no all-gather.
  stats   separate_scores means: per-rank partial sums + one 4-scalar all-reduce.
  perms   permutation test: munk_b200/permtest.py (round-robin, one all-gather of 3 doubles each).

The driver is written against a small `ops` object so that the schedule, the ownership map and the
packing can be exercised on CPU tensors with gloo (tests/test_dist_cpu.py supplies a torch-CPU
`ops`); the product always uses CudaOps (no CPU fallback: CudaOps needs a Device).
"""
import numpy as np
import torch
import torch.distributed as dist

from .device import KC, MC, TILE_LOWER, pad16, landmark_groups

NB = 512


class ColumnCyclic:
    """Ownership map of a 1 x P block-column-cyclic layout."""

    def __init__(self, n, world, nb=NB):
        self.n, self.world, self.nb = int(n), int(world), int(nb)
        self.nblk = (self.n + self.nb - 1) // self.nb

    def owner(self, j):
        """Serpentine rounds (0 .. P-1, P-1 .. 0, 0 .. P-1, ...): the work attached to a block
        column (trailing updates, the staircase solve, the k-range of its score rows) shrinks
        linearly with j, and plain round-robin would give rank 0 the heaviest block of every
        round (measured at 8 GPUs: 28 ms vs 20 ms in the score product)."""
        r, rnd = j % self.world, j // self.world
        return r if rnd % 2 == 0 else self.world - 1 - r

    def cols(self, j):
        return j * self.nb, min(self.n, (j + 1) * self.nb)

    def blocks_of(self, rank):
        return [j for j in range(self.nblk) if self.owner(j) == rank]

    def rows_of(self, rank):
        """Global row indices (of C1 / S) owned by `rank`, ascending."""
        out = []
        for j in self.blocks_of(rank):
            a, b = self.cols(j)
            out.extend(range(a, b))
        return out


class CudaOps:
    """The four operations of the distributed Cholesky on a munk_b200 Device."""

    def __init__(self, dev):
        self.dev = dev
        self.device = dev.torch_device

    def syrk_lower(self, C, X, Y, m, nc, k):
        """C[:m, :nc] -= X[:m, :k] Y[:nc, :k]^T (at least the tiles touching the lower triangle)."""
        self.dev.gemm(KC, KC, m, nc, k, -1.0, X, Y, 1.0, C, TILE_LOWER)

    def factor_panel(self, A, n, c0, w, linv):
        """In place: Cholesky of the w x w diagonal block at (c0, c0) and the panel solve below."""
        diag = A[c0:, c0:]
        lv = linv[(c0 // 128) * 128 * 128:]
        self.dev.potrf(diag, w, linv=lv, check_info=False)
        below = n - (c0 + w)
        if below > 0:
            self.dev.trsm_right(A[c0 + w:, c0:], below, diag, w, lv)

    def info(self):
        return self.dev.potrf_info()


def _panel_elems(n, c0, w):
    ldp = pad16(w)
    return (n - c0) * ldp, ldp, ((w + 127) // 128) * 128 * 128


def dist_potrf_steps(ops, A, n, linv, rank, world, group=None, nb=NB, stream=None):
    """Generator form of dist_potrf: enqueues one panel step per next() (on `stream` if given),
    so that two independent factorisations can be interleaved from one host thread and overlap
    on the device (each is latency-bound in its panel chain)."""
    import contextlib
    ctx = (lambda: torch.cuda.stream(stream)) if stream is not None else contextlib.nullcontext
    cc = ColumnCyclic(n, world, nb)
    max_elems = n * pad16(min(nb, n)) + ((min(nb, n) + 127) // 128) * 128 * 128
    with ctx():
        bufs = [torch.empty(max_elems, dtype=torch.float64, device=A.device) for _ in range(2)]

    def pack(j, buf):
        c0, c1 = cc.cols(j)
        w = c1 - c0
        ne, ldp, nl = _panel_elems(n, c0, w)
        view = buf[:ne].view(n - c0, ldp)
        view[:, :w].copy_(A[c0:, c0:c1])
        if ldp != w:
            view[:, w:].zero_()
        lo = (c0 // 128) * 128 * 128
        buf[ne:ne + nl].copy_(linv[lo:lo + nl])

    def unpack(j, buf):
        c0, c1 = cc.cols(j)
        w = c1 - c0
        ne, ldp, nl = _panel_elems(n, c0, w)
        A[c0:, c0:c1].copy_(buf[:ne].view(n - c0, ldp)[:, :w])
        lo = (c0 // 128) * 128 * 128
        linv[lo:lo + nl].copy_(buf[ne:ne + nl])

    def bcast(j, buf, async_op):
        c0, c1 = cc.cols(j)
        ne, _, nl = _panel_elems(n, c0, c1 - c0)
        if world == 1:
            return None
        return dist.broadcast(buf[:ne + nl], src=_global_rank(group, cc.owner(j)), group=group,
                              async_op=async_op)

    cur, nxt = bufs
    with ctx():
        if rank == cc.owner(0):
            ops.factor_panel(A, n, 0, cc.cols(0)[1], linv)
            pack(0, cur)
        bcast(0, cur, async_op=False)
        if rank != cc.owner(0):
            unpack(0, cur)
    yield

    for j in range(cc.nblk):
        with ctx():
            c0, c1 = cc.cols(j)
            w = c1 - c0
            ne, ldp, _ = _panel_elems(n, c0, w)
            pan = cur[:ne].view(n - c0, ldp)                     # panel j, rows c0.., on every rank
            work = None
            if j + 1 < cc.nblk:
                r1, r1e = cc.cols(j + 1)
                w1 = r1e - r1
                if rank == cc.owner(j + 1):                      # look-ahead: update, factor, send
                    ops.syrk_lower(A[r1:, r1:], pan[r1 - c0:], pan[r1 - c0:], n - r1, w1, w)
                    ops.factor_panel(A, n, r1, w1, linv)
                    pack(j + 1, nxt)
                work = bcast(j + 1, nxt, async_op=True)
            for c in cc.blocks_of(rank):                         # my other trailing block columns
                if c <= j + 1:
                    continue
                cs, ce = cc.cols(c)
                ops.syrk_lower(A[cs:, cs:], pan[cs - c0:], pan[cs - c0:], n - cs, ce - cs, w)
            if j + 1 < cc.nblk:
                if work is not None:
                    work.wait()
                if rank != cc.owner(j + 1):
                    unpack(j + 1, nxt)
            cur, nxt = nxt, cur
        yield


def dist_potrf(ops, A, n, linv, rank, world, group=None, nb=NB, check=True):
    """Distributed in-place Cholesky of the (replicated-storage) matrix A; on return every rank
    holds the complete factor in the lower triangle of A and all leaf inverses in `linv`.
    Returns the non-SPD flag (max over ranks; 0 = ok), or None with check=False."""
    for _ in dist_potrf_steps(ops, A, n, linv, rank, world, group, nb):
        pass
    if not check:
        return None          # the caller collects the flag later (keeps the host running ahead)
    return dist_potrf_flag(ops, A.device, world, group)


def dist_potrf_flag(ops, device, world, group=None):
    """Synchronises, then max over ranks of the non-SPD flag of `ops` (0 = ok)."""
    flag = torch.tensor([ops.info()], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    return int(flag.item())


_SIDE = {}


def _side_resources(dev, world, group):
    """(second Device/workspace, second stream, second communicator) of this process for work
    that runs concurrently with the main chain; created once.  Collective on first call when
    world > 1 (dist.new_group)."""
    key = (dev.index, id(group))
    res = _SIDE.get(key)
    if res is None:
        from .device import Device
        group2 = None
        if world > 1:
            ranks = list(range(world)) if group is None else dist.get_process_group_ranks(group)
            group2 = dist.new_group(ranks=ranks)
        res = _SIDE[key] = (Device(dev.index), torch.cuda.Stream(device=dev.index), group2)
    return res


def _global_rank(group, group_rank):
    if group is None:
        return group_rank
    return dist.get_global_rank(group, group_rank)


# ---------------------------------------------------------------------------------------------
def staircase_rhs(n, owned_cols, landmark_cols, device):
    """One-hot right-hand side [E_owned | E_landmarks] (n x ncols, zero padded to 128-column
    tiles) and its staircase: the first non-zero row of each 128-column tile.

    owned_cols: ascending global column indices owned by this rank (whole blocks);
    landmark_cols: ascending distinct landmark indices.  Returns (X, stair, n_owned_padded)."""
    own = np.asarray(owned_cols, dtype=np.int64)
    lm = np.asarray(landmark_cols, dtype=np.int64)
    n_own_pad = (len(own) + 127) // 128 * 128
    n_lm_pad = (len(lm) + 127) // 128 * 128
    ncols = n_own_pad + n_lm_pad
    X = torch.zeros((n, max(ncols, 2)), dtype=torch.float64, device=device)
    col_of = np.concatenate([np.arange(len(own)), n_own_pad + np.arange(len(lm))])
    row_of = np.concatenate([own, lm])
    if len(row_of):
        X[torch.as_tensor(row_of, device=device), torch.as_tensor(col_of, device=device)] = 1.0
    stair = np.full(ncols // 128, n, dtype=np.int32)           # empty tile: nothing to do
    for t in range(ncols // 128):
        rows = row_of[(col_of >= 128 * t) & (col_of < 128 * (t + 1))]
        if len(rows):
            stair[t] = int(rows.min())
    return X, torch.from_numpy(stair).to(device), n_own_pad


class ShardSink:
    """Pinned host buffers for one rank's share of the three results and a copy stream: its rows
    of C1 and S (ColumnCyclic.rows_of(rank)) and its rows of C2hat, i.e. the column block
    [col0, col1) of the replicated P = C2hat^T, so that every result leaves the box sharded and
    no rank moves more than (n1 + 2 n2) n1 / P doubles.  Reusable across calls."""

    def __init__(self, dev, n1, n2, rank, world):
        self.dev, self.n1, self.n2 = dev, n1, n2
        self.rows = ColumnCyclic(n1, world).rows_of(rank)
        per = (n2 + world - 1) // world
        self.col0, self.col1 = min(n2, rank * per), min(n2, (rank + 1) * per)
        m = len(self.rows)
        opts = dict(dtype=torch.float64, pin_memory=True)
        self.C1 = torch.zeros((m, n1), **opts)       # C1 is upper triangular: the part left of each
        self.blocks, l0 = [], 0                      # block's first column stays zero, never copied
        cc = ColumnCyclic(n1, world)
        for j in cc.blocks_of(rank):
            a, b = cc.cols(j)
            self.blocks.append((l0, l0 + b - a, a))
            l0 += b - a
        self.S = torch.empty((m, n2), **opts)
        self.C2hat = torch.empty((self.col1 - self.col0, n1), **opts)
        self.stream = torch.cuda.Stream(device=dev.index)
        self.done = {}                     # name -> event recorded after that copy (timelines)

    def _copy_out(self, name, host, src):
        """host <- src (contiguous CUDA tensor produced on the current stream), on the copy stream.
        Only the DMA runs there: anything that needs SMs (the transposes below) is enqueued on the
        compute stream *before* the next GEMM, because a kernel on a second stream does not get
        an SM while a tensor-core launch with long tiles occupies them all (measured: the copies
        were delayed to the end of the score product)."""
        self.stream.wait_stream(torch.cuda.current_stream(self.dev.index))
        with torch.cuda.stream(self.stream):
            host.copy_(src, non_blocking=True)
            self.done[name] = torch.cuda.Event(enable_timing=True)
            self.done[name].record()
        src.record_stream(self.stream)

    def _transposed(self, X, rows, cols):
        T = torch.empty((cols, rows), dtype=torch.float64, device=self.dev.torch_device)
        return self.dev.transpose(X, rows, cols, out=T)

    def push_C1(self, X, m_own):
        """Rows of C1 = the first m_own columns of the solve result X (n1 x ...), transposed."""
        if not m_own:
            return
        T = self._transposed(X, self.n1, m_own)
        self.stream.wait_stream(torch.cuda.current_stream(self.dev.index))
        with torch.cuda.stream(self.stream):
            for l0, l1, c0 in self.blocks:       # rows of block column c0.. are zero left of c0
                self.dev.download_block(T[l0:l1, c0:], self.C1[l0:l1, c0:])
            self.done["C1"] = torch.cuda.Event(enable_timing=True)
            self.done["C1"].record()
        T.record_stream(self.stream)

    def push_P(self, P):
        if self.col1 > self.col0:
            self._copy_out("C2hat", self.C2hat, self._transposed(P[:, self.col0:], self.n1, self.col1 - self.col0))

    def push_S(self, S, r0, r1):
        """Local rows [r0, r1) of the rank's score rows."""
        if r1 > r0:
            src = S[r0:r1, :self.n2]
            self._copy_out("S", self.S[r0:r1], src if src.is_contiguous() else src.contiguous())

    def d2h_bytes(self):
        c1 = sum((l1 - l0) * (self.n1 - c0) for l0, l1, c0 in self.blocks)
        return (c1 + self.S.numel() + self.C2hat.numel()) * 8

    def wait(self):
        self.stream.synchronize()


def dist_embed_scores(dev, n1, edges1, n2, edges2, landmark_idxs, rank, world, group=None,
                      src_lam=0.05, tgt_lam=0.05, homolog_idxs=None, marks=None, sink=None):
    """The whole hot path on `world` GPUs.  Returns a dict with the rank's shard
    (`sink`: optional ShardSink; each result is copied to pinned host memory as soon as it is
    enqueued, overlapping the remaining stages):
       rows   global row indices of C1 / S owned by this rank
       C1     (len(rows), n1) rows of the source embedding       (CUDA tensor)
       P      (n1, n2) C2hat^T, replicated                        (CUDA tensor, padded ld)
       S      (len(rows), n2) rows of the score matrix            (CUDA tensor)
       stats  (hom_mean, other_mean, diff) over the full matrix if homolog_idxs is given."""
    ops = CudaOps(dev)

    def mark(name):
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    mark("start")
    s_idx = [int(a) for a, _ in landmark_idxs]
    t_idx = [int(b) for _, b in landmark_idxs]
    # distinct source landmarks in ascending order; groups of target rows to average per landmark
    uniq_first, gptr_f, members_f = landmark_groups(s_idx)
    order = np.argsort(np.asarray(uniq_first, dtype=np.int64), kind="stable")
    uniq = [uniq_first[g] for g in order]
    gptr, members = [0], []
    for g in order:
        members.extend(members_f[gptr_f[g]:gptr_f[g + 1]])
        gptr.append(len(members))

    # 1. factors, replicated on every rank.  The target chain (assembly, Cholesky, the k-column
    #    solve for D2[Lt,:]) is independent of the source chain and, like it, latency-bound in its
    #    panel steps, so it runs concurrently: own stream, own workspace, own communicator.
    main = torch.cuda.current_stream(dev.index)
    dev2, side, group2 = _side_resources(dev, world, group)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        A2 = dev2.assemble(n2, edges2[0], edges2[1], tgt_lam)
        linv2 = dev2.new_linv(n2)
        ops2 = CudaOps(dev2)
    A1 = dev.assemble(n1, edges1[0], edges1[1], src_lam)
    linv1 = dev.new_linv(n1)
    # interleave the two panel loops step by step (one host thread feeds both streams)
    src_steps = dist_potrf_steps(ops, A1, n1, linv1, rank, world, group)
    tgt_steps = dist_potrf_steps(ops2, A2, n2, linv2, rank, world, group2, stream=side)
    src_live = tgt_live = True
    while src_live or tgt_live:
        if src_live:
            src_live = next(src_steps, "done") != "done"
        if tgt_live:
            tgt_live = next(tgt_steps, "done") != "done"
    with torch.cuda.stream(side):
        B = dev2.target_landmark_rows_from_factor(A2, linv2, n2, t_idx)
    mark("potrf_src")
    # 2. my columns of W1 and the landmark columns Q1, one staircase solve
    cc = ColumnCyclic(n1, world)
    rows = cc.rows_of(rank)
    X, stair, n_own_pad = staircase_rhs(n1, rows, uniq, dev.torch_device)
    ncols = X.shape[1]
    if ncols >= 128:
        dev.trsm_left(A1, n1, linv1, X, ncols, transposed=False, stair=stair)
    kq = len(uniq)
    Q1 = X[:, n_own_pad:]                                        # n1 x pad128(k'), view

    mark("solve_src")
    # 3. projection (replicated): join the target chain, check both factorisations
    main.wait_stream(side)
    for t in (A2, linv2, B):
        t.record_stream(main)
    # the factorisation flags are parked in device ints (no host sync here: the host keeps
    # enqueuing) and inspected once at the end of the step
    flags = torch.cat([dev.potrf_info_async(), dev2.potrf_info_async()]).to(torch.int64)
    P = dev.project(Q1, n1, B, n2, (gptr, members))
    if sink is not None:
        sink.push_C1(X, len(rows))
        sink.push_P(P)

    mark("embed")
    # 4. my rows of S = C1 C2hat^T:  S_r = X_own^T P, k-range clipped by the same staircase
    m_own = len(rows)
    S = dev.empty(max(m_own, 1), n2)
    if m_own and sink is None:
        dev.gemm_stair(MC, MC, m_own, n2, n1, 1.0, X, P, 0.0, S, 0, stair_m=stair, k0=0)
    elif m_own:
        # row blocks, bottom-up, each copied out while the next one is computed
        from .pipeline import streamed_row_blocks
        for t0, t1 in streamed_row_blocks((m_own + 127) // 128, (n2 + 127) // 128):
            r0, r1 = t0 * 128, min(m_own, t1 * 128)
            dev.gemm_stair(MC, MC, r1 - r0, n2, n1, 1.0, X[:, r0:], P, 0.0, S[r0:], 0,
                           stair_m=stair[t0:t1], k0=0)
            sink.push_S(S, r0, r1)
    C1 = X[:, :m_own].t() if m_own else X[:, :0].t()            # (m_own, n1) view of the solve

    mark("score")
    bad = flags.max().reshape(1)
    if world > 1:
        dist.all_reduce(bad, op=dist.ReduceOp.MAX, group=group)
    bad = int(bad.item())
    if bad:
        from .device import NotPositiveDefinite
        raise NotPositiveDefinite("matrix is not positive definite (pivot %d)" % (bad - 1))
    mark("flag_check")
    out = dict(rows=rows, C1=C1, P=P, S=S[:m_own], n1=n1, n2=n2)
    if homolog_idxs is not None:
        out["stats"] = dist_score_means(dev, S, rows, n1, n2, landmark_idxs, homolog_idxs, world, group)
    return out


def gather_rows(local, n_rows, rank, world, group=None, dst=0):
    """Row-sharded (len(rows_of(rank)) x ncols) CUDA tensors -> the full (n_rows x ncols) matrix in
    global row order on rank `dst` (None elsewhere).  One NCCL gather of zero-padded shards."""
    cc = ColumnCyclic(n_rows, world)
    rows_by_rank = [cc.rows_of(r) for r in range(world)]
    ncols = int(local.shape[1])
    if world == 1:
        return local[:, :ncols].contiguous()
    mx = max(len(r) for r in rows_by_rank)
    pad = torch.zeros((mx, ncols), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=_global_rank(group, dst), group=group)
    if rank != dst:
        return None
    full = torch.empty((n_rows, ncols), dtype=local.dtype, device=local.device)
    for r, rows in enumerate(rows_by_rank):
        if rows:
            full[torch.as_tensor(rows, device=local.device)] = parts[r][:len(rows)]
    return full


def embed_and_score_distributed(dev, inputs, rank, world, src_lam=0.05, tgt_lam=0.05, group=None):
    """pipeline.embed_and_score on `world` GPUs (scripts/compute_embeddings.py under torchrun):
    every rank runs dist_embed_scores on the same index inputs, the row shards of C1 and S are
    gathered on rank 0 over NVLink and rank 0 returns the same dict as the single-GPU call
    (host arrays, reference layouts); the other ranks return None."""
    import time
    inp = inputs
    n1, n2 = len(inp["source_nodes"]), len(inp["target_nodes"])
    marks = []
    t0 = time.time()
    out = dist_embed_scores(dev, n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"], rank, world,
                            group=group, src_lam=src_lam, tgt_lam=tgt_lam, marks=marks)
    torch.cuda.current_stream(dev.index).synchronize()
    phase = {name: a.elapsed_time(b) * 1e-3 for (_, a), (name, b) in zip(marks[:-1], marks[1:])}
    C1 = gather_rows(out["C1"].contiguous(), n1, rank, world, group)
    S = gather_rows(out["S"][:, :n2].contiguous(), n1, rank, world, group)
    if rank != 0:
        return None
    source_X = dev.download(C1, n1, n1)
    target_X = dev.download(out["P"], n1, n2).T        # (n2, n1) transposed view, munk.py:109-110
    scores = dev.download(S, n1, n2)
    # the reference's four timer keys (munk.py:161-164) mapped onto the distributed phases: both
    # Cholesky chains run interleaved in the first, the W-column solves are the factorisation
    runtimes = dict(source_regularized_laplacian_runtime=phase.get("potrf_src", 0.0),
                    source_rkhs_factorization_runtime=phase.get("solve_src", 0.0),
                    target_regularized_laplacian_runtime=0.0,
                    embed_runtime=phase.get("embed", 0.0))
    return dict(source_X=source_X, target_X=target_X, scores=scores, source_nodes=inp["source_nodes"],
                target_nodes=inp["target_nodes"], landmarks=inp["landmark_homs"], runtimes=runtimes,
                score_runtime=phase.get("score", 0.0), wall=time.time() - t0)


def dist_score_means(dev, S_local, rows, n1, n2, landmark_idxs, homolog_idxs, world, group=None):
    """mean(H_H), mean(other), difference over the row-sharded score matrix: local partial sums
    (munk_separate_scores_means on the local rows) + one all-reduce of 4 doubles."""
    sums = torch.zeros(4, dtype=torch.float64, device=dev.torch_device)
    if rows:
        plan = _local_plan(len(rows), n2, rows, landmark_idxs, homolog_idxs)
        out = plan.means(dev, S_local)
        torch.cuda.current_stream(dev.index).synchronize()
        o = out.cpu().numpy()
        hh_sum = o[0] * o[3] if o[3] > 0 else 0.0
        ot_sum = o[1] * o[4] if o[4] > 0 else 0.0
        sums = torch.tensor([hh_sum, o[3], ot_sum, o[4]], dtype=torch.float64, device=dev.torch_device)
    if world > 1:
        dist.all_reduce(sums, group=group)
    s = sums.cpu().numpy()
    hom_mean, other_mean = s[0] / s[1], s[2] / s[3]
    return hom_mean, other_mean, hom_mean - other_mean


def _local_plan(m_local, n2, rows, landmark_idxs, homolog_idxs):
    """SeparatePlan of the local row block: row flags / cell lists restricted to owned rows,
    column flags from every landmark."""
    from .scores_plan import SeparatePlan
    local_of = {g: i for i, g in enumerate(rows)}
    plan = SeparatePlan.__new__(SeparatePlan)
    ls = np.array([a for a, _ in landmark_idxs], dtype=np.int64)
    lt = np.array([b for _, b in landmark_idxs], dtype=np.int64)
    hs = np.array([a for a, _ in homolog_idxs], dtype=np.int64)
    ht = np.array([b for _, b in homolog_idxs], dtype=np.int64)
    rowL = np.zeros(m_local, dtype=np.uint8)
    for a in ls:
        if int(a) in local_of:
            rowL[local_of[int(a)]] = 1
    colL = np.zeros(n2, dtype=np.uint8)
    colL[lt] = 1
    pair_keys = set(zip(ls.tolist(), lt.tolist()))
    cells = sorted({(local_of[a], b) for a, b in zip(hs.tolist(), ht.tolist())
                    if a in local_of and (a, b) not in pair_keys})
    h_row = np.array([c[0] for c in cells], dtype=np.int32)
    h_col = np.array([c[1] for c in cells], dtype=np.int32)
    n_nonl_rows = int(m_local - rowL.sum())
    n_nonl_cols = int(n2 - colL.sum())
    in_other = sum(1 for r, c in cells if not rowL[r] and not colL[c])
    plan.n1, plan.n2 = m_local, n2
    plan.rowL, plan.colL = rowL, colL
    plan.h_row, plan.h_col = h_row, h_col
    plan.n_hh = len(cells)
    plan.n_other = n_nonl_rows * n_nonl_cols - in_other
    return plan
