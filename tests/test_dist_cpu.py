"""Multi-rank host logic on CPU: world_size-2 gloo runs of the permutation sharding / gather
(munk_b200/permtest.py) and of the block-cyclic index maps (munk_b200/dist.py)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_ranks(body, world=2, timeout=120):
    """Runs `body` (python source; sees rank, world, dist) on `world` gloo ranks; returns stdouts."""
    script = textwrap.dedent("""
        import os, sys
        sys.path.insert(0, %r)
        import numpy as np, torch, torch.distributed as dist
        rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%%s" %% os.environ["PORT"], rank=rank, world_size=world)
    """ % ROOT) + textwrap.dedent(body) + "\ndist.destroy_process_group()\n"
    port = str(29500 + os.getpid() % 2000)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), PORT=port)
        procs.append(subprocess.Popen([sys.executable, "-c", script], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=timeout)
        assert p.returncode == 0, err[-2000:]
        outs.append(out)
    return outs


def test_shard_is_a_partition():
    from munk_b200.permtest import shard
    for n, w in ((10, 3), (1000, 8), (5, 8), (0, 2)):
        got = sorted(sum((shard(n, r, w) for r in range(w)), []))
        assert got == list(range(n))
        sizes = [len(shard(n, r, w)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_gather_results_two_ranks_gloo():
    outs = run_ranks("""
        from munk_b200.permtest import shard, gather_results
        n = 7
        local = {p: (p + 0.5, 10.0 * p, -1.0 * p) for p in shard(n, rank, world)}
        if rank == 1:
            local[3] = None                      # a failed permutation stays NaN (errs accounting)
        table = gather_results(local, n, rank, world)
        print(repr(table.tolist()))
    """)
    tables = [np.array(eval(o.strip().replace("nan", "float('nan')"))) for o in outs]
    assert np.array_equal(tables[0], tables[1], equal_nan=True)
    t = tables[0]
    assert np.isnan(t[3]).all() and not np.isnan(np.delete(t, 3, axis=0)).any()
    for p in (0, 1, 2, 4, 5, 6):
        assert t[p].tolist() == [p + 0.5, 10.0 * p, -1.0 * p]


def test_statistics_match_reference_formulae():
    from munk_b200.permtest import effect_size, one_tail_pval, permuted_homolog_idxs
    d = np.array([0.1, 0.5, -0.2, 0.3])
    assert one_tail_pval(0.25, d) == (0.5, 2.0, 4.0)
    assert abs(effect_size(0.25, d) - (0.25 - d.mean()) / d.std()) < 1e-15
    idx = [(i, i + 1) for i in range(20)]
    a, b = permuted_homolog_idxs(idx, 3), permuted_homolog_idxs(idx, 3)
    assert a == b and sorted(a) == idx and a != permuted_homolog_idxs(idx, 4)


def test_gather_rows_reassembles_global_order_gloo():
    """dist.gather_rows: serpentine row shards of three ranks (one of them with a short last
    block) come back in global row order on rank 0, nothing on the other ranks."""
    outs = run_ranks("""
        from munk_b200.dist import ColumnCyclic, gather_rows
        n, ncols = 2500, 3                       # 5 blocks of 512 (last: 452 rows) over 3 ranks
        rows = ColumnCyclic(n, world).rows_of(rank)
        local = torch.tensor(rows, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0, -1.0])
        full = gather_rows(local, n, rank, world)
        if rank == 0:
            want = torch.arange(n, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0, -1.0])
            print("ok" if torch.equal(full, want) else "bad")
        else:
            print("none" if full is None else "bad")
    """, world=3)
    assert [o.strip() for o in outs] == ["ok", "none", "none"]


def test_sharded_statistic_plans_add_up_to_the_oracle():
    """dist._local_plan per rank (row flags, H_H cells, |other| of the rank's rows), evaluated with
    numpy on the rank's rows of S and combined like dist_score_means' all-reduce, reproduces the
    oracle's mean(H_H) and mean(other) over the whole matrix — with repeated landmarks, one-to-many
    homologs and homolog cells inside landmark rows / columns."""
    from munk_b200.dist import ColumnCyclic, _local_plan
    from oracle import munk_oracle as oracle
    rng = np.random.default_rng(11)
    n1, n2, world, k = 1700, 260, 3, 12
    ls, lt = rng.choice(n1, k, replace=True), rng.choice(n2, k, replace=True)
    hs = np.concatenate([ls, rng.choice(n1, 90)])
    ht = np.concatenate([lt, rng.choice(n2 // 4, 90)])
    l_idx, h_idx = list(zip(ls.tolist(), lt.tolist())), list(zip(hs.tolist(), ht.tolist()))
    S = rng.random((n1, n2))
    sums = np.zeros(4)
    for rank in range(world):
        rows = ColumnCyclic(n1, world).rows_of(rank)
        plan = _local_plan(len(rows), n2, rows, l_idx, h_idx)
        S_loc = S[rows]
        hh = S_loc[plan.h_row, plan.h_col]
        keep = np.outer(plan.rowL == 0, plan.colL == 0)
        other_mask = keep.copy()
        other_mask[plan.h_row, plan.h_col] = False
        assert int(other_mask.sum()) == plan.n_other and plan.n_hh == len(hh)
        sums += [hh.sum(), plan.n_hh, S_loc[other_mask].sum(), plan.n_other]
    want = oracle.separate_scores(S, l_idx, h_idx)
    assert int(sums[1]) == len(want[3]) and int(sums[3]) == len(want[4])
    assert abs(sums[0] / sums[1] - want[3].mean()) < 1e-12
    assert abs(sums[2] / sums[3] - want[4].mean()) < 1e-12


def test_column_cyclic_map():
    from munk_b200.dist import ColumnCyclic
    cc = ColumnCyclic(2000, 3, nb=512)
    assert cc.nblk == 4 and cc.cols(3) == (1536, 2000)
    # serpentine rounds: 0 1 2 | 2 1 0 | 0 1 2 ... (balances work that shrinks with the block index)
    assert [cc.owner(j) for j in range(4)] == [0, 1, 2, 2]
    assert cc.blocks_of(0) == [0] and cc.blocks_of(2) == [2, 3]
    rows = sorted(sum((cc.rows_of(r) for r in range(3)), []))
    assert rows == list(range(2000))
    assert cc.rows_of(2) == list(range(1024, 2000))
    big = ColumnCyclic(20000, 8)
    assert [big.owner(j) for j in range(18)] == [0, 1, 2, 3, 4, 5, 6, 7, 7, 6, 5, 4, 3, 2, 1, 0, 0, 1]
    assert sorted(sum((big.blocks_of(r) for r in range(8)), [])) == list(range(big.nblk))
    load = [sum(20000 - big.cols(j)[0] for j in big.blocks_of(r)) for r in range(8)]
    assert max(load) < 1.1 * min(load)


def test_staircase_rhs_layout():
    from munk_b200.dist import staircase_rhs
    import torch
    X, stair, n_own_pad = staircase_rhs(1000, list(range(512, 700)), [3, 600, 999], torch.device("cpu"))
    assert n_own_pad == 256 and X.shape == (1000, 384)
    assert stair.tolist() == [512, 640, 3]
    assert X.sum() == 188 + 3 and X[512, 0] == 1 and X[699, 187] == 1 and X[600, 257] == 1
    assert X[:, 188:256].abs().sum() == 0


def test_distributed_cholesky_schedule_two_ranks_gloo():
    """dist_potrf on 2 gloo ranks with the torch-CPU ops: every rank must end with the complete
    factor of the same SPD matrix (look-ahead order, packing and broadcasts are exercised;
    block width 96 gives 5 block columns incl. a ragged one)."""
    outs = run_ranks("""
        sys.path.insert(0, os.path.join(%r, "tests"))
        from cpu_ops import TorchCpuOps
        from munk_b200.dist import dist_potrf
        n = 430
        g = torch.Generator().manual_seed(7)
        X = torch.randn(n, n, dtype=torch.float64, generator=g)
        A0 = X @ X.t() / n + 2 * torch.eye(n, dtype=torch.float64)
        A = torch.zeros(n, 432, dtype=torch.float64)
        A[:, :n] = A0
        linv = torch.zeros(((n + 127) // 128) * 128 * 128, dtype=torch.float64)
        bad = dist_potrf(TorchCpuOps(), A, n, linv, rank, world, nb=96)
        L = torch.tril(A[:, :n])
        err = float((L - torch.linalg.cholesky(A0)).abs().max())
        print(bad, err)
    """ % ROOT)
    for o in outs:
        bad, err = o.split()
        assert int(bad) == 0 and float(err) < 1e-12


def test_onehot_rhs_columns_and_staircase():
    """Right-hand side of the target's fused forward substitution: one-hot columns in landmark-list
    order (repeats allowed), staircase = first non-zero row of each 128-column tile."""
    import torch
    from munk_b200.dist import onehot_rhs
    cols = [700, 5, 5, 999] + list(range(300, 430))
    X, stair = onehot_rhs(1000, cols, torch.device("cpu"))
    assert X.shape == (1000, 256) and stair.tolist() == [5, 424]
    assert float(X.sum()) == len(cols)
    for l, c in enumerate(cols):
        assert X[c, l] == 1.0
    X0, stair0 = onehot_rhs(64, [], torch.device("cpu"))
    assert X0.shape == (64, 128) and stair0.tolist() == [64] and float(X0.abs().sum()) == 0.0
