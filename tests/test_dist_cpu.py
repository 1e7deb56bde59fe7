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
