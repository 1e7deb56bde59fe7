"""Host-side logic of munk_b200 (no GPU): node/homolog index mapping must be bit-exact with the
reference, the separate_scores plan must reproduce the reference's mask order, file formats."""
import re

import networkx as nx
import numpy as np
import pytest

from conftest import graph_from_golden, pairs
from oracle import munk_oracle as oracle


def test_util_matches_oracle_and_reference_indices(golden):
    from munk_b200 import util
    from munk_b200.pipeline import graph_inputs
    G1, G2 = graph_from_golden(golden, "source"), graph_from_golden(golden, "target")
    homs = [tuple(str(x) for x in h) for h in golden["homologs"]]
    assert util.sorted_nodes(G1) == [str(x) for x in golden["source_nodes"]]
    assert util.node_to_index(G2) == oracle.node_to_index(G2)
    assert util.index_to_node(G2) == oracle.index_to_node(G2)
    assert util.node_to_index(["b", "a"], dtype="list") == {"b": 0, "a": 1}
    with pytest.raises(NotImplementedError):
        util.node_to_index(G1, dtype="nope")
    assert util.homologs_in_graphs(G1, G2, homs + [("zz", "yy")]) == homs
    inp = graph_inputs(G1, G2, homs, int(golden["n_landmarks"]))
    assert inp["landmark_idxs"] == pairs(golden["landmark_idxs"])
    assert [list(h) for h in inp["landmark_homs"]] == [list(h) for h in golden["landmark_homs"]]
    # edge arrays describe the same undirected edge set as the oracle's
    def canon(u, v):
        e = np.stack([np.minimum(u, v), np.maximum(u, v)], axis=1)
        return np.unique(e, axis=0)
    assert np.array_equal(canon(*inp["edges1"]), golden["source_edges"])
    assert np.array_equal(canon(*inp["edges2"]), golden["target_edges"])
    with pytest.raises(AssertionError):
        graph_inputs(G1, G2, homs, len(homs) + 1)        # munk.py:129


def test_two_core_and_simplify_match_oracle():
    from munk_b200 import util
    G = nx.barabasi_albert_graph(120, 2, seed=4)          # loses a node in the 2-core
    G.add_edge(5, 5)
    G.add_edges_from([(1000, 1001), (1001, 1002), (1002, 1000)])   # second component
    a = util.simple_two_core(G.copy(), verbose=False)
    b = oracle.simple_two_core(G.copy())
    assert sorted(a.nodes()) == sorted(b.nodes()) and sorted(map(sorted, a.edges())) == sorted(map(sorted, b.edges()))
    assert a.number_of_nodes() < 120 and 1000 not in a and not list(nx.selfloop_edges(a))
    assert min(d for _, d in a.degree()) >= 2


def test_edge_index_arrays_rules():
    from munk_b200 import util
    G = nx.Graph([("a", "b"), ("b", "c"), ("c", "c"), ("c", "d")])
    u, v = util.edge_index_arrays(G, ["c", "a", "b"])     # subset + custom order, self loop dropped
    assert sorted(zip(u.tolist(), v.tolist())) in ([(1, 2), (2, 0)], [(0, 2), (1, 2)], [(1, 2), (0, 2)], [(2, 0), (1, 2)]) \
        or sorted(map(sorted, zip(u.tolist(), v.tolist()))) == [[0, 2], [1, 2]]
    assert u.dtype == np.int32
    with pytest.raises(NotImplementedError):
        util.edge_index_arrays(nx.DiGraph([("a", "b")]), ["a", "b"])


def test_weighted_and_multigraph_edges_carry_networkx_laplacian_entries():
    """Weights and parallel edges: (u, v, lval, ldiag) reproduce nx.laplacian_matrix (munk.py:82)."""
    from munk_b200 import util
    G = nx.Graph()
    G.add_weighted_edges_from([("a", "b", 2.5), ("b", "c", 1.0), ("c", "d", 0.25), ("d", "a", 3.0), ("c", "c", 9.0)])
    M = nx.MultiGraph([("a", "b"), ("a", "b"), ("b", "c"), ("c", "a"), ("c", "a"), ("c", "a")])
    for H in (G, M):
        nodes = sorted(H.nodes())
        u, v, lval, ldiag = util.edge_index_arrays(H, nodes)
        L = np.zeros((len(nodes), len(nodes)))
        L[u, v] = lval
        L[v, u] = lval
        L[np.arange(len(nodes)), np.arange(len(nodes))] = ldiag
        assert np.array_equal(L, np.array(nx.laplacian_matrix(H, nodelist=nodes).todense(), dtype=np.float64))
        assert u.dtype == np.int32 and lval.dtype == np.float64 and (u < v).all()


def test_file_formats_roundtrip(tmp_path):
    from munk_b200 import util
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(60, 50, 3, 20, seeds=(5, 6))
    synthetic.write_edgelist(G1, tmp_path / "s.txt", extras=True)
    synthetic.write_homologs(homs, tmp_path / "h.tsv")
    R = nx.read_edgelist(tmp_path / "s.txt", encoding="ascii")
    R = util.simple_two_core(R, verbose=False)
    assert sorted(R.nodes()) == sorted(G1.nodes()) and R.number_of_edges() == G1.number_of_edges()
    assert util.read_homolog_list(tmp_path / "h.tsv") == homs


def emulate_plan(plan, S):
    """numpy re-statement of separate_scores_kernel's placement rule (munk_b200/csrc/scores.cu)."""
    o_lloff = np.full(plan.n_lloff, np.nan); o_lnonl = np.full(plan.n_lnonl, np.nan)
    o_other = np.full(plan.n_other, np.nan)
    for i in range(plan.n1):
        pc = plan.p_col[plan.p_ptr[i]:plan.p_ptr[i + 1]]
        hc = plan.h_col[plan.h_ptr[i]:plan.h_ptr[i + 1]]
        for j in range(plan.n2):
            rk = plan.colrank[j]
            if plan.rowL[i]:
                if plan.colL[j]:
                    if j not in pc:
                        o_lloff[plan.off_lloff[i] + rk - np.sum(pc < j)] = S[i, j]
                else:
                    o_lnonl[plan.off_lnonl[i] + j - rk] = S[i, j]
            elif plan.colL[j]:
                o_lnonl[plan.off_lnonl[i] + rk] = S[i, j]
            elif j not in hc:
                before = sum(1 for c in hc if c < j and not plan.colL[c])
                o_other[plan.off_other[i] + (j - rk) - before] = S[i, j]
    hh = S[plan.h_row, plan.h_col]
    return S[plan.ls, plan.lt], o_lloff, o_lnonl, hh, o_other


def test_separate_scores_plan_reproduces_reference_order(golden):
    from munk_b200.scores_plan import plan_separate_scores
    S = golden["S"]
    plan = plan_separate_scores(S.shape[0], S.shape[1], pairs(golden["landmark_idxs"]), pairs(golden["homolog_idxs"]))
    got = emulate_plan(plan, S)
    for g, key in zip(got, ("sep_LL_diag", "sep_LL_off", "sep_L_nonL", "sep_HH", "sep_other")):
        assert np.array_equal(g, golden[key]), key


def test_separate_scores_plan_edge_cases():
    from munk_b200.scores_plan import plan_separate_scores
    rng = np.random.default_rng(0)
    S = rng.standard_normal((9, 7))
    # repeated landmark rows/cols, homolog cells inside landmark rows/cols, duplicate homologs
    lidx = [(1, 2), (1, 5), (4, 2), (8, 6)]
    hidx = lidx + [(0, 0), (0, 3), (0, 0), (1, 1), (3, 2), (7, 6), (5, 4)]
    plan = plan_separate_scores(9, 7, lidx, hidx)
    ref = oracle.separate_scores(S, lidx, hidx)
    for g, r in zip(emulate_plan(plan, S), ref):
        assert np.array_equal(g, r)
    with pytest.raises(IndexError):
        plan_separate_scores(9, 7, [(9, 0)], hidx)


def test_cli_flag_surface_matches_reference():
    """Same flags and defaults as scripts/compute_embeddings.py:17-31 of the reference."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "compute_embeddings.py")
    src = open(path).read()
    for flag in ("-se", "--source_edgelist", "-te", "--target_edgelist", "-hf", "--homolog_list", "-so",
                 "--source_output_file", "-to", "--target_output_file", "-sim", "--sim_scores_output_file",
                 "-lo", "--landmarks_output_file", "-r", "--runtimes_file", "-n", "--n_landmarks", "-rs",
                 "--random_seed", "--src_lam", "--tgt_lam"):
        assert re.search(r"'%s'" % re.escape(flag), src), flag
    assert "default=400" in src and "default=28791" in src and src.count("default=0.05") == 2


@pytest.mark.parametrize("n1,n2", [(20000, 16000), (16000, 6000), (1500, 900), (100, 120), (50000, 40000),
                                   (129, 1), (2500, 16000)])
def test_streamed_row_blocks_partition_bottom_up(n1, n2):
    """Launch schedule of the streamed score product: every 128-row tile exactly once, bottom
    block first, the thin top block last."""
    from munk_b200.pipeline import streamed_row_blocks
    tiles, col_tiles = (n1 + 127) // 128, (n2 + 127) // 128
    blocks = streamed_row_blocks(tiles, col_tiles)
    assert blocks[0][1] == tiles and blocks[-1][0] == 0
    assert all(a < b for a, b in blocks)
    assert all(prev[0] == nxt[1] for prev, nxt in zip(blocks[:-1], blocks[1:]))   # contiguous, descending
    if tiles >= 25:
        assert blocks[-1][1] - blocks[-1][0] == min(b - a for a, b in blocks)     # the last one is the thin one
        assert blocks[-1][1] <= 0.04 * tiles + 3


def _digest_c(buf, threads):
    import ctypes
    from munk_b200._lib import lib
    out = (ctypes.c_ulonglong * 2)()
    assert lib.munk_host_digest(ctypes.c_void_p(buf.ctypes.data), buf.nbytes, threads, out) == 0
    return int(out[0]), int(out[1])


def test_host_digest_sees_every_single_element_edit():
    """munk_host_digest (host code of the C library, no GPU): independent of the thread count, and a
    change of any one element - first, last, chunk borders, the odd tail bytes - changes it."""
    rng = np.random.default_rng(5)
    a = rng.standard_normal(3 * (1 << 20) + 5)                   # 24 MB + 40 B: three full 8 MB chunks and a tail
    ref = _digest_c(a, 1)
    assert all(_digest_c(a, t) == ref for t in (2, 3, 8, 64))
    words_per_chunk = (8 << 20) // 8
    spots = [0, 1, 2, 3, 4, words_per_chunk - 1, words_per_chunk, 2 * words_per_chunk + 7, a.size - 1]
    spots += rng.integers(0, a.size, 20).tolist()
    for i in spots:
        keep = a[i]
        a[i] = np.nextafter(keep, np.inf)                        # one ulp
        assert _digest_c(a, 4) != ref, i
        a[i] = keep
    assert _digest_c(a, 4) == ref
    a[5], a[9] = a[9], a[5]                                      # same multiset of words, other order
    assert _digest_c(a, 4)[0] != ref[0] and _digest_c(a, 4)[1] == ref[1]
    b = np.arange(13, dtype=np.uint8)                            # a buffer that is not a whole number of words
    d = _digest_c(b, 2)
    b[12] ^= 1
    assert _digest_c(b, 2) != d
    assert _digest_c(np.empty(0), 4) == (0, 0)
    assert _digest_c(a[:10], 4) != _digest_c(a[:11], 4)          # the length is part of the digest


def test_scores_cache_is_dropped_for_any_in_place_edit():
    """munk.scores serves the device-resident embedding only for the arrays embed_networks returned,
    byte for byte (munk_b200/munk.py::_cached_embedding): an edit of any element, not only of a
    sampled one, invalidates it; so do copies and other views of the same buffer."""
    from munk_b200 import munk as impl
    rng = np.random.default_rng(6)
    C1, P = rng.standard_normal((301, 301)), rng.standard_normal((301, 257))
    C2 = P.T
    emb = object()
    saved = impl._cache
    try:
        impl._cache = (C1, P, emb, impl._digest(C1), impl._digest(P))
        assert impl._cached_embedding(C1, C2) is emb
        assert impl._cached_embedding(C1.copy(), C2) is None and impl._cached_embedding(C1, C2.copy()) is None
        assert impl._cached_embedding(C1, P.T[::-1]) is None and impl._cached_embedding(C1, P) is None
        for arr, idx in ((C1, (123, 45)), (P, (300, 256)), (C1, (0, 0)), (P, (17, 101))):
            keep = arr[idx]
            arr[idx] = keep * 3.0
            assert impl._cached_embedding(C1, C2) is None
            arr[idx] = keep
            assert impl._cached_embedding(C1, C2) is emb
        impl.release_device_cache()
        assert impl._cached_embedding(C1, C2) is None
    finally:
        impl._cache = saved


@pytest.mark.parametrize("seed", range(40))
def test_separate_scores_plan_random_shapes(seed):
    """Random index lists - empty landmark / homolog lists (an error on both sides), every row or column a landmark, repeated
    and one-to-many pairs - through the plan + the kernel's placement rule, against the oracle's
    restatement of munk.py:20-68 (five groups, exact values, exact order)."""
    from munk_b200.scores_plan import plan_separate_scores
    rng = np.random.default_rng(100 + seed)
    n1, n2 = int(rng.integers(1, 14)), int(rng.integers(1, 14))
    S = rng.standard_normal((n1, n2))
    k = int(rng.integers(0, min(n1, n2) + 3)) if seed % 5 else 0
    h = int(rng.integers(0, 2 * (n1 + n2))) if seed % 7 else 0

    def draw(count):
        return [(int(rng.integers(0, n1)), int(rng.integers(0, n2))) for _ in range(count)]

    lidx = draw(k)
    hidx = (lidx if seed % 3 else []) + draw(h)          # the reference takes landmarks = homologs[:k]; both ways here
    if seed % 11 == 0:                                   # every row and every column is a landmark row / column
        lidx = [(i % n1, i % n2) for i in range(max(n1, n2))]
    if not lidx or not hidx:
        # an empty list fails the same way on both sides: `zip(*[])` cannot be unpacked (munk.py:30-31)
        with pytest.raises(ValueError):
            oracle.separate_scores(S, lidx, hidx)
        with pytest.raises(ValueError):
            plan_separate_scores(n1, n2, lidx, hidx)
        return
    plan = plan_separate_scores(n1, n2, lidx, hidx)
    ref = oracle.separate_scores(S, lidx, hidx)
    for name, g, r in zip(("LL_diag", "LL_off", "L_nonL", "HH", "other"), emulate_plan(plan, S), ref):
        assert np.array_equal(np.asarray(g), np.asarray(r)), (name, seed)


def test_numa_local_binding_is_scoped_and_never_fails():
    """munk_b200/hostmem.py: the allocating thread is bound to a CPU set only inside the block, the
    previous affinity comes back afterwards (also after an exception), and every reason not to bind
    - no GPU / sysfs entry, a CPU set outside the process's own, MUNK_NUMA_BIND=0 - is a no-op."""
    import os
    from munk_b200 import hostmem
    assert hostmem.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem.parse_cpulist("5") == {5} and hostmem.parse_cpulist("") == set()
    if not hasattr(os, "sched_getaffinity"):
        pytest.skip("no sched_getaffinity on this platform")
    before = os.sched_getaffinity(0)
    one = {min(before)}
    with hostmem.bound_to(one) as bound:
        assert os.sched_getaffinity(0) == one and bound == (one != before)
    assert os.sched_getaffinity(0) == before
    with pytest.raises(RuntimeError):
        with hostmem.bound_to(one):
            raise RuntimeError("allocation failed")
    assert os.sched_getaffinity(0) == before
    for cpus in (None, set(), {10 ** 6}, set(before)):
        with hostmem.bound_to(cpus) as bound:
            assert not bound and os.sched_getaffinity(0) == before
    with hostmem.numa_local(0) as bound:                 # no GPU here: the lookup fails, nothing is bound
        assert not bound and os.sched_getaffinity(0) == before
    assert os.sched_getaffinity(0) == before
