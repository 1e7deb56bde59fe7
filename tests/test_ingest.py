"""munk_b200.ingest (array path) must reproduce the reference's networkx path bit for bit:
nx.read_edgelist -> util.simple_two_core -> util.sorted_nodes / node_to_index."""
import random

import networkx as nx
import numpy as np
import pytest

from munk_b200 import ingest, util
from oracle import munk_oracle as oracle


def messy_edgelist(path, seed, n=300, m=2, extra_components=3):
    """Scale-free core + pendant trees + extra components (one as large as possible ties) + self
    loops + duplicate / reversed lines + comments + blank lines, unpadded names, shuffled order."""
    rng = random.Random(seed)
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    lines = ["g%d g%d" % (a, b) for a, b in G.edges()]
    for t in range(20):                                  # pendant chains (removed by the 2-core)
        a = rng.randrange(n)
        lines += ["g%d p%d_0" % (a, t), "p%d_0 p%d_1" % (t, t)]
    for c in range(extra_components):                    # small cycles in their own components
        k = rng.randrange(3, 7)
        lines += ["c%d_%d c%d_%d" % (c, i, c, (i + 1) % k) for i in range(k)]
    lines += ["g%d g%d" % (a, a) for a in rng.sample(range(n), 5)]          # self loops
    lines += ["lonely lonely"]                                               # a self-loop-only node
    lines += [" ".join(reversed(l.split())) for l in rng.sample(lines[:200], 30)]   # reversed duplicates
    rng.shuffle(lines)
    lines.insert(3, "# a comment line")
    lines.insert(10, "")
    lines[20] += "   # trailing comment"
    lines[25] += " {}"                                   # empty data dict (legal third column)
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4])
def test_array_ingest_equals_networkx_path(tmp_path, seed):
    p = tmp_path / "edges.txt"
    messy_edgelist(p, seed)
    G = util.simple_two_core(nx.read_edgelist(p, encoding="ascii"), verbose=False)
    want_nodes = util.sorted_nodes(G)
    wu, wv = util.edge_index_arrays(G, want_nodes)
    nodes, u, v = ingest.load_two_core(p)
    assert nodes == want_nodes
    assert u.dtype == np.int32 and v.dtype == np.int32

    def canon(a, b):
        e = np.stack([np.minimum(a, b), np.maximum(a, b)], axis=1)
        return np.unique(e, axis=0)
    assert np.array_equal(canon(u, v), canon(wu, wv))
    assert len(u) == G.number_of_edges()
    assert np.array_equal(canon(u, v), oracle.edge_index_arrays(oracle.simple_two_core(
        nx.read_edgelist(p, encoding="ascii")), want_nodes))


def test_component_tie_goes_to_first_in_file_order(tmp_path):
    """Two components of equal size: networkx keeps the one whose first node appears first."""
    tri = lambda s: ["%s1 %s2" % (s, s), "%s2 %s3" % (s, s), "%s3 %s1" % (s, s)]  # noqa: E731
    for order in (("z", "a"), ("a", "z")):
        p = tmp_path / ("tie_%s.txt" % order[0])
        p.write_text("\n".join(tri(order[0]) + tri(order[1])) + "\n")
        G = util.simple_two_core(nx.read_edgelist(p, encoding="ascii"), verbose=False)
        nodes, u, v = ingest.load_two_core(p)
        assert nodes == util.sorted_nodes(G) == ["%s1" % order[0], "%s2" % order[0], "%s3" % order[0]]


def _cli_module():
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "compute_embeddings.py")
    spec = importlib.util.spec_from_file_location("cli_compute_embeddings", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_weighted_edge_list_takes_the_networkx_route(tmp_path):
    """nx.read_edgelist keeps a {'weight': w} column and nx.laplacian_matrix honours it (munk.py:82):
    the array ingestion hands such a file over (WeightedEdgeList) and the CLI's loader returns the
    weighted Laplacian's own entries, equal to networkx's matrix on the 2-core."""
    import logging
    p = tmp_path / "w.txt"
    p.write_text("a b {'weight': 2.0}\nb c\nc a {'weight': 0.5}\nc d\nd a\ne a\n# comment\nc c\n")
    with pytest.raises(ingest.WeightedEdgeList):
        ingest.load_two_core(p)
    assert issubclass(ingest.WeightedEdgeList, NotImplementedError)
    cli = _cli_module()
    nodes, edges = cli.load_with_arrays(str(p), logging.getLogger("test"), "source")
    G = util.simple_two_core(nx.read_edgelist(p, encoding="ascii"), verbose=False)
    assert nodes == util.sorted_nodes(G) == ["a", "b", "c", "d"]            # 'e' is pruned, the self loop dropped
    u, v, lval, ldiag = edges
    L = np.zeros((4, 4))
    L[u, v] = lval
    L = L + L.T + np.diag(ldiag)
    assert np.array_equal(L, nx.laplacian_matrix(G, nodelist=nodes).toarray())
    # an unweighted file stays on the array path: plain (u, v)
    p.write_text("a b\nb c\nc a {'weight': 1}\n")
    nodes, edges = cli.load_with_arrays(str(p), logging.getLogger("test"), "source")
    assert nodes == ["a", "b", "c"] and len(edges) == 2


def test_ingest_refuses_bad_data(tmp_path):
    p = tmp_path / "w.txt"
    p.write_text("a b 3.5\nb c\nc a\n")
    with pytest.raises(TypeError):                       # nx.read_edgelist raises TypeError here too
        ingest.load_two_core(p)
    with pytest.raises(TypeError):
        nx.read_edgelist(p, encoding="ascii")


def test_homolog_filter_and_index_inputs_match_graph_path(tmp_path):
    from munk_b200 import pipeline
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(120, 90, 3, 30, seeds=(5, 6), padded=False)
    synthetic.write_edgelist(G1, tmp_path / "s.txt", extras=True)
    synthetic.write_edgelist(G2, tmp_path / "t.txt", extras=True)
    homs = homs + [("missing", "t3"), (homs[0][0], homs[1][1])]
    n1, u1, v1 = ingest.load_two_core(tmp_path / "s.txt")
    n2, u2, v2 = ingest.load_two_core(tmp_path / "t.txt")
    H1 = util.simple_two_core(nx.read_edgelist(tmp_path / "s.txt", encoding="ascii"), verbose=False)
    H2 = util.simple_two_core(nx.read_edgelist(tmp_path / "t.txt", encoding="ascii"), verbose=False)
    valid = ingest.homologs_in_node_lists(n1, n2, homs)
    assert valid == util.homologs_in_graphs(H1, H2, homs)
    a = pipeline.index_inputs(n1, (u1, v1), n2, (u2, v2), valid, 10)
    b = pipeline.graph_inputs(H1, H2, valid, 10)
    for key in ("source_nodes", "target_nodes", "landmark_homs", "landmark_idxs", "s_pos", "t_pos"):
        assert a[key] == b[key], key
