"""Synthetic scale-free PPI-like inputs for tests and bench.py.

Barabasi-Albert graphs with string node names, a homolog list drawn from seeded permutations,
and writers for the reference's file formats (edgelist: `u v` per line, read by
nx.read_edgelist at scripts/compute_embeddings.py:49; homologs: `src<TAB>tgt`, util.py:55-58).
"""
import networkx as nx
import numpy as np


def ba_graph(n, m, seed, prefix, padded=True):
    """nx.barabasi_albert_graph(n, m, seed) relabelled to names such as s000123 (padded: string
    order == numeric order) or s123 (unpadded: pure lexicographic order)."""
    G = nx.barabasi_albert_graph(n, m, seed=seed)
    fmt = (prefix + "%06d") if padded else (prefix + "%d")
    return nx.relabel_nodes(G, {i: fmt % i for i in range(n)})


def homolog_pairs(source_nodes, target_nodes, count, seed=0):
    """`count` one-to-one pairs: rng.permutation(n1)[:H] zipped with rng.permutation(n2)[:H]."""
    rng = np.random.default_rng(seed)
    a = rng.permutation(len(source_nodes))[:count]
    b = rng.permutation(len(target_nodes))[:count]
    return [(source_nodes[i], target_nodes[j]) for i, j in zip(a, b)]


def make_case(n1, n2, m, n_homologs, seeds=(1, 2), padded=True):
    """(source_G, target_G, homologs) of the BASELINE.json shape n1 / n2 / n_homologs."""
    G1 = ba_graph(n1, m, seeds[0], "s", padded)
    G2 = ba_graph(n2, m, seeds[1], "t", padded)
    homs = homolog_pairs(sorted(G1.nodes()), sorted(G2.nodes()), n_homologs)
    return G1, G2, homs


CONFIGS = {
    # name: (n1, n2, BA m, homologs, landmarks)  — BASELINE.json configs[0..2,4]
    "C1": (2000, 1500, 5, 300, 100),
    "C2": (16000, 6000, 10, 1000, 400),
    "C3": (20000, 16000, 10, 4000, 400),
    "C5": (50000, 40000, 10, 4000, 400),
}


def write_edgelist(G, path, extras=False):
    """One `u v` line per undirected edge; with extras=True also a reversed duplicate, a repeated
    line and a self loop, which nx.Graph de-duplicates and util.simplify_graph removes."""
    edges = list(G.edges())
    with open(path, "w") as f:
        for a, b in edges:
            f.write("%s %s\n" % (a, b))
        if extras and edges:
            a, b = edges[0]
            f.write("%s %s\n" % (b, a))
            f.write("%s %s\n" % (a, b))
            f.write("%s %s\n" % (a, a))


def write_homologs(homologs, path):
    with open(path, "w") as f:
        for a, b in homologs:
            f.write("%s\t%s\n" % (a, b))
