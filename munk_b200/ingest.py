"""Fast graph ingestion: edge-list file -> simple 2-core as sorted node names + int32 edge arrays.

Array (numpy / scipy.sparse.csgraph) restatement of what the reference CLI does with networkx
before the numerics start (scripts/compute_embeddings.py:49-58 -> util.simple_two_core,
munk/munk/util.py:12-53, then util.sorted_nodes / node_to_index, util.py:72-87):

    nx.read_edgelist (whitespace separated, '#' comments, undirected, duplicates collapse)
    -> largest connected component (the first one of maximal size in node-insertion order)
    -> self loops removed  -> 2-core  -> nodes sorted as Python strings  -> index pairs

At 20 000+ nodes the networkx path is the dominant wall-time term of the CLI once the dense
algebra runs on a GPU.  The result is bit-exact with the networkx
path: same node set, same order, same undirected edge set (tests/test_host.py compares them on
messy random inputs).  Host-side integer/string work only; nothing here touches the GPU.
"""
import ast

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components


class WeightedEdgeList(NotImplementedError):
    """The edge list carries non-unit weights: the array path only builds the unweighted index
    arrays; the caller takes the networkx route (util.edge_index_arrays returns the weighted
    Laplacian's entries, assembled by munk_assemble_reglap_entries)."""


def read_edge_pairs(path, comments="#"):
    """(names in first-appearance order, a, b): node ids per line, like nx.read_edgelist
    (lines with fewer than two tokens are skipped; a third column must be a dict literal; a
    non-unit 'weight' in it raises WeightedEdgeList)."""
    ids, names, a, b = {}, [], [], []
    with open(path, "r", encoding="ascii") as fh:
        for line in fh:
            cut = line.find(comments)
            if cut >= 0:
                line = line[:cut]
            tok = line.split()
            if len(tok) < 2:
                continue
            if len(tok) > 2:
                try:
                    data = ast.literal_eval(" ".join(tok[2:]))
                    if not isinstance(data, dict):
                        raise ValueError
                except Exception:
                    raise TypeError("Failed to convert edge data (%s) to dictionary." % tok[2:])
                if data.get("weight", 1) != 1:
                    raise WeightedEdgeList("%s: weighted edge (%s %s %r); load it with networkx "
                                           "(util.edge_index_arrays keeps the weights)" % (path, tok[0], tok[1], data))
            for name, out in ((tok[0], a), (tok[1], b)):
                i = ids.get(name)
                if i is None:
                    i = ids[name] = len(names)
                    names.append(name)
                out.append(i)
    return names, np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64)


def two_core_ids(n, a, b):
    """Boolean mask (length n, first-appearance ids) of the nodes in the simple 2-core of the
    largest connected component; ties between components of equal size go to the one that
    contains the earliest node (networkx yields components in node order, max() keeps the first)."""
    if n == 0:
        raise ValueError("Connectivity is undefined for the null graph.")
    adj = coo_matrix((np.ones(len(a), dtype=np.int8), (a, b)), shape=(n, n))
    ncomp, label = connected_components(adj, directed=False)
    if ncomp > 1:
        sizes = np.bincount(label, minlength=ncomp)
        first = np.full(ncomp, n, dtype=np.int64)
        np.minimum.at(first, label, np.arange(n))
        best = sizes.max()
        cand = np.flatnonzero(sizes == best)
        keep_comp = cand[np.argmin(first[cand])]
        alive = label == keep_comp
    else:
        alive = np.ones(n, dtype=bool)
    # undirected simple edge set inside the component (self loops dropped, duplicates collapsed)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    ok = (lo != hi) & alive[lo] & alive[hi]
    key = np.unique(lo[ok] * n + hi[ok])
    lo, hi = key // n, key % n
    # peel vertices of degree < 2 until none is left (the 2-core is unique)
    while True:
        deg = np.bincount(lo, minlength=n) + np.bincount(hi, minlength=n)
        drop = alive & (deg < 2)
        if not drop.any():
            break
        alive &= ~drop
        ok = alive[lo] & alive[hi]
        lo, hi = lo[ok], hi[ok]
    return alive, lo, hi


def load_two_core(path):
    """Edge-list file -> (nodes, u, v): node names of the simple 2-core sorted like Python's
    sorted() (util.sorted_nodes) and the undirected edges as int32 positions in that order."""
    names, a, b = read_edge_pairs(path)
    alive, lo, hi = two_core_ids(len(names), a, b)
    kept = np.flatnonzero(alive)
    kept_names = [names[i] for i in kept]
    order = sorted(range(len(kept_names)), key=kept_names.__getitem__)      # Python string order
    nodes = [kept_names[i] for i in order]
    pos = np.full(len(names), -1, dtype=np.int64)
    pos[kept[np.asarray(order, dtype=np.int64)]] = np.arange(len(order))
    return nodes, pos[lo].astype(np.int32), pos[hi].astype(np.int32)


def homologs_in_node_lists(source_nodes, target_nodes, homologs):
    """util.homologs_in_graphs (util.py:60-69) on node lists: order and repeats preserved."""
    s, t = set(source_nodes), set(target_nodes)
    return [(a, b) for a, b in homologs if a in s and b in t]
