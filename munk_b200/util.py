"""Graph pruning and the node/homolog index mapping (reference: munk/munk/util.py).

Integer/string host logic; results are bit-exact with the reference: node order is Python's
sorted() over node names (util.py:86-87), homolog filtering keeps file order and duplicates
(util.py:60-69).  simplify_graph / simple_two_core are restated for networkx >= 3, where
connected_component_subgraphs and Graph.selfloop_edges (util.py:19,30) no longer exist.
"""
import numpy as np

from .io import get_logger

# networkx is imported where it is used (simplify_graph, simple_two_core): it costs ~0.7 s at
# start-up and the array ingestion path of the CLI never touches it.


def sorted_nodes(G):
    return sorted(G.nodes())


def node_to_index(obj, dtype='graph'):
    if dtype == 'graph':
        names = sorted_nodes(obj)
    elif dtype == 'list':
        names = obj
    else:
        raise NotImplementedError()
    return {name: pos for pos, name in enumerate(names)}


def index_to_node(obj, dtype='graph'):
    if dtype == 'graph':
        return dict(enumerate(sorted_nodes(obj)))
    if dtype == 'list':
        return dict(enumerate(obj))
    return None  # the reference falls through and returns None for other dtypes


def read_homolog_list(fp):
    """Two-column (whitespace separated) homolog file -> list of tuples, file order."""
    pairs = []
    with open(fp, 'r') as handle:
        for line in handle:
            pairs.append(tuple(line.split()))
    return pairs


def homologs_in_graphs(G1, G2, homologs):
    """Pairs whose source is in G1 and target in G2; order and repeats preserved."""
    in1, in2 = set(G1.nodes()), set(G2.nodes())
    return [(a, b) for a, b in homologs if a in in1 and b in in2]


def simplify_graph(G, verbose=True):
    """Largest connected component (first one of maximal size), then self loops removed."""
    import networkx as nx
    log = get_logger()
    if not nx.is_connected(G):
        parts = [G.subgraph(nodes).copy() for nodes in nx.connected_components(G)]
        sizes = [len(p) for p in parts]
        biggest = max(sizes)
        sizes.remove(biggest)
        if verbose:
            log.warning('Network has %d connected components', len(parts))
            log.warning('\tLargest is size %d and all the rest are %d or smaller',
                        biggest, max(sizes))
            log.warning('\tUsing largest connected component')
        G = max(parts, key=len)
    G.remove_edges_from(list(nx.selfloop_edges(G)))
    return G


def simple_two_core(G, verbose=True):
    """2-core of the simplified graph."""
    import networkx as nx
    log = get_logger()
    G = simplify_graph(G, verbose)
    n_nodes, n_edges = G.number_of_nodes(), G.number_of_edges()
    if verbose:
        log.info('PPI info - # Nodes: %d, # Edges: %d', n_nodes, n_edges)
        log.info('Computing 2 core')
    core = nx.k_core(G, 2)
    if verbose:
        log.info('2 core info - # Nodes: %d, # Edges: %d',
                 core.number_of_nodes(), core.number_of_edges())
        log.info('2 core removed %d nodes and %d edges',
                 n_nodes - core.number_of_nodes(), n_edges - core.number_of_edges())
    return core


def edge_index_arrays(G, nodes):
    """Edges of G restricted to `nodes`, as int32 arrays of positions in `nodes`.

    Simple unweighted graphs (every PPI edge list the reference reads): returns (u, v), each
    undirected edge once (canonical (min, max) pairs, duplicates removed), self loops dropped —
    they cancel in the Laplacian nx.laplacian_matrix builds (munk.py:82).

    Weighted graphs and MultiGraphs: returns (u, v, lval, ldiag) — the off-diagonal entries
    L[u, v] (each unordered pair once) and the diagonal of nx.laplacian_matrix(G, nodelist=nodes)
    itself, float64, so that parallel edges and weights are summed exactly as the reference sums
    them; the assembly kernel then forms 1 + lam * L entry by entry.  Directed graphs are refused:
    their Laplacian is not symmetric, and the Cholesky route needs an SPD matrix.
    """
    pos = {name: i for i, name in enumerate(nodes)}
    if len(pos) != len(nodes):
        raise ValueError('node list contains duplicates')
    if G.is_directed():
        raise NotImplementedError('directed graphs are not supported: I + lam*L must be symmetric')
    weighted = G.is_multigraph()
    us, vs = [], []
    for a, b, data in G.edges(data=True):
        if a == b or a not in pos or b not in pos:
            continue
        if data and data.get('weight', 1) != 1:
            weighted = True
            break
        us.append(pos[a])
        vs.append(pos[b])
    if weighted:
        return _laplacian_entries(G, nodes)
    u, v = np.asarray(us, dtype=np.int64), np.asarray(vs, dtype=np.int64)
    lo, hi = np.minimum(u, v), np.maximum(u, v)
    key = np.unique(lo * len(nodes) + hi)               # nx.Graph yields each edge once already;
    if len(key) != len(lo):                             # this guards graph-like inputs that do not
        lo, hi = key // len(nodes), key % len(nodes)
        return lo.astype(np.int32), hi.astype(np.int32)
    return u.astype(np.int32), v.astype(np.int32)


def _laplacian_entries(G, nodes):
    """(u, v, lval, ldiag) from nx.laplacian_matrix itself (munk.py:82), strict upper triangle."""
    import networkx as nx
    import scipy.sparse
    L = scipy.sparse.coo_matrix(nx.laplacian_matrix(G, nodelist=nodes).astype(np.float64))
    ldiag = np.zeros(len(nodes), dtype=np.float64)
    on = L.row == L.col
    np.add.at(ldiag, L.row[on], L.data[on])
    up = L.row < L.col
    return (L.row[up].astype(np.int32), L.col[up].astype(np.int32),
            np.ascontiguousarray(L.data[up], dtype=np.float64), ldiag)
