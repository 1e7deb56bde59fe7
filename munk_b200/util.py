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
    """Edges of G restricted to `nodes`, as two int32 arrays of positions in `nodes`.

    Each undirected edge appears once, self loops are dropped (they cancel in the Laplacian
    nx.laplacian_matrix builds), and a weighted graph is refused: the reference's PPI edge
    lists are unweighted and the assembly kernel writes -lam per edge.
    """
    pos = {name: i for i, name in enumerate(nodes)}
    if len(pos) != len(nodes):
        raise ValueError('node list contains duplicates')
    us, vs = [], []
    for a, b, data in G.edges(data=True):
        if a == b or a not in pos or b not in pos:
            continue
        if data and data.get('weight', 1) != 1:
            raise NotImplementedError('weighted graphs are not supported by the B200 assembly kernel')
        us.append(pos[a])
        vs.append(pos[b])
    return np.asarray(us, dtype=np.int32), np.asarray(vs, dtype=np.int32)
