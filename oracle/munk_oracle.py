"""CPU oracle for MUNK's dense hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference algorithm (lrgr/munk), function by function, each
citing the reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the checker or the
timed CPU baseline; nothing under munk_b200/ imports it.

Parity pin: the reference ships no tests or golden vectors, so the oracle
is pinned against outputs of the UNMODIFIED reference itself, generated in the build container
by oracle/gen_golden.py (imports /root/reference/munk under a sklearn.externals.joblib shim) and
committed under tests/golden/.  tests/test_oracle.py checks every function here against them.

The arithmetic lives in third-party libraries the reference calls (numpy.linalg.inv -> LAPACK
gesv, numpy.linalg.pinv -> gesdd with rcond 1e-15, scipy.linalg.eigh -> syevr, numpy.dot -> BLAS
dgemm, networkx.laplacian_matrix); the reference pins them loosely (munk/setup.py:26-31:
numpy>=1.14,<2  scipy>=0.19.1,<1.0.0  networkx>=2.0,<3.0).  This file restates the same calls on
the versions installed here (numpy 2.3, scipy 1.18, networkx 3.6).
"""
import numpy as np
import scipy.linalg


# --------------------------------------------------------------------------------------
# graph / index helpers  (reference: munk/munk/util.py)
# --------------------------------------------------------------------------------------
def sorted_nodes(G):
    """util.py:86-87 — canonical node order is Python's sorted() of the node names."""
    return sorted(G.nodes())


def node_to_index(obj, dtype="graph"):
    """util.py:72-78."""
    if dtype == "graph":
        return {n: i for i, n in enumerate(sorted_nodes(obj))}
    if dtype == "list":
        return {n: i for i, n in enumerate(obj)}
    raise NotImplementedError()


def index_to_node(obj, dtype="graph"):
    """util.py:80-84."""
    if dtype == "graph":
        return dict(enumerate(sorted_nodes(obj)))
    if dtype == "list":
        return dict(enumerate(obj))


def read_homolog_list(fp):
    """util.py:55-58 — whitespace-split tuples in file order."""
    with open(fp, "r") as f:
        return [tuple(line.split()) for line in f]


def homologs_in_graphs(G1, G2, homologs):
    """util.py:60-69 — keep pairs present in both graphs, order preserved, no de-duplication."""
    n1, n2 = set(G1.nodes()), set(G2.nodes())
    return [(a, b) for a, b in homologs if a in n1 and b in n2]


def simplify_graph(G):
    """util.py:12-31 restated for networkx 3 (connected_component_subgraphs / G.selfloop_edges
    are gone): largest connected component (first maximum on ties), then drop self loops."""
    import networkx as nx
    if not nx.is_connected(G):
        comps = [G.subgraph(c).copy() for c in nx.connected_components(G)]
        G = max(comps, key=len)
    G.remove_edges_from(list(nx.selfloop_edges(G)))
    return G


def simple_two_core(G):
    """util.py:33-53 — simplify_graph then nx.k_core(G, 2)."""
    import networkx as nx
    return nx.k_core(simplify_graph(G), 2)


def edge_index_arrays(G, nodes):
    """Undirected edges of G as index pairs in `nodes` order (each edge once, u < v)."""
    n2i = {n: i for i, n in enumerate(nodes)}
    e = np.array([(n2i[a], n2i[b]) for a, b in G.edges() if a != b], dtype=np.int64).reshape(-1, 2)
    e.sort(axis=1)
    return np.unique(e, axis=0)


# --------------------------------------------------------------------------------------
# numeric core  (reference: munk/munk/munk.py)
# --------------------------------------------------------------------------------------
def laplacian_dense(n, edges):
    """munk.py:82 — np.array(nx.laplacian_matrix(G, nodelist=nodes).todense()) for an unweighted
    simple graph: integer matrix, deg on the diagonal, -1 per edge."""
    L = np.zeros((n, n), dtype=np.int64)
    u, v = edges[:, 0], edges[:, 1]
    L[u, v] = -1
    L[v, u] = -1
    np.add.at(L, (u, u), 1)
    np.add.at(L, (v, v), 1)
    return L


def reglap_matrix(n, edges, lam):
    """munk.py:83, the matrix being inverted: np.eye(n) + lam * L."""
    return np.eye(n) + (lam * laplacian_dense(n, edges))


def regularized_laplacian(G, nodes, lam):
    """munk.py:74-83 — D = inv(I + lam * L)."""
    n = len(nodes)
    return np.linalg.inv(reglap_matrix(n, edge_index_arrays(G, nodes), lam))


def rkhs_factor(D):
    """munk.py:85-88 — e, v = eigh(D); C = v diag(sqrt(e))."""
    e, v = scipy.linalg.eigh(D)
    return v.dot(np.diag(np.sqrt(e)))


def embed_matrices(source_C, target_D, landmark_idxs):
    """munk.py:97-112 — C2hat = (pinv(C1[Ls,:]) . D2[Lt,:]).T ; returns (source_C, C2hat)."""
    s_idx, t_idx = zip(*landmark_idxs)
    C2 = np.linalg.pinv(source_C[s_idx, :]).dot(target_D[t_idx, :]).T
    return source_C, C2


def scores(source_C, target_C_hat):
    """scripts/compute_embeddings.py:81 / permtest.py:55 — S = C1 . C2hat^T."""
    return np.dot(source_C, target_C_hat.T)


def embed_networks(source_G, target_G, homologs, n_landmarks, src_lam=0.05, tgt_lam=0.05):
    """munk.py:114-176 (return_idxs=True shape, plus the landmark name pairs)."""
    assert n_landmarks <= len(homologs)                           # munk.py:129
    s_nodes, t_nodes = sorted_nodes(source_G), sorted_nodes(target_G)
    D1 = regularized_laplacian(source_G, s_nodes, src_lam)
    C1 = rkhs_factor(D1)
    D2 = regularized_laplacian(target_G, t_nodes, tgt_lam)
    s_n2i, t_n2i = node_to_index(source_G), node_to_index(target_G)
    landmark_homs = homologs[:n_landmarks]                        # munk.py:152
    landmark_idxs = [(s_n2i[a], t_n2i[b]) for a, b in landmark_homs]
    homolog_idxs = [(s_n2i[a], t_n2i[b]) for a, b in homologs]    # munk.py:167-168
    C1, C2 = embed_matrices(C1, D2, landmark_idxs)
    return dict(C1=C1, C2=C2, D1=D1, D2=D2, source_nodes=s_nodes, target_nodes=t_nodes,
                landmark_homs=landmark_homs, landmark_idxs=landmark_idxs, homolog_idxs=homolog_idxs)


def separate_scores(S, landmark_pair_idxs, homolog_pair_idxs):
    """munk.py:20-68 — five boolean-mask selections of S, each in row-major order."""
    ls, lt = zip(*landmark_pair_idxs)
    hs, ht = zip(*homolog_pair_idxs)
    rows = np.zeros(S.shape, dtype=bool); rows[ls, :] = True          # munk.py:34-35
    cols = np.zeros(S.shape, dtype=bool); cols[:, lt] = True          # munk.py:36-37
    pair = np.zeros(S.shape, dtype=bool); pair[ls, lt] = True         # munk.py:38-39
    LL_diag = S[ls, lt]                                               # munk.py:42
    LL_off = S[(rows & cols) & ~pair]                                 # munk.py:45-47
    L_nonL = S[rows ^ cols]                                           # munk.py:50-51
    hh = np.zeros(S.shape, dtype=bool); hh[hs, ht] = True; hh[ls, lt] = False   # munk.py:54-57
    H_H = S[hh]
    other = S[~(rows | cols | hh)]                                    # munk.py:60-62
    return LL_diag, LL_off, L_nonL, H_H, other


# --------------------------------------------------------------------------------------
# permutation-test unit  (reference: experiments/homology-scores-permutation-test/permtest.py)
# --------------------------------------------------------------------------------------
def difference_in_means(source_C, target_D, homolog_idxs, n_landmarks):
    """permtest.py:37-80 — embed with the first n_landmarks pairs, S = C1 @ C2hat.T, then
    mean(H_H) - mean(other); returns (diff, other_mean, hom_mean)."""
    landmark_idxs = homolog_idxs[:n_landmarks]
    C1, C2 = embed_matrices(source_C, target_D, landmark_idxs)
    S = C1 @ C2.T
    _, _, _, hh, other = separate_scores(S, landmark_idxs, homolog_idxs)
    other_mean, hom_mean = np.mean(other), np.mean(hh)
    return hom_mean - other_mean, other_mean, hom_mean


def one_tail_pval(observed, permuted):
    """permtest.py:101-104."""
    permuted = np.asarray(permuted)
    n_sig = float(np.sum(permuted > observed))
    return n_sig / float(len(permuted)), n_sig, float(len(permuted))


def effect_size(observed, control):
    """permtest.py:106-109 — population std (np.std)."""
    return (observed - np.mean(control)) / np.std(control)


# --------------------------------------------------------------------------------------
# full-size checker: the same scores without inv / eigh / pinv (seconds instead of minutes)
# --------------------------------------------------------------------------------------
def spd_matrix(n, edges, lam):
    """np.eye(n) + lam * L (munk.py:83) built in one float64 buffer: the diagonal is
    fl(1 + fl(lam * deg)), off-diagonals fl(lam * -1) — the same roundings as reglap_matrix()
    (checked bit for bit in tests/test_oracle.py) without its int64 and float temporaries, so a
    20 000-node matrix costs 3.2 GB instead of 13."""
    A = np.zeros((n, n), dtype=np.float64)
    u, v = edges[:, 0], edges[:, 1]
    A[u, v] = lam * -1.0
    A[v, u] = lam * -1.0
    deg = np.bincount(np.concatenate([u, v]), minlength=n).astype(np.float64)
    A[np.arange(n), np.arange(n)] = 1.0 + lam * deg
    return A


def scores_by_cholesky(n1, edges1, n2, edges2, landmark_idxs, src_lam=0.05, tgt_lam=0.05):
    """The score matrix of embed_networks + scores (munk.py:114-176, compute_embeddings.py:81)
    computed on the CPU WITHOUT the reference's inv / eigh / pinv calls, for checks at sizes where
    those take minutes (BASELINE configs 2, 3, 5).  For distinct source landmarks Ls the rows
    M = C1[Ls,:] are independent (M M^T = D1[Ls,Ls] is a principal block of an SPD matrix), so
    pinv(M) = M^T (M M^T)^-1 exactly and

        S = C1 pinv(M) D2[Lt,:] = D1[:,Ls] D1[Ls,Ls]^-1 D2[Lt,:]         

    which needs only k columns of each kernel: two LAPACK Cholesky factorisations (scipy
    cho_factor) and 2k right-hand sides.  Independent of every GPU result.  Pinned against the
    reference's own S on the golden cases in tests/test_oracle.py (<= 1e-12 of max|S|).
    Returns dict(S, D1_cols = D1[:,Ls] (n1 x k), D2_rows = D2[Lt,:] (k x n2))."""
    ls = np.array([a for a, _ in landmark_idxs], dtype=np.int64)
    lt = np.array([b for _, b in landmark_idxs], dtype=np.int64)
    assert len(set(ls.tolist())) == len(ls), "scores_by_cholesky needs distinct source landmarks"
    k = len(ls)

    def kernel_columns(n, edges, lam, cols):
        A = spd_matrix(n, np.asarray(edges).reshape(-1, 2), lam)
        c = scipy.linalg.cho_factor(A, lower=True, overwrite_a=True, check_finite=False)
        E = np.zeros((n, len(cols)), order="F")
        E[cols, np.arange(len(cols))] = 1.0
        return scipy.linalg.cho_solve(c, E, overwrite_b=True, check_finite=False)

    X = kernel_columns(n1, edges1, src_lam, ls)                   # D1[:, Ls]
    B = kernel_columns(n2, edges2, tgt_lam, lt).T                 # D2[Lt, :] (D2 symmetric)
    G = X[ls, :]                                                  # D1[Ls, Ls]
    G = 0.5 * (G + G.T)
    Z = scipy.linalg.cho_solve(scipy.linalg.cho_factor(G, lower=True), np.ascontiguousarray(B))
    return dict(S=X @ Z, D1_cols=X, D2_rows=np.ascontiguousarray(B))


# --------------------------------------------------------------------------------------
# parity metrics used by the tests
# --------------------------------------------------------------------------------------
def rel_max_err(x, ref):
    """max|x - ref| / max|ref| — the north-star score metric (tolerance 1e-8)."""
    return float(np.max(np.abs(np.asarray(x) - ref)) / np.max(np.abs(ref)))


def procrustes_residual(X, X_ref, *others):
    """min_Q |X Q - X_ref|_max over orthogonal Q (SVD solution); the same Q is applied to each
    (Y, Y_ref) in `others`.  Returns the list of max-abs residuals relative to max|ref|."""
    U, _, Vt = np.linalg.svd(X.T @ X_ref)
    Q = U @ Vt
    out = [rel_max_err(X @ Q, X_ref)]
    for Y, Y_ref in others:
        out.append(rel_max_err(Y @ Q, Y_ref))
    return out
