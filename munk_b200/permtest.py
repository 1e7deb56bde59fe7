"""Homology-scores permutation test on B200s
(reference: experiments/homology-scores-permutation-test/permtest.py, gen_random_network.py).

The per-permutation unit is the reference's `difference_in_means` (permtest.py:68-80):
embed_matrices + S = C1 @ C2hat.T + separate_scores + mean(H_H) - mean(other).  Two kinds of
permutation are supported ("C4 semantics"):

  "indices" (BASELINE.json config 4): the graphs, hence the factors W1, W2, stay fixed and
      permutation p shuffles the homolog list with np.random.default_rng(1000 + p) before the
      first n_landmarks pairs are taken as landmarks;
  "graphs"  (the reference's own experiment): every permutation draws a degree-preserving random
      graph pair (gen_random_network.py:50-84), factors it on the GPU, and keeps the homolog list.

Permutations are independent units: rank r of `world` takes p = r, r + world, ... with no
data-path collective; one all-gather of 3 doubles per permutation at the end.
"""
import numpy as np
import torch

from . import util
from .device import default_device, landmark_groups
from .pipeline import graph_inputs
from .scores_plan import plan_separate_scores


# ------------------------------------------------------------------ statistics (host scalars)
def one_tail_pval(observed, permuted):
    """permtest.py:101-104: (#permuted > observed) / N, the count, N."""
    permuted = np.asarray(permuted)
    n_significant = float(np.sum(permuted > observed))
    return n_significant / float(len(permuted)), n_significant, float(len(permuted))


def effect_size(observed, control):
    """permtest.py:106-109: z-score with the population standard deviation."""
    return (observed - np.mean(control)) / np.std(control)


# ------------------------------------------------------------------ the unit, host-array API
def munk_embed_wrapper(source_data, target_data, homologs, n_landmarks, ref_source_nodes,
                       ref_target_nodes):
    """permtest.py:37-56 — host dicts in (source_data['C'], target_data['D'], node lists),
    (sim_scores, landmark_idxs, homolog_idxs) out."""
    from .munk import embed_matrices, scores
    source_nodes, target_nodes = source_data['nodes'], target_data['nodes']
    assert source_nodes == ref_source_nodes
    assert target_nodes == ref_target_nodes
    s_pos = {n: i for i, n in enumerate(source_nodes)}
    t_pos = {n: i for i, n in enumerate(target_nodes)}
    homolog_idxs = [(s_pos[a], t_pos[b]) for a, b in homologs]
    landmark_idxs = homolog_idxs[:n_landmarks]
    source_C, target_C_hat = embed_matrices(source_data['C'], target_data['D'], landmark_idxs)
    return scores(source_C, target_C_hat), landmark_idxs, homolog_idxs


def dissim_diff(sim_scores, landmark_idxs, homolog_idxs):
    """permtest.py:58-66 — (hom_mean - other_mean, other_mean, hom_mean), computed on the device
    without materialising the five selections."""
    dev = default_device()
    S = sim_scores if torch.is_tensor(sim_scores) else \
        torch.from_numpy(np.ascontiguousarray(sim_scores, dtype=np.float64)).to(dev.torch_device)
    plan = plan_separate_scores(S.shape[0], S.shape[1], landmark_idxs, homolog_idxs)
    hom_mean, other_mean, diff = plan.means(dev, S).cpu().numpy()[:3]
    return float(diff), float(other_mean), float(hom_mean)


def difference_in_means(source_data, target_data, homologs, n_landmarks, ref_source_nodes,
                        ref_target_nodes):
    """permtest.py:68-80."""
    return dissim_diff(*munk_embed_wrapper(source_data, target_data, homologs, n_landmarks,
                                           ref_source_nodes, ref_target_nodes))


# ------------------------------------------------------------------ the unit, device resident
class FactoredPair:
    """W1, W2 (inverse Cholesky factors of I + lam L) of one graph pair, on the device."""

    def __init__(self, dev, n1, edges1, n2, edges2, lam=0.05):
        self.dev, self.n1, self.n2 = dev, n1, n2
        self.W1 = dev.kernel_factor_edges(n1, edges1, lam)
        self.W2 = dev.kernel_factor_edges(n2, edges2, lam)
        self._S = None

    def statistic(self, homolog_idxs, n_landmarks):
        """CUDA tensor {hom_mean, other_mean, diff, |H_H|, |other|}; nothing is synchronised."""
        dev = self.dev
        landmark_idxs = homolog_idxs[:n_landmarks]
        uniq, gptr, members = landmark_groups([int(a) for a, _ in landmark_idxs])
        Q1 = dev.gather_cols_lower(self.W1, self.n1, uniq)
        B = dev.target_landmark_rows(self.W2, self.n2, [int(b) for _, b in landmark_idxs])
        P = dev.project(Q1, self.n1, B, self.n2, (gptr, members))
        if self._S is None:
            self._S = dev.empty(self.n1, self.n2)
        S = dev.score(self.W1, self.n1, P, self.n2, out=self._S)
        plan = plan_separate_scores(self.n1, self.n2, landmark_idxs, homolog_idxs)
        return plan.means(dev, S)


    def statistic_rank_k(self, homolog_idxs, n_landmarks):
        """The same three numbers WITHOUT forming S (opt-in fast path).

        With distinct source landmarks S = C1 C2hat^T = X Z with X = D1[:,Ls] = W1^T Q1 (n1 x k)
        and Z = D1[Ls,Ls]^-1 mean_groups(D2[Lt,:]) (k x n2), so
            sum over a row set R and a column set C of S  =  (1_R^T X) (Z 1_C)
        and a single score is a k-long dot product.  Cost per permutation: n1^2 k flops for X
        instead of n1^2 n2 for S (about 16x less time at 20k/16k/400).  Returns a host tuple
        (hom_mean, other_mean, diff); validated against `statistic` in the GPU tests."""
        from .device import KC, MC, KLO_M
        dev, n1, n2 = self.dev, self.n1, self.n2
        landmark_idxs = homolog_idxs[:n_landmarks]
        uniq, gptr, members = landmark_groups([int(a) for a, _ in landmark_idxs])
        Q1 = dev.gather_cols_lower(self.W1, n1, uniq)
        B = dev.target_landmark_rows(self.W2, n2, [int(b) for _, b in landmark_idxs])
        Z, kq, verdict = dev.landmark_solve(Q1, n1, B, n2, (gptr, members))
        if verdict.fallback_needed():          # dependent landmark rows: the factored form does not hold
            return tuple(float(x) for x in self.statistic(homolog_idxs, n_landmarks)[:3].cpu())
        X = self._d1_columns(Q1, kq, dev.empty(n1, kq))
        plan = plan_separate_scores(n1, n2, landmark_idxs, homolog_idxs)
        # masks as n x 2 operands (second column zero): a = X^T m_R, b = Z m_C on the tensor cores
        mR = torch.zeros((n1, 2), dtype=torch.float64, device=dev.torch_device)
        mR[:, 0] = torch.from_numpy(1.0 - plan.rowL.astype(np.float64)).to(dev.torch_device)
        mC = torch.zeros((n2, 2), dtype=torch.float64, device=dev.torch_device)
        mC[:, 0] = torch.from_numpy(1.0 - plan.colL.astype(np.float64)).to(dev.torch_device)
        a = dev.empty(kq, 2, ld=2)
        dev.gemm(MC, MC, kq, 1, n1, 1.0, X, mR, 0.0, a, 0)            # a[l] = sum_{i not in R} X[i,l]
        b = dev.empty(kq, 2, ld=2)
        dev.gemm(KC, MC, kq, 1, n2, 1.0, Z, mC, 0.0, b, 0)            # b[l] = sum_{j not in C} Z[l,j]
        hh = dev.pair_dots(X, Z, kq, plan.h_row, plan.h_col)
        in_other = torch.from_numpy(((plan.rowL[plan.h_row] == 0) & (plan.colL[plan.h_col] == 0))
                                    .astype(np.float64)).to(dev.torch_device)
        block_sum = float((a[:, 0] * b[:, 0]).sum())
        hh_sum, hh_other = float(hh.sum()), float((hh * in_other).sum())
        hom_mean = hh_sum / plan.n_hh
        other_mean = (block_sum - hh_other) / plan.n_other
        return hom_mean, other_mean, hom_mean - other_mean

    def _d1_columns(self, Q1, kq, X):
        """X = D1[:, Ls] = W1^T Q1 (n1 x k'), contracting only k >= row of the triangular W1."""
        from .device import MC, KLO_M
        self.dev.gemm(MC, MC, self.n1, kq, self.n1, 1.0, self.W1, Q1, 0.0, X, KLO_M)
        return X


def permuted_homolog_idxs(homolog_idxs, p):
    """Permutation p of the 'indices' mode: seeded shuffle of the homolog index pairs."""
    order = np.random.default_rng(1000 + p).permutation(len(homolog_idxs))
    return [homolog_idxs[i] for i in order]


# ------------------------------------------------------------------ random graphs (mode "graphs")
def perturbed_graph(seed_graph, rng):
    """Degree-preserving random graph on the same node names (gen_random_network.py:50-65):
    configuration model on a shuffled degree sequence, nodes renamed so that every name keeps
    its degree, parallel edges and self loops dropped, then the simple 2-core.  Unlike the
    reference (unseeded np.random.shuffle) the generator is passed in, so runs are repeatable."""
    import networkx as nx
    names_degs = list(seed_graph.degree)
    order = rng.permutation(len(names_degs))
    names_degs = [names_degs[i] for i in order]
    new_G = nx.configuration_model([d for _, d in names_degs], seed=int(rng.integers(2 ** 31)))
    by_deg_new, by_deg_old = {}, {}
    for name, d in names_degs:
        by_deg_new.setdefault(d, []).append(name)
    for node, d in new_G.degree:
        by_deg_old.setdefault(d, []).append(node)
    mapping = {}
    for d, new_names in by_deg_new.items():
        shuffled = [new_names[i] for i in rng.permutation(len(new_names))]
        mapping.update(zip(by_deg_old[d], shuffled))
    return util.simple_two_core(nx.Graph(nx.relabel_nodes(new_G, mapping)), verbose=False)


def safe_perturbed_graph(seed_graph, rng, n_tries=10):
    """gen_random_network.py:67-73 — retry until the 2-core keeps every node of the seed."""
    for attempt in range(n_tries):
        G = perturbed_graph(seed_graph, rng)
        if not (set(G.nodes) ^ set(seed_graph.nodes)):
            return G, attempt + 1
    raise Exception('no random graph kept all nodes in %d tries' % n_tries)


def gen_random_network_data(seed_graph, n_tries=10, rng=None, lam=0.05, host=True):
    """gen_random_network.py:78-84 — dict(G, D, C, nodes) of one random graph (lam fixed at 0.05
    in the reference).  With host=False D and C are replaced by the device factor W."""
    rng = rng or np.random.default_rng()
    G, tries = safe_perturbed_graph(seed_graph, rng, n_tries)
    nodes = sorted(G.nodes())
    dev = default_device()
    n = len(nodes)
    W = dev.kernel_factor_edges(n, util.edge_index_arrays(G, nodes), lam)
    if not host:
        return dict(G=G, W=W, nodes=nodes), tries
    D = dev.download(dev.gram_lower(W, n, ld=n + (n & 1)), n, n)
    C = dev.download(dev.transpose_lower(W, n, ld=n + (n & 1)), n, n)
    return dict(G=G, D=D, C=C, nodes=nodes), tries


# ------------------------------------------------------------------ sharding + gather
def shard(n_units, rank, world):
    """Units of this rank: rank, rank + world, ... (round-robin)."""
    return list(range(rank, n_units, world))


def gather_results(local, n_units, rank, world, device=None):
    """All ranks' (unit -> 3 values) merged into an (n_units, 3) array on every rank.

    One all-gather of a padded (units_per_rank, 3) float64 tensor; works with NCCL (CUDA
    tensors) and gloo (CPU tensors).  Missing units (errors) are NaN rows."""
    import torch.distributed as dist
    per = (n_units + world - 1) // world
    buf = torch.full((per, 3), float('nan'), dtype=torch.float64)
    for slot, unit in enumerate(shard(n_units, rank, world)):
        if unit in local and local[unit] is not None:
            buf[slot] = torch.as_tensor(local[unit], dtype=torch.float64)
    if world == 1 or not (dist.is_available() and dist.is_initialized()):
        parts = [buf]
    else:
        if device is not None:
            buf = buf.to(device)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        parts = [p.cpu() for p in parts]
    out = np.full((n_units, 3), np.nan)
    for r, part in enumerate(parts):
        units = shard(n_units, r, world)
        out[units] = part.numpy()[:len(units)]
    return out


def run_permutation_test(source_G, target_G, homologs, n_landmarks, n_permutations, lam=0.05,
                         mode="indices", rank=0, world=1, seed=0, n_tries=50):
    """The whole experiment (permtest.py:111-187) with permutations sharded over `world` ranks.

    Returns dict(pval, effect_size, n_permutations, n_less_than, errs, real, random) on every rank.
    """
    dev = default_device()
    inp = graph_inputs(source_G, target_G, homologs, n_landmarks)
    s_pos, t_pos = inp["s_pos"], inp["t_pos"]
    homolog_idxs = [(s_pos[a], t_pos[b]) for a, b in homologs]
    n1, n2 = len(inp["source_nodes"]), len(inp["target_nodes"])
    real_pair = FactoredPair(dev, n1, inp["edges1"], n2, inp["edges2"], lam)
    real = real_pair.statistic(homolog_idxs, n_landmarks).cpu().numpy()

    pending = {}
    for p in shard(n_permutations, rank, world):
        try:
            if mode == "indices":
                pending[p] = real_pair.statistic(permuted_homolog_idxs(homolog_idxs, p), n_landmarks)
            elif mode == "graphs":
                rng = np.random.default_rng([seed, p])
                G1, _ = safe_perturbed_graph(source_G, rng, n_tries)
                G2, _ = safe_perturbed_graph(target_G, rng, n_tries)
                e1 = util.edge_index_arrays(G1, inp["source_nodes"])
                e2 = util.edge_index_arrays(G2, inp["target_nodes"])
                pair = FactoredPair(dev, n1, e1, n2, e2, lam)
                pending[p] = pair.statistic(homolog_idxs, n_landmarks)
                del pair
            else:
                raise ValueError("mode must be 'indices' or 'graphs'")
        except Exception as exc:     # permtest.py:97-99: count the failure, keep going
            print('ERROR with permutation {}: {}'.format(p, exc))
            pending[p] = None
    torch.cuda.current_stream(dev.index).synchronize()
    local = {}
    for p, t in pending.items():
        if t is None:
            local[p] = None
        else:
            hom_mean, other_mean, diff = t.cpu().numpy()[:3]
            local[p] = (diff, other_mean, hom_mean)
    table = gather_results(local, n_permutations, rank, world, device=dev.torch_device)
    ok = ~np.isnan(table[:, 0])
    diffs = table[ok, 0]
    pval, n_less, n_obs = one_tail_pval(real[2], diffs)
    return dict(pval=pval, effect_size=float(effect_size(real[2], diffs)), n_permutations=n_obs,
                n_less_than=n_less, errs=int((~ok).sum()),
                real=dict(mean_diff=float(real[2]), non_hom_mean=float(real[1]), hom_mean=float(real[0])),
                random=dict(mean_diffs=table[ok, 0], non_hom_means=table[ok, 1], hom_means=table[ok, 2]))


def run_permutation_test_from_files(source_G, target_G, homologs, n_landmarks, source_data_files,
                                    target_data_files, lam=0.05, rank=0, world=1):
    """The reference's own mode (permtest.py:111-187): permutation p is the pickled pair
    source_data_files[p] = dict(C, nodes), target_data_files[p] = dict(D, nodes) written by
    gen_random_network.py; the unit is difference_in_means_from_files (permtest.py:82-99: any
    exception is printed and counted).  The file pairs are dealt round-robin to the ranks."""
    import joblib
    from .munk import regularized_laplacian, rkhs_factor
    dev = default_device()
    source_nodes, target_nodes = sorted(source_G.nodes()), sorted(target_G.nodes())
    source_data = dict(C=rkhs_factor(regularized_laplacian(source_G, source_nodes, lam)), nodes=source_nodes)
    target_data = dict(D=regularized_laplacian(target_G, target_nodes, lam), nodes=target_nodes)
    real_diff, other_mean, hom_mean = difference_in_means(source_data, target_data, homologs, n_landmarks,
                                                          source_nodes, target_nodes)
    del source_data, target_data
    n_permutations = len(source_data_files)
    local = {}
    for p in shard(n_permutations, rank, world):
        try:
            local[p] = difference_in_means(joblib.load(source_data_files[p]), joblib.load(target_data_files[p]),
                                           homologs, n_landmarks, source_nodes, target_nodes)
        except Exception:
            print('ERROR with {}, {}'.format(source_data_files[p], target_data_files[p]))
            local[p] = None
    table = gather_results(local, n_permutations, rank, world, device=dev.torch_device)
    ok = ~np.isnan(table[:, 0])
    diffs = table[ok, 0]
    pval, n_less, n_obs = one_tail_pval(real_diff, diffs)
    return dict(pval=pval, effect_size=float(effect_size(real_diff, diffs)), n_permutations=n_obs,
                n_less_than=n_less, errs=int((~ok).sum()),
                real=dict(mean_diff=float(real_diff), non_hom_mean=float(other_mean), hom_mean=float(hom_mean)),
                random=dict(mean_diffs=table[ok, 0], non_hom_means=table[ok, 1], hom_means=table[ok, 2]))
