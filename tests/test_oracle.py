"""Pins the CPU oracle (oracle/munk_oracle.py) to outputs of the unmodified reference
(tests/golden/*.npz, made by oracle/gen_golden.py from /root/reference)."""
import numpy as np

from conftest import graph_from_golden, pairs
from oracle import munk_oracle as oracle


def test_kernels_match_reference(golden):
    for which, key in (("source", "D1"), ("target", "D2")):
        G = graph_from_golden(golden, which)
        nodes = [str(x) for x in golden[which + "_nodes"]]
        assert oracle.sorted_nodes(G) == nodes
        D = oracle.regularized_laplacian(G, nodes, float(golden["lam"]))
        assert np.array_equal(D, golden[key])          # same LAPACK call on the same matrix


def test_assembly_matches_networkx(golden):
    import networkx as nx
    G = graph_from_golden(golden, "source")
    nodes = [str(x) for x in golden["source_nodes"]]
    L = np.array(nx.laplacian_matrix(G, nodelist=nodes).todense())
    e = oracle.edge_index_arrays(G, nodes)
    assert np.array_equal(oracle.laplacian_dense(len(nodes), e), L)
    assert np.array_equal(e, golden["source_edges"])


def test_rkhs_and_embedding_match_reference(golden):
    C1 = oracle.rkhs_factor(golden["D1"])
    assert np.array_equal(C1, golden["C1"])
    _, C2 = oracle.embed_matrices(C1, golden["D2"], pairs(golden["landmark_idxs"]))
    assert np.array_equal(C2, golden["C2"])
    assert np.array_equal(oracle.scores(C1, C2), golden["S"])


def test_embed_networks_indices_match_reference(golden):
    G1, G2 = graph_from_golden(golden, "source"), graph_from_golden(golden, "target")
    homs = [tuple(str(x) for x in h) for h in golden["homologs"]]
    out = oracle.embed_networks(G1, G2, homs, int(golden["n_landmarks"]))
    assert out["landmark_idxs"] == pairs(golden["landmark_idxs"])
    assert out["homolog_idxs"] == pairs(golden["homolog_idxs"])
    assert [list(h) for h in out["landmark_homs"]] == [list(h) for h in golden["landmark_homs"]]
    assert oracle.rel_max_err(oracle.scores(out["C1"], out["C2"]), golden["S"]) < 1e-12


def test_separate_scores_matches_reference(golden):
    sep = oracle.separate_scores(golden["S"], pairs(golden["landmark_idxs"]), pairs(golden["homolog_idxs"]))
    for got, key in zip(sep, ("sep_LL_diag", "sep_LL_off", "sep_L_nonL", "sep_HH", "sep_other")):
        assert np.array_equal(got, golden[key]), key


def test_known_answer_properties(golden):
    D1, D2, S = golden["D1"], golden["D2"], golden["S"]
    n1 = D1.shape[0]
    # D (I + lam L) = I, row sums 1, symmetric positive
    A = oracle.reglap_matrix(n1, golden["source_edges"].astype(np.int64), float(golden["lam"]))
    assert np.abs(D1 @ A - np.eye(n1)).max() < 1e-12
    assert np.abs(D1.sum(axis=1) - 1).max() < 1e-12
    assert (D1 > 0).all() and np.abs(D1 - D1.T).max() < 1e-15
    # interpolation property S[Ls,:] = D2[Lt,:] for distinct source landmarks (item 6)
    l = golden["landmark_idxs"]
    if golden["name"] != "dup_source_landmark":
        assert np.abs(S[l[:, 0], :] - D2[l[:, 1], :]).max() < 1e-12
        # rank-k identity (item 7)
        Ls, Lt = l[:, 0], l[:, 1]
        S2 = D1[:, Ls] @ np.linalg.solve(D1[np.ix_(Ls, Ls)], D2[Lt, :])
        assert oracle.rel_max_err(S2, S) < 1e-11


def test_complete_graph_closed_form():
    import networkx as nx
    n, lam = 17, 0.05
    G = nx.relabel_nodes(nx.complete_graph(n), {i: "v%02d" % i for i in range(n)})
    D = oracle.regularized_laplacian(G, oracle.sorted_nodes(G), lam)
    # (I + lam (nI - J))^-1 = (I + lam J) / (1 + lam n)
    assert np.abs(D - (np.eye(n) + lam * np.ones((n, n))) / (1 + lam * n)).max() < 1e-14


def test_duplicate_source_landmark_folding():
    """pinv with a repeated source row == de-duplicated rows with averaged right-hand sides
    (the identity the GPU projection relies on)."""
    rng = np.random.default_rng(3)
    C = rng.standard_normal((30, 30))
    D2 = rng.standard_normal((25, 25))
    idx = [(3, 1), (7, 2), (3, 9), (11, 4), (7, 5), (3, 6)]
    _, ref = oracle.embed_matrices(C, D2, idx)
    from munk_b200.device import landmark_groups
    uniq, gptr, members = landmark_groups([s for s, _ in idx])
    assert uniq == [3, 7, 11] and gptr == [0, 3, 5, 6] and members == [0, 2, 5, 1, 4, 3]
    B = D2[[t for _, t in idx], :]
    Bm = np.stack([B[members[gptr[g]:gptr[g + 1]]].mean(axis=0) for g in range(len(uniq))])
    M = C[uniq, :]
    folded = (M.T @ np.linalg.solve(M @ M.T, Bm)).T
    assert oracle.rel_max_err(folded, ref) < 1e-10


def test_permutation_statistics():
    d = np.array([0.1, 0.5, -0.2, 0.3])
    p, n_sig, n = oracle.one_tail_pval(0.25, d)
    assert (p, n_sig, n) == (0.5, 2.0, 4.0)
    assert abs(oracle.effect_size(0.25, d) - (0.25 - d.mean()) / d.std()) < 1e-15


def test_cholesky_checker_matches_reference_scores(golden):
    """oracle.scores_by_cholesky (the full-size checker used at BASELINE configs 2/3/5 and inside
    bench.py's `parity` object) against the reference's own S."""
    if golden["name"] == "dup_source_landmark":
        return                                         # the checker requires distinct source landmarks
    n1, n2 = len(golden["source_nodes"]), len(golden["target_nodes"])
    e1, e2 = golden["source_edges"].astype(np.int64), golden["target_edges"].astype(np.int64)
    assert np.array_equal(oracle.spd_matrix(n1, e1, float(golden["lam"])),
                          oracle.reglap_matrix(n1, e1, float(golden["lam"])))
    l = golden["landmark_idxs"]
    out = oracle.scores_by_cholesky(n1, e1, n2, e2, pairs(l))
    assert oracle.rel_max_err(out["S"], golden["S"]) < 1e-12
    assert oracle.rel_max_err(out["D1_cols"], golden["D1"][:, l[:, 0]]) < 1e-12
    assert oracle.rel_max_err(out["D2_rows"], golden["D2"][l[:, 1], :]) < 1e-12


def test_installed_reference_agrees_with_oracle():
    """oracle/_ref (the unmodified reference, pip-installed by oracle/build_ref.py; what
    `bench.py --impl reference` times) gives bit-for-bit the oracle's numbers."""
    import subprocess
    import sys
    import os
    import pytest
    from oracle import build_ref
    if not build_ref.installed():
        pytest.skip("oracle/_ref not installed (run python oracle/build_ref.py where /root/reference exists)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # a fresh interpreter: this process already has the drop-in `munk` alias importable
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle import build_ref, munk_oracle as oracle\n"
        "import synth\n"
        "ref = build_ref.load()\n"
        "assert 'munk_b200' not in sys.modules\n"
        "G1, G2, homs = synth.make_case(90, 70, 3, 25, seeds=(3, 4))\n"
        "(C1, sn), (C2, tn), lh, rt = ref.embed_networks(G1, G2, homs, 10)\n"
        "S = np.dot(C1, C2.T)\n"
        "o = oracle.embed_networks(G1, G2, homs, 10)\n"
        "assert sn == o['source_nodes'] and tn == o['target_nodes'] and lh == o['landmark_homs']\n"
        "assert np.array_equal(C1, o['C1']) and np.array_equal(C2, o['C2'])\n"
        "assert np.array_equal(S, oracle.scores(o['C1'], o['C2']))\n"
        "assert sorted(rt) == ['embed_runtime', 'source_regularized_laplacian_runtime',"
        " 'source_rkhs_factorization_runtime', 'target_regularized_laplacian_runtime']\n"
        "print('ok')\n" % root)
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert res.returncode == 0 and "ok" in res.stdout, res.stderr
