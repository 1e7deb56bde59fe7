"""End-to-end drop-in check of scripts/compute_embeddings.py on a GPU: same flags, same five
output files and schemas as the reference CLI (scripts/compute_embeddings.py:17-31,72-98), scores
equal to the oracle's on the same edge lists."""
import json
import os
import subprocess
import sys

import joblib
import networkx as nx
import numpy as np
import pytest

from oracle import munk_oracle as oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_outputs_match_reference_schema_and_oracle(tmp_path):
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(400, 350, 3, 70, seeds=(71, 72), padded=False)
    # a pendant chain that the 2-core must remove, a second component, a self loop, duplicate lines
    G1.add_edges_from([("s0", "tail1"), ("tail1", "tail2"), ("iso1", "iso2"), ("iso2", "iso3"), ("iso3", "iso1")])
    synthetic.write_edgelist(G1, tmp_path / "src.txt", extras=True)
    synthetic.write_edgelist(G2, tmp_path / "tgt.txt", extras=True)
    raw_homs = homs[:10] + [("tail1", homs[0][1]), ("nope", "t5")] + homs[10:]
    synthetic.write_homologs(raw_homs, tmp_path / "homs.tsv")
    out = {k: str(tmp_path / v) for k, v in dict(so="src.pkl", to="tgt.pkl", sim="scores.pkl", lo="landmarks.tsv",
                                                 r="runtimes.json").items()}
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "compute_embeddings.py"),
           "-se", str(tmp_path / "src.txt"), "-te", str(tmp_path / "tgt.txt"), "-hf", str(tmp_path / "homs.tsv"),
           "-so", out["so"], "-to", out["to"], "-sim", out["sim"], "-lo", out["lo"], "-r", out["r"],
           "-n", "25", "-rs", "7", "--src_lam", "0.05", "--tgt_lam", "0.1"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]

    # the oracle on the same files
    S1 = oracle.simple_two_core(nx.read_edgelist(tmp_path / "src.txt", encoding="ascii"))
    S2 = oracle.simple_two_core(nx.read_edgelist(tmp_path / "tgt.txt", encoding="ascii"))
    assert "tail1" not in S1 and "iso1" not in S1
    valid = oracle.homologs_in_graphs(S1, S2, oracle.read_homolog_list(tmp_path / "homs.tsv"))
    ref = oracle.embed_networks(S1, S2, valid, 25, src_lam=0.05, tgt_lam=0.1)
    S_ref = oracle.scores(ref["C1"], ref["C2"])

    src, tgt, sim = joblib.load(out["so"]), joblib.load(out["to"]), joblib.load(out["sim"])
    assert sorted(src) == sorted(tgt) == ["X", "landmarks", "nodes", "weight"]          # munk.py:180
    assert sorted(sim) == ["A_nodes", "B_nodes", "X", "homologs", "landmarks"]          # compute_embeddings.py:82-85
    assert src["nodes"] == ref["source_nodes"] and tgt["nodes"] == ref["target_nodes"]
    assert sim["A_nodes"] == ref["source_nodes"] and sim["B_nodes"] == ref["target_nodes"]
    assert [tuple(h) for h in sim["homologs"]] == [tuple(h) for h in valid]
    assert [tuple(l) for l in sim["landmarks"]] == [tuple(l) for l in ref["landmark_homs"]]
    assert src["landmarks"] == [a for a, _ in ref["landmark_homs"]] and src["weight"] is None
    assert tgt["landmarks"] == [b for _, b in ref["landmark_homs"]]
    assert src["X"].shape == ref["C1"].shape and tgt["X"].shape == ref["C2"].shape
    assert oracle.rel_max_err(sim["X"], S_ref) < 1e-10
    assert oracle.rel_max_err(np.dot(src["X"], tgt["X"].T), S_ref) < 1e-10              # saved embeddings reproduce S
    lines = open(out["lo"]).read().splitlines()
    assert lines == ["%s\t%s" % (a, b) for a, b in ref["landmark_homs"]]
    rt = json.load(open(out["r"]))
    assert rt["n_source_nodes"] == len(ref["source_nodes"]) and rt["n_target_nodes"] == len(ref["target_nodes"])
    assert sorted(rt["runtimes"]) == ["embed_runtime", "source_regularized_laplacian_runtime",
                                      "source_rkhs_factorization_runtime", "target_regularized_laplacian_runtime"]
    assert rt["total_time"] > 0


def _oracle_real_pair(tmp_path, lam):
    S1 = oracle.simple_two_core(nx.read_edgelist(tmp_path / "src.txt", encoding="ascii"))
    S2 = oracle.simple_two_core(nx.read_edgelist(tmp_path / "tgt.txt", encoding="ascii"))
    homs = oracle.homologs_in_graphs(S1, S2, oracle.read_homolog_list(tmp_path / "homs.tsv"))
    n1, n2 = oracle.sorted_nodes(S1), oracle.sorted_nodes(S2)
    p1, p2 = oracle.node_to_index(S1), oracle.node_to_index(S2)
    idx = [(p1[a], p2[b]) for a, b in homs]
    C1 = oracle.rkhs_factor(oracle.regularized_laplacian(S1, n1, lam))
    D2 = oracle.regularized_laplacian(S2, n2, lam)
    return S1, S2, n1, n2, idx, C1, D2


def test_permtest_cli_from_pickles_matches_oracle(tmp_path):
    """scripts/permtest.py in the reference's own mode (permtest.py:111-187): random-graph pickles
    dict(C, nodes) / dict(D, nodes) in, results JSON + diffs pickle out; every number against the
    oracle's difference_in_means / one_tail_pval / effect_size on the same pickles.  One pickle pair
    is corrupt: it must be reported and counted in `errs`, not abort the run (permtest.py:97-99)."""
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(260, 220, 4, 50, seeds=(81, 82))
    synthetic.write_edgelist(G1, tmp_path / "src.txt")
    synthetic.write_edgelist(G2, tmp_path / "tgt.txt")
    synthetic.write_homologs(homs, tmp_path / "homs.tsv")
    lam, k = 0.05, 15
    S1, S2, n1, n2, idx, C1, D2 = _oracle_real_pair(tmp_path, lam)
    rng = np.random.default_rng(3)
    sf, tf, want = [], [], []
    for p in range(5):
        # any graph on the same node names is a valid "random network" for the unit under test
        R1 = nx.relabel_nodes(nx.barabasi_albert_graph(len(n1), 4, seed=100 + p), dict(enumerate(n1)))
        R2 = nx.relabel_nodes(nx.barabasi_albert_graph(len(n2), 4, seed=200 + p), dict(enumerate(n2)))
        Dr1, Dr2 = oracle.regularized_laplacian(R1, n1, lam), oracle.regularized_laplacian(R2, n2, lam)
        Cr1 = oracle.rkhs_factor(Dr1)
        sf.append(str(tmp_path / ("s%d.pkl" % p)))
        tf.append(str(tmp_path / ("t%d.pkl" % p)))
        joblib.dump(dict(G=R1, D=Dr1, C=Cr1, nodes=n1), sf[-1])        # gen_random_network.py:84 schema
        joblib.dump(dict(G=R2, D=Dr2, C=None, nodes=n2), tf[-1])
        want.append(oracle.difference_in_means(Cr1, Dr2, idx, k))
    sf.append(str(tmp_path / "broken_s.pkl")); tf.append(str(tmp_path / "broken_t.pkl"))
    open(sf[-1], "wb").write(b"not a pickle"); open(tf[-1], "wb").write(b"not a pickle")
    out, dof = str(tmp_path / "res.json"), str(tmp_path / "diffs.pkl")
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "permtest.py"), "-s", str(tmp_path / "src.txt"),
           "-t", str(tmp_path / "tgt.txt"), "-hom", str(tmp_path / "homs.tsv"), "-nl", str(k), "-lam", str(lam),
           "-sf"] + sf + ["-tf"] + tf + ["-o", out, "-dof", dof, "-j", "3"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    assert "ERRORS: 1" in res.stdout and "ERROR with" in res.stdout
    real = oracle.difference_in_means(C1, D2, idx, k)
    got, diffs = json.load(open(out)), joblib.load(dof)
    assert sorted(got) == ["effect_size", "errs", "n_less_than", "n_permutations", "pval"]      # permtest.py:176
    assert sorted(diffs) == ["random", "real"] and sorted(diffs["random"]) == ["hom_means", "mean_diffs", "non_hom_means"]
    assert np.allclose([diffs["real"]["mean_diff"], diffs["real"]["non_hom_mean"], diffs["real"]["hom_mean"]],
                       real, rtol=1e-9, atol=0)
    assert np.allclose(np.array([diffs["random"]["mean_diffs"], diffs["random"]["non_hom_means"],
                                 diffs["random"]["hom_means"]]).T, np.array(want), rtol=1e-9, atol=0)
    p, n_less, n_obs = oracle.one_tail_pval(real[0], [w[0] for w in want])
    assert got["errs"] == 1 and got["n_permutations"] == n_obs == 5 and got["n_less_than"] == n_less
    assert got["pval"] == p
    assert abs(got["effect_size"] - oracle.effect_size(real[0], [w[0] for w in want])) < 1e-6 * abs(got["effect_size"])


def test_permtest_cli_index_permutations_match_oracle(tmp_path):
    """BASELINE config 4 semantics (--mode indices): graphs fixed, permutation p shuffles the
    homolog list with default_rng(1000 + p); the unit is checked against the oracle per permutation."""
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(240, 200, 4, 60, seeds=(91, 92))
    synthetic.write_edgelist(G1, tmp_path / "src.txt")
    synthetic.write_edgelist(G2, tmp_path / "tgt.txt")
    synthetic.write_homologs(homs, tmp_path / "homs.tsv")
    lam, k, n_perm = 0.05, 12, 6
    S1, S2, n1, n2, idx, C1, D2 = _oracle_real_pair(tmp_path, lam)
    want = []
    for p in range(n_perm):
        order = np.random.default_rng(1000 + p).permutation(len(idx))
        want.append(oracle.difference_in_means(C1, D2, [idx[i] for i in order], k))
    out, dof = str(tmp_path / "res.json"), str(tmp_path / "diffs.pkl")
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "permtest.py"), "-s", str(tmp_path / "src.txt"),
           "-t", str(tmp_path / "tgt.txt"), "-hom", str(tmp_path / "homs.tsv"), "-nl", str(k), "-lam", str(lam),
           "--n_permutations", str(n_perm), "--mode", "indices", "-o", out, "-dof", dof]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    got, diffs = json.load(open(out)), joblib.load(dof)
    real = oracle.difference_in_means(C1, D2, idx, k)
    assert np.allclose(diffs["real"]["mean_diff"], real[0], rtol=1e-9, atol=0)
    assert np.allclose(diffs["random"]["mean_diffs"], [w[0] for w in want], rtol=1e-9, atol=0)
    p, n_less, n_obs = oracle.one_tail_pval(real[0], [w[0] for w in want])
    assert (got["pval"], got["n_less_than"], got["n_permutations"], got["errs"]) == (p, n_less, n_obs, 0)
