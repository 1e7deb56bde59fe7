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
