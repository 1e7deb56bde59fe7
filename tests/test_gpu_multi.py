"""Multi-GPU parity: the distributed pipeline on 2 ranks (NCCL) against the single-GPU pipeline.
Skipped unless at least two GPUs are visible (the single-rank form of the same code path is
covered by tests/test_gpu_kernels.py::test_distributed_driver_single_rank_matches_oracle and the
schedule itself by the gloo tests in tests/test_dist_cpu.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29541",
           os.path.join(ROOT, "tools", "dist_check.py"), "2600", "1900", "80"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert "-> OK" in res.stdout


def test_cli_under_torchrun_matches_single_gpu_cli(tmp_path):
    """scripts/compute_embeddings.py launched with torchrun on 2 GPUs writes the same five files as
    the plain single-GPU launch (scores to rounding, index data exactly)."""
    import joblib
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(1300, 900, 4, 90, seeds=(5, 6))
    synthetic.write_edgelist(G1, tmp_path / "src.txt")
    synthetic.write_edgelist(G2, tmp_path / "tgt.txt")
    synthetic.write_homologs(homs, tmp_path / "homs.tsv")
    script = os.path.join(ROOT, "scripts", "compute_embeddings.py")

    def run(tag, launcher):
        out = {k: str(tmp_path / ("%s_%s" % (tag, v))) for k, v in dict(
            so="src.pkl", to="tgt.pkl", sim="scores.pkl", lo="landmarks.tsv", r="runtimes.json").items()}
        cmd = launcher + [script, "-se", str(tmp_path / "src.txt"), "-te", str(tmp_path / "tgt.txt"),
                          "-hf", str(tmp_path / "homs.tsv"), "-so", out["so"], "-to", out["to"],
                          "-sim", out["sim"], "-lo", out["lo"], "-r", out["r"], "-n", "40"]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
        return out

    one = run("one", [sys.executable])
    two = run("two", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                      "--master-addr", "127.0.0.1", "--master-port", "29543"])
    a, b = joblib.load(one["sim"]), joblib.load(two["sim"])
    assert a["A_nodes"] == b["A_nodes"] and a["B_nodes"] == b["B_nodes"]
    assert [tuple(x) for x in a["landmarks"]] == [tuple(x) for x in b["landmarks"]]
    assert [tuple(x) for x in a["homologs"]] == [tuple(x) for x in b["homologs"]]
    scale = np.abs(a["X"]).max()
    assert np.abs(a["X"] - b["X"]).max() < 1e-12 * scale
    sa, sb = joblib.load(one["so"]), joblib.load(two["so"])
    ta, tb = joblib.load(one["to"]), joblib.load(two["to"])
    assert sa["nodes"] == sb["nodes"] and ta["nodes"] == tb["nodes"] and sa["landmarks"] == sb["landmarks"]
    assert sb["X"].shape == (1300, 1300) and tb["X"].shape == (900, 1300)
    assert np.abs(sb["X"] @ tb["X"].T - a["X"]).max() < 1e-11 * scale          # saved embeddings reproduce S
    assert np.abs(sa["X"] - sb["X"]).max() < 1e-11 * np.abs(sa["X"]).max()     # same factor (both are L^-T)
    assert open(one["lo"]).read() == open(two["lo"]).read()
    import json
    rt = json.load(open(two["r"]))
    assert sorted(rt["runtimes"]) == ["embed_runtime", "source_regularized_laplacian_runtime",
                                      "source_rkhs_factorization_runtime", "target_regularized_laplacian_runtime"]
