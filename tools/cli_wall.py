"""Wall time of the drop-in CLI (scripts/compute_embeddings.py) on a BASELINE.json configuration:
writes the synthetic edge lists and homolog list to a scratch directory, runs the CLI as a
subprocess, and reports the process wall time, the CLI's own runtimes JSON, the log timeline and
the size of the written files (SURVEY.md section 8(d) item iii).
Usage: python tools/cli_wall.py [C1|C2|C3] [--networkx_ingest] [--keep]"""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import synth as synthetic  # noqa: E402


def main():
    cfg_name = next((a for a in sys.argv[1:] if not a.startswith("--")), "C3")
    extra = [a for a in sys.argv[1:] if a == "--networkx_ingest"]
    n1, n2, m, n_homs, k = synthetic.CONFIGS[cfg_name]
    tmp = tempfile.mkdtemp(prefix="munk_cli_")
    try:
        t0 = time.time()
        G1, G2, homs = synthetic.make_case(n1, n2, m, n_homs)
        synthetic.write_edgelist(G1, os.path.join(tmp, "src.txt"))
        synthetic.write_edgelist(G2, os.path.join(tmp, "tgt.txt"))
        synthetic.write_homologs(homs, os.path.join(tmp, "homs.tsv"))
        print("inputs written in %.1f s (%d / %d nodes, %d / %d edges)"
              % (time.time() - t0, G1.number_of_nodes(), G2.number_of_nodes(), G1.number_of_edges(),
                 G2.number_of_edges()), flush=True)
        out = {k: os.path.join(tmp, v) for k, v in dict(so="src.pkl", to="tgt.pkl", sim="scores.pkl",
                                                        lo="landmarks.tsv", r="runtimes.json").items()}
        cmd = [sys.executable, os.path.join(ROOT, "scripts", "compute_embeddings.py"),
               "-se", os.path.join(tmp, "src.txt"), "-te", os.path.join(tmp, "tgt.txt"),
               "-hf", os.path.join(tmp, "homs.tsv"), "-so", out["so"], "-to", out["to"],
               "-sim", out["sim"], "-lo", out["lo"], "-r", out["r"], "-n", str(k)] + extra
        t0 = time.time()
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
        wall = time.time() - t0
        print(res.stderr[-4000:])
        print(res.stdout[-2000:])
        assert res.returncode == 0
        sizes = {k: os.path.getsize(v) for k, v in out.items()}
        report = dict(config=cfg_name, cli_wall_s=wall, runtimes_json=json.load(open(out["r"])),
                      output_bytes=sizes, ingest="networkx" if extra else "arrays")
        print("CLI_WALL " + json.dumps(report))
    finally:
        if "--keep" not in sys.argv:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
