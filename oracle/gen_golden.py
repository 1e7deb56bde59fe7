"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/munk).

Run once in the build container (the GPU box has no /root/reference):
    python oracle/gen_golden.py
The reference imports `sklearn.externals.joblib`, removed from scikit-learn >= 0.23; an in-memory
alias to the stand-alone joblib is installed before the import.  No
reference file is edited or copied.  util.simplify_graph / simple_two_core do not run under
networkx 3, so inputs are made 2-cores beforehand with the oracle's restatement (BA graphs with
m >= 2 at these seeds are already their own 2-core; asserted below).
"""
import os
import sys

import joblib
import numpy as np
import sklearn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

sklearn.externals = type(sys)("sklearn.externals")
sklearn.externals.joblib = joblib
sys.modules["sklearn.externals"] = sklearn.externals
sys.modules["sklearn.externals.joblib"] = joblib
sys.path.insert(0, "/root/reference/munk")
import munk as ref  # noqa: E402  (the reference package)

import synth as synthetic  # noqa: E402
from oracle import munk_oracle as oracle  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# name: (n1, n2, BA m, n_homologs, n_landmarks, padded names, inject duplicate source landmark)
CASES = {
    "alice_bob": (100, 120, 2, 40, 20, True, False),      # the reference's own toy shape
    "small_unpadded": (150, 90, 3, 50, 25, False, False),  # pure lexicographic node order
    "dup_source_landmark": (80, 70, 3, 30, 12, True, True),
}


def run_case(name, n1, n2, m, H, k, padded, dup):
    G1, G2, homs = synthetic.make_case(n1, n2, m, H, seeds=(11, 12), padded=padded)
    G1, G2 = oracle.simple_two_core(G1), oracle.simple_two_core(G2)
    homs = ref.util.homologs_in_graphs(G1, G2, homs)
    if dup:
        # one-to-many homolog list: source landmark 0 appears twice (two different targets)
        homs = homs[:5] + [(homs[0][0], homs[7][1])] + homs[5:]
    (C1, s_nodes), (C2, t_nodes), l_idx, h_idx = ref.embed_networks(G1, G2, homs, k, return_idxs=True)
    (_, _), (_, _), l_homs, runtimes = ref.embed_networks(G1, G2, homs, k)
    D1 = ref.regularized_laplacian(G1, s_nodes, 0.05)
    D2 = ref.regularized_laplacian(G2, t_nodes, 0.05)
    S = np.dot(C1, C2.T)
    sep = ref.separate_scores(S, l_idx, h_idx)
    e1 = oracle.edge_index_arrays(G1, s_nodes)
    e2 = oracle.edge_index_arrays(G2, t_nodes)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        source_nodes=np.array(s_nodes), target_nodes=np.array(t_nodes),
        source_edges=e1.astype(np.int32), target_edges=e2.astype(np.int32),
        homologs=np.array(homs), n_landmarks=k, lam=0.05,
        landmark_idxs=np.array(l_idx, dtype=np.int64), homolog_idxs=np.array(h_idx, dtype=np.int64),
        landmark_homs=np.array(l_homs),
        D1=D1, D2=D2, C1=C1, C2=C2, S=S,
        sep_LL_diag=sep[0], sep_LL_off=sep[1], sep_L_nonL=sep[2], sep_HH=sep[3], sep_other=sep[4],
        runtime_keys=np.array(sorted(runtimes.keys())),
    )
    print("%-22s n1=%d n2=%d k=%d  max S %.4f  ||C1C1^T-D1|| %.2e" % (
        name, len(s_nodes), len(t_nodes), k, S.max(), np.abs(C1 @ C1.T - D1).max()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, spec in CASES.items():
        run_case(name, *spec)
