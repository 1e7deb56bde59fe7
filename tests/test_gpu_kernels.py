"""Kernel-level GPU tests through the C ABI: GEMM layouts and structure flags, factorisations,
triangular solves, permutation-test unit.  torch float64 on the same device is the checker for
the linear algebra; the oracle is the checker for the permutation statistic."""
import numpy as np
import pytest

from oracle import munk_oracle as oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from munk_b200.device import default_device
    return default_device()


def relerr(x, y):
    return float((x - y).abs().max() / y.abs().max())


def spd(n, seed):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.randn(n, n, dtype=torch.float64, generator=g).cuda()
    return X @ X.t() / n + 2.0 * torch.eye(n, dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("M,N,K", [(128, 128, 16), (200, 150, 100), (130, 70, 33), (1000, 900, 1100),
                                   (64, 64, 2000), (400, 400, 5000), (17, 1, 5), (2000, 300, 640)])
def test_gemm_all_layouts(dev, M, N, K):
    """C = alpha A B^T + beta C for both storages of either operand, ragged sizes, padded ld,
    including shapes that take the split-K path."""
    import torch
    from munk_b200.device import KC, MC
    torch.manual_seed(M + N + K)
    for al in (KC, MC):
        for bl in (KC, MC):
            A = torch.randn(M, K, dtype=torch.float64, device="cuda")
            B = torch.randn(N, K, dtype=torch.float64, device="cuda")
            Ab = dev.empty(M, K) if al == KC else dev.empty(K, M)
            Bb = dev.empty(N, K) if bl == KC else dev.empty(K, N)
            Ab.zero_(); Bb.zero_()
            (Ab[:, :K] if al == KC else Ab[:, :M]).copy_(A if al == KC else A.t())
            (Bb[:, :K] if bl == KC else Bb[:, :N]).copy_(B if bl == KC else B.t())
            C0 = torch.randn(M, N + (N & 1) + 2, dtype=torch.float64, device="cuda")
            for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (0.5, 2.0)):
                C = C0.clone()
                dev.gemm(al, bl, M, N, K, alpha, Ab, Bb, beta, C)
                ref = alpha * (A @ B.t()) + beta * C0[:, :N]
                assert relerr(C[:, :N], ref) < 1e-12
                assert torch.equal(C[:, N:], C0[:, N:])          # padding untouched


def test_gemm_structure_flags_never_read_stale_blocks(dev):
    """Triangular k-ranges are clipped per 128-tile: NaNs planted in the blocks above the diagonal
    must never reach a result."""
    import torch
    from munk_b200.device import KC, MC, KLO_M, KLO_N, KHI_M, KHI_N
    n = 700
    torch.manual_seed(1)
    W = torch.tril(torch.randn(n, n, dtype=torch.float64, device="cuda"))
    blk = torch.arange(n, device="cuda") // 128
    Wg = W.clone()
    Wg[blk[:, None] < blk[None, :]] = float("nan")
    D = dev.gram_lower(Wg, n, ld=n)
    assert relerr(D, W.t() @ W) < 1e-13 and torch.equal(D, D.t())
    P = torch.randn(n, 500, dtype=torch.float64, device="cuda")
    X = torch.randn(300, n, dtype=torch.float64, device="cuda")
    out = dev.empty(n, 500, ld=500)
    dev.gemm(MC, MC, n, 500, n, 1.0, Wg, P, 0.0, out, KLO_M)
    assert relerr(out, W.t() @ P) < 1e-13
    dev.gemm(KC, MC, n, 500, n, 1.0, Wg, P, 0.0, out, KHI_M)
    assert relerr(out, W @ P) < 1e-13
    out2 = dev.empty(300, n, ld=n)
    dev.gemm(KC, MC, 300, n, n, 1.0, X, Wg, 0.0, out2, KLO_N)
    assert relerr(out2, X @ W) < 1e-13
    dev.gemm(KC, KC, 300, n, n, 1.0, X, Wg, 0.0, out2, KHI_N)
    assert relerr(out2, X @ W.t()) < 1e-13


@pytest.mark.parametrize("n", [1, 64, 128, 129, 300, 1000, 1300, 2500, 4000])
def test_potrf_trtri_trsm(dev, n):
    """Cholesky (recursive below 1024, look-ahead driver above), triangular inverse, W^T W and the
    left solves against torch on the same matrix."""
    import torch
    A0 = spd(n, n)
    A = dev.empty(n, n)
    A[:, :n] = A0
    linv = dev.potrf(A, n)
    L = torch.tril(A[:, :n])
    Lref = torch.linalg.cholesky(A0)
    assert relerr(L, Lref) < 1e-12
    rhs = torch.randn(n, 37, dtype=torch.float64, device="cuda")
    B = dev.empty(n, 37)
    B[:, :37] = rhs
    dev.trsm_left(A, n, linv, B, 37, transposed=False)
    assert relerr(Lref @ B[:, :37], rhs) < 1e-11
    B[:, :37] = rhs
    dev.trsm_left(A, n, linv, B, 37, transposed=True)
    assert relerr(Lref.t() @ B[:, :37], rhs) < 1e-11
    dev.trtri(A, n, linv)
    W = torch.tril(A[:, :n])
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert relerr(W @ Lref, eye) < 1e-11
    D = dev.gram_lower(A, n)
    assert relerr(D[:, :n] @ A0, eye) < 1e-11


def test_potrf_reports_first_bad_pivot(dev):
    import torch
    from munk_b200.device import NotPositiveDefinite
    A = dev.empty(300, 300)
    A.zero_()
    A[:, :300] = torch.eye(300, dtype=torch.float64, device="cuda")
    A[200, 200] = -1.0
    with pytest.raises(NotPositiveDefinite) as err:
        dev.potrf(A, 300)
    assert "pivot 200" in str(err.value)
    assert dev.potrf_info() == 0                                   # flag is cleared once read


def test_target_rows_from_factor_equals_inverse_rows(dev):
    import networkx as nx
    import torch
    from munk_b200 import util
    n = 900
    G = nx.barabasi_albert_graph(n, 6, seed=9)
    u, v = util.edge_index_arrays(G, list(range(n)))
    idx = [5, 880, 0, 5, 431]
    A = dev.assemble(n, u, v, 0.05)
    linv = dev.potrf(A, n)
    B = dev.target_landmark_rows_from_factor(A, linv, n, idx)
    D = np.linalg.inv(oracle.reglap_matrix(n, oracle.edge_index_arrays(G, list(range(n))), 0.05))
    assert oracle.rel_max_err(B[:len(idx), :n].cpu().numpy(), D[idx, :]) < 1e-12
    W = dev.trtri(A, n, linv)
    B2 = dev.target_landmark_rows(W, n, idx)
    assert oracle.rel_max_err(B2[:len(idx), :n].cpu().numpy(), D[idx, :]) < 1e-12


def test_permutation_unit_and_experiment(dev):
    """difference_in_means (permtest.py:68-80) on host arrays vs the oracle, and the sharded
    experiment in 'indices' mode vs the oracle run permutation by permutation."""
    from munk_b200 import permtest
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(260, 220, 4, 60, seeds=(31, 32))
    k = 15
    ref = oracle.embed_networks(G1, G2, homs, k)
    src = dict(C=ref["C1"], nodes=ref["source_nodes"])
    tgt = dict(D=ref["D2"], nodes=ref["target_nodes"])
    got = permtest.difference_in_means(src, tgt, homs, k, ref["source_nodes"], ref["target_nodes"])
    want = oracle.difference_in_means(ref["C1"], ref["D2"], ref["homolog_idxs"], k)
    assert np.allclose(got, want, rtol=1e-10, atol=0)
    res = permtest.run_permutation_test(G1, G2, homs, k, n_permutations=6, mode="indices")
    assert res["errs"] == 0 and res["n_permutations"] == 6.0
    assert abs(res["real"]["mean_diff"] - want[0]) <= 1e-10 * abs(want[0])
    diffs = []
    for p in range(6):
        perm = permtest.permuted_homolog_idxs(ref["homolog_idxs"], p)
        diffs.append(oracle.difference_in_means(ref["C1"], ref["D2"], perm, k)[0])
    assert np.allclose(res["random"]["mean_diffs"], diffs, rtol=1e-9, atol=0)
    pval, n_less, n_obs = oracle.one_tail_pval(want[0], np.array(diffs))
    assert (res["pval"], res["n_less_than"]) == (pval, n_less)
    assert abs(res["effect_size"] - oracle.effect_size(want[0], np.array(diffs))) < 1e-6 * abs(res["effect_size"])


def test_streamed_outputs_equal_direct_outputs(dev):
    """HostSink path (results copied out on a second stream while later stages run, scores in
    row blocks) gives the same arrays as the plain serial device path (to rounding: the row-block
    launches pick different split-K factors, i.e. a different summation order)."""
    from munk_b200 import pipeline
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(1500, 900, 4, 80, seeds=(51, 52))
    res = pipeline.embed_and_score(G1, G2, homs, 30)
    inp = pipeline.graph_inputs(G1, G2, homs, 30)
    emb = pipeline.embed_device(1500, inp["edges1"], 900, inp["edges2"], inp["landmark_idxs"], overlap=False)
    assert oracle.rel_max_err(res["scores"], emb.scores_host()) < 1e-13
    assert oracle.rel_max_err(res["source_X"], emb.source_C()) < 1e-13
    assert oracle.rel_max_err(res["target_X"], emb.target_C_hat()) < 1e-13
    assert res["target_X"].shape == (900, 1500) and res["target_X"].flags.f_contiguous
    ref = oracle.embed_networks(G1, G2, homs, 30)
    assert oracle.rel_max_err(res["scores"], oracle.scores(ref["C1"], ref["C2"])) < 1e-11


@pytest.mark.parametrize("n,row_block,pinned", [(1, 4, False), (37, 8, True), (300, 128, True),
                                                (1029, 1024, False)])
def test_download_upper_moves_exactly_the_triangle(dev, n, row_block, pinned):
    """munk_download_upper: bit-exact upper triangle on the host, strictly lower part untouched."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(n)
    full = torch.randn(n, n + 3, dtype=torch.float64, generator=g)
    C = torch.triu(full[:, :n]).contiguous()
    Cd = torch.zeros(n, n + 3, dtype=torch.float64, device=dev.torch_device)
    Cd[:, :n] = C.to(dev.torch_device)
    host = torch.full((n, n + 1), -7.0, dtype=torch.float64, pin_memory=pinned)
    dev.download_upper(Cd, n, host, row_block=row_block)
    torch.cuda.synchronize()
    got = host[:, :n]
    assert torch.equal(torch.triu(got), C)
    blk = torch.arange(n) // row_block * row_block               # first column each row receives
    untouched = torch.arange(n)[None, :] < blk[:, None]
    assert bool((got[untouched] == -7.0).all()) and bool((host[:, n] == -7.0).all())
    assert bool((got[~untouched & (torch.arange(n)[None, :] < torch.arange(n)[:, None])] == 0).all())


def test_distributed_driver_single_rank_matches_oracle(dev):
    """munk_b200.dist with world = 1 (same code path as N ranks, minus the broadcasts): panel
    driver, staircase solve for the W columns + landmark columns, staircase score product, means."""
    from munk_b200 import dist as mdist, pipeline
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(1300, 700, 4, 90, seeds=(41, 42))
    homs = homs[:7] + [(homs[3][0], homs[50][1])] + homs[7:]            # one repeated source landmark
    k = 40
    ref = oracle.embed_networks(G1, G2, homs, k)
    S_ref = oracle.scores(ref["C1"], ref["C2"])
    inp = pipeline.graph_inputs(G1, G2, homs, k)
    out = mdist.dist_embed_scores(dev, 1300, inp["edges1"], 700, inp["edges2"], inp["landmark_idxs"],
                                  rank=0, world=1, homolog_idxs=ref["homolog_idxs"])
    assert out["rows"] == list(range(1300))
    assert oracle.rel_max_err(out["S"][:, :700].cpu().numpy(), S_ref) < 1e-10
    C1 = out["C1"].cpu().numpy()
    assert oracle.rel_max_err(C1 @ C1.T, ref["D1"]) < 1e-11
    want = oracle.difference_in_means(ref["C1"], ref["D2"], ref["homolog_idxs"], k)
    got = out["stats"]
    assert np.allclose([got[2], got[1], got[0]], want, rtol=1e-9, atol=0)


def test_staircase_solve_matches_dense(dev):
    import torch
    from munk_b200.dist import staircase_rhs
    n = 1100
    A0 = spd(n, 3)
    A = dev.empty(n, n)
    A[:, :n] = A0
    linv = dev.potrf(A, n)
    own = list(range(512, 900))
    lms = [0, 5, 300, 777, 1099]
    X, stair, n_own_pad = staircase_rhs(n, own, lms, dev.torch_device)
    dev.trsm_left(A, n, linv, X, X.shape[1], transposed=False, stair=stair)
    W = torch.linalg.inv(torch.linalg.cholesky(A0))
    assert relerr(X[:, :len(own)], W[:, own]) < 1e-11
    assert relerr(X[:, n_own_pad:n_own_pad + len(lms)], W[:, lms]) < 1e-11
    assert float(X[:, len(own):n_own_pad].abs().max()) == 0.0


def test_rank_deficient_landmark_rows_follow_pinv(dev):
    """Landmark rows that are linearly dependent without being duplicates: np.linalg.pinv
    truncates the null direction; the Gram/Cholesky fast path cannot, so the device-side condition
    bound must send the projection to the robust path (munk_pinv_rows_solve), which gives the
    same minimum-norm answer."""
    import munk_b200
    rng = np.random.default_rng(11)
    C = rng.standard_normal((60, 60))
    C[41] = C[3] + 0.5 * C[17]                                   # dependent, not a duplicate
    D2 = rng.standard_normal((50, 50))
    idx = [(3, 1), (17, 4), (41, 7), (8, 9), (22, 30), (3, 12)]   # plus a true duplicate (3 twice)
    _, want = oracle.embed_matrices(C, D2, idx)
    _, got = munk_b200.embed_matrices(C, D2, idx)
    assert dev.last_rank == 4                                     # 5 distinct rows, one dependent
    assert oracle.rel_max_err(got, want) < 1e-9
    # well-conditioned rows take the fast path and agree as well
    idx2 = [(3, 1), (17, 4), (8, 9), (22, 30)]
    _, want2 = oracle.embed_matrices(C, D2, idx2)
    _, got2 = munk_b200.embed_matrices(C, D2, idx2)
    assert dev.last_rank == 4 and oracle.rel_max_err(got2, want2) < 1e-12


@pytest.mark.parametrize("smallest", [1e-6, 1e-9, 1e-12, 0.0])
def test_pinv_rows_follow_numpy_cutoff(dev, smallest):
    """munk.py:109 semantics: singular values <= 1e-15 sigma_max are dropped, everything above is
    kept - also in the band (1e-12 ... 1e-9) sigma_max that a Gram-matrix method (cut ~1e-8) cannot
    resolve.  Rows M = U diag(s) V^T with prescribed singular values.  Checked: the retained rank
    and the singular values against LAPACK; the projection against np.linalg.pinv(M) @ B to 1e-8
    where that product is itself determined to 1e-8 (its conditioning is eps / smallest: below
    smallest ~ 1e-8 two correct SVD codes differ by more, so there the comparison tolerance is
    100 eps / smallest and the backward error |M P - B| carries the check)."""
    rng = np.random.default_rng(5)
    k, m, n = 24, 700, 90
    U, _ = np.linalg.qr(rng.standard_normal((k, k)))
    V, _ = np.linalg.qr(rng.standard_normal((m, k)))
    s = np.geomspace(1.0, 1e-3, k)
    s[-1] = smallest
    M = (U * s) @ V.T
    # right-hand sides in the range of M, as D2[Lt,:] is for kernel factors: B = M X
    B = M @ rng.standard_normal((m, n))
    want = np.linalg.pinv(M) @ B
    ref = np.linalg.svd(M, compute_uv=False)
    want_rank = int((ref > 1e-15 * ref[0]).sum())
    assert want_rank == (k if smallest > 0 else k - 1)
    Q1 = dev.upload(np.ascontiguousarray(M.T))
    Bd = dev.upload(B)
    P = dev.project(Q1, m, Bd, n, (list(range(k + 1)), list(range(k))))
    got = P[:m, :n].cpu().numpy()
    assert dev.last_rank == want_rank
    eps = np.finfo(float).eps
    tol = 1e-8 if smallest == 0 else max(1e-8, 100 * eps / smallest)
    assert oracle.rel_max_err(got, want) < tol
    assert np.abs(M @ got - B).max() < 1e-11 * np.abs(B).max()           # backward error
    assert np.linalg.norm(got) <= np.linalg.norm(want) * (1 + 1e-6)      # minimum norm
    # the kernel's singular values against LAPACK's (absolute accuracy eps * sigma_max, as LAPACK's own)
    R = dev.upload(M)
    _, rank, sigma = dev.pinv_rows_solve(R, k, m, Bd, n)
    assert rank == want_rank
    assert np.abs(np.sort(sigma)[::-1] - ref).max() < 100 * eps * ref[0]


def test_rank_k_statistic_matches_dense_path(dev):
    """Opt-in fast path (no S formed) vs the dense unit, incl. a repeated source landmark and
    homolog cells inside landmark rows/columns."""
    from munk_b200 import permtest, pipeline
    import synth as synthetic
    G1, G2, homs = synthetic.make_case(900, 700, 4, 120, seeds=(61, 62))
    homs = homs[:5] + [(homs[2][0], homs[40][1]), (homs[50][0], homs[3][1])] + homs[5:]
    k = 25
    inp = pipeline.graph_inputs(G1, G2, homs, k)
    hidx = [(inp["s_pos"][a], inp["t_pos"][b]) for a, b in homs]
    pair = permtest.FactoredPair(dev, 900, inp["edges1"], 700, inp["edges2"])
    for p in (None, 0, 1):
        idx = hidx if p is None else permtest.permuted_homolog_idxs(hidx, p)
        dense = pair.statistic(idx, k).cpu().numpy()[:3]
        fast = pair.statistic_rank_k(idx, k)
        assert np.allclose(fast, dense, rtol=1e-9, atol=0), (fast, dense)


def test_random_network_generator(dev):
    """gen_random_network_data (gen_random_network.py:78-84): same nodes and degrees as the seed
    graph, D and C consistent with the oracle on the generated graph."""
    import networkx as nx
    from munk_b200 import permtest
    import synth as synthetic
    seed_G = oracle.simple_two_core(synthetic.ba_graph(150, 4, 3, "g"))
    data, tries = permtest.gen_random_network_data(seed_G, n_tries=50, rng=np.random.default_rng(5))
    G = data["G"]
    assert sorted(G.nodes()) == sorted(seed_G.nodes()) == data["nodes"] and tries >= 1
    assert min(d for _, d in G.degree()) >= 2
    D_ref = oracle.regularized_laplacian(G, data["nodes"], 0.05)
    assert oracle.rel_max_err(data["D"], D_ref) < 1e-11
    assert oracle.rel_max_err(data["C"] @ data["C"].T, D_ref) < 1e-11
    res = permtest.run_permutation_test(seed_G, seed_G, [(n, n) for n in data["nodes"][:40]], 10,
                                        n_permutations=3, mode="graphs", seed=1)
    assert res["errs"] == 0 and len(res["random"]["mean_diffs"]) == 3


@pytest.mark.parametrize("n1,n2,seed", [(37, 40, 0), (64, 52, 1), (50, 30, 2), (300, 2052, 3), (9, 4, 4),
                                        (120, 4096, 5)])
def test_separate_scores_random_cells_exact(dev, n1, n2, seed):
    """separate_scores against the oracle on random index lists: repeated landmarks, one-to-many
    homologs, homolog cells inside landmark columns and several per four-column chunk, row lengths
    that do and do not take the vectorised path.  Exact, in order."""
    import munk_b200 as m
    rng = np.random.default_rng(seed)
    k = int(rng.integers(1, min(n1, n2, 12) + 1))
    ls, lt = rng.choice(n1, k, replace=True), rng.choice(n2, k, replace=True)
    extra = int(rng.integers(k, 6 * k + 8))
    hs = np.concatenate([ls, rng.choice(n1, extra)])
    ht = np.concatenate([lt, rng.choice(max(1, n2 // 3), extra)])          # crowded columns
    l_idx, h_idx = list(zip(ls.tolist(), lt.tolist())), list(zip(hs.tolist(), ht.tolist()))
    S = rng.random((n1, n2))
    for got, want in zip(m.separate_scores(S, l_idx, h_idx), oracle.separate_scores(S, l_idx, h_idx)):
        assert np.array_equal(got, want)


def test_weighted_and_multigraph_assembly_bit_exact(dev):
    """Weights / parallel edges follow nx.laplacian_matrix (munk.py:82): A == np.eye + lam * L bit
    for bit, and regularized_laplacian on such a graph matches the reference formula."""
    import networkx as nx
    from munk_b200 import util
    import munk_b200 as munk
    rng = np.random.default_rng(4)
    G = nx.barabasi_albert_graph(300, 4, seed=2)
    for a, b in G.edges():
        G[a][b]["weight"] = float(rng.integers(1, 9)) / 4.0
    M = nx.MultiGraph(nx.barabasi_albert_graph(200, 3, seed=3))
    M.add_edges_from(list(M.edges())[:50])                      # 50 doubled edges
    for H, lam in ((G, 0.05), (M, 0.13)):
        nodes = sorted(H.nodes())
        n = len(nodes)
        edges = util.edge_index_arrays(H, nodes)
        assert len(edges) == 4
        A = dev.assemble_edges(n, edges, lam)
        L = np.array(nx.laplacian_matrix(H, nodelist=nodes).todense())
        want = np.eye(n) + lam * L
        assert np.array_equal(A[:, :n].cpu().numpy(), want)
        D = munk.regularized_laplacian(H, nodes, lam)
        assert oracle.rel_max_err(D, np.linalg.inv(want)) < 1e-12


def test_assembly_skips_and_reports_out_of_range_edges(dev):
    import torch
    n = 64
    u = np.array([0, 1, 70, 3, -2], dtype=np.int32)
    v = np.array([1, 2, 3, 64, 5], dtype=np.int32)
    guard = torch.full((n + 8, 80), 7.0, dtype=torch.float64, device=dev.torch_device)
    A = guard[:n]
    dev.assemble(n, u, v, 0.05, out=A)
    assert dev.assemble_info() == 3                                # first bad edge is index 2
    assert dev.assemble_info() == 0                                # flag cleared
    assert bool((guard[n:] == 7.0).all())                          # nothing written past the matrix
    want = oracle.reglap_matrix(n, np.array([[0, 1], [1, 2]]), 0.05)
    assert np.array_equal(A[:, :n].cpu().numpy(), want)


@pytest.mark.parametrize("n", [1300, 700, 2049])
def test_dist_factor_panel_major_single_rank(dev, n):
    """csrc/dist.cu with one rank (the N-rank program minus the broadcasts): assembly into
    panel-major storage is bit-exact, the factor matches LAPACK, the fused forward substitution
    and the backward substitution give A^-1 E."""
    import networkx as nx
    import torch
    from munk_b200 import util
    from munk_b200.dist import DistFactor, onehot_rhs, staircase_rhs
    G = nx.barabasi_albert_graph(n, 5, seed=n)
    u, v = util.edge_index_arrays(G, list(range(n)))
    want_A = oracle.reglap_matrix(n, oracle.edge_index_arrays(G, list(range(n))), 0.05)
    fac = DistFactor(dev, None, n)
    fac.assemble((u, v), 0.05)
    got_A = fac.unpack()[:, :n].cpu().numpy()
    assert np.array_equal(got_A, np.tril(want_A))
    own = list(range(512, min(n, 1024)))
    lms = sorted({0, 3, n // 2, n - 1})
    X, stair, n_own_pad = staircase_rhs(n, own, lms, dev.torch_device)
    fac.set_rhs(X, X.shape[1], stair)
    fac.factor()
    assert dev.potrf_info() == 0
    L = fac.unpack()[:, :n]
    L_ref = torch.linalg.cholesky(torch.from_numpy(want_A).to(dev.torch_device))
    assert relerr(L, L_ref) < 1e-13
    W = torch.linalg.inv(L_ref)
    assert relerr(X[:, :len(own)], W[:, own]) < 1e-11
    assert relerr(X[:, n_own_pad:n_own_pad + len(lms)], W[:, lms]) < 1e-11
    # A^-1 E for unsorted, repeated columns: forward (fused, second factorisation) + backward
    cols = [n - 2, 5, 5, n // 3]
    fac.assemble((u, v), 0.05)
    X2, stair2 = onehot_rhs(n, cols, dev.torch_device)
    fac.set_rhs(X2, X2.shape[1], stair2)
    fac.factor()
    fac.solve_backward(X2, X2.shape[1])
    D = torch.linalg.inv(torch.from_numpy(want_A).to(dev.torch_device))
    assert relerr(X2[:, :len(cols)], D[:, cols]) < 1e-11


@pytest.mark.parametrize("m,s", [(1, 128), (37, 512), (300, 384), (1000, 300), (2500, 512)])
def test_trsm_right_strip_and_recursive(dev, m, s):
    """munk_trsm_right: B <- B L^-T.  Up to 1024 rows it is the one-launch strip solve
    (trsm_strip.cu, also with a ragged last block), beyond that the recursion; both against torch."""
    import torch
    A0 = spd(s, 7)
    A = dev.empty(s, s)
    A[:, :s] = A0
    linv = dev.potrf(A, s)
    L = torch.tril(A[:, :s])
    B0 = torch.randn(m, s, dtype=torch.float64, device="cuda")
    guard = torch.full((m + 4, s + 16), 3.0, dtype=torch.float64, device="cuda")
    B = guard[:m]
    B[:, :s] = B0
    dev.trsm_right(B, m, A, s, linv)
    want = torch.linalg.solve_triangular(L, B0.t(), upper=False).t()
    assert relerr(B[:, :s], want) < 1e-12
    assert bool((guard[m:] == 3.0).all()) and bool((guard[:m, s:] == 3.0).all())


def test_graph_replay_survives_workspace_growth():
    """Factorisations of >= 4096 columns replay cached CUDA graphs from their second use on
    (factor.cu graph_or_direct).  A larger matrix on the same workspace moves the triangular-inverse
    temporary: the cached graphs must be dropped, not replayed with a dangling address."""
    import torch
    from munk_b200.device import Device
    d = Device(0)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        def run(A0, A, linv, n):
            A[:, :n] = A0
            d.potrf(A, n, linv=linv, check_info=False)
            d.trtri(A, n, linv)
            W = torch.tril(A[:, :n])
            x = torch.ones(n, 3, dtype=torch.float64, device="cuda")
            return float((W.t() @ (W @ (A0 @ x)) - x).abs().max())       # W^T W = A0^-1
        small, big = 4224, 5000
        As, Ab = spd(small, 1), spd(big, 2)
        bufs = d.empty(small, small), d.new_linv(small)
        bufb = d.empty(big, big), d.new_linv(big)
        errs = [run(As, *bufs, small) for _ in range(3)]                # direct, capture, replay
        errs += [run(Ab, *bufb, big) for _ in range(3)]                 # grows the temporary
        errs += [run(As, *bufs, small) for _ in range(3)]               # re-capture, replay
        assert d.potrf_info() == 0
    assert max(errs) < 1e-10, errs
