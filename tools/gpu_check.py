"""Development check of the CUDA primitives on a B200 (run under gpurun).

Compares every kernel with torch float64 on the same device, then times the big cases and the
FP64 ceilings (cuBLAS dgemm via torch.matmul, raw DMMA / DFMA issue rate).  Not part of the
product; the parity tests proper live in tests/.
"""
import ctypes
import json
import sys
import os
import time
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from munk_b200._lib import lib, check

dev = torch.device("cuda:0")
results = {}


def ws_create():
    h = ctypes.c_void_p()
    check(lib.munk_workspace_create(0, ctypes.byref(h)), "ws_create")
    return h


WS = ws_create()


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def dgemm(al, bl, M, N, K, alpha, A, B, beta, C, flags=0):
    check(lib.munk_dgemm(WS, al, bl, M, N, K, alpha, ptr(A), A.stride(0), ptr(B), B.stride(0), beta,
                         ptr(C), C.stride(0), flags, stream()), "dgemm")


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def relerr(x, y):
    return float((x - y).abs().max() / y.abs().max().clamp_min(1e-300))


def section(name):
    def deco(fn):
        try:
            t = time.time()
            fn()
            print("[ok] %s (%.1fs)" % (name, time.time() - t), flush=True)
        except Exception:
            print("[FAIL] %s" % name, flush=True)
            traceback.print_exc()
            results.setdefault("failures", []).append(name)
        return fn
    return deco


@section("microbench")
def _():
    v = ctypes.c_double()
    check(lib.munk_microbench_dmma(0, 20000, ctypes.byref(v)), "dmma")
    results["dmma_peak_tflops"] = v.value
    check(lib.munk_microbench_dfma(0, 20000, ctypes.byref(v)), "dfma")
    results["dfma_peak_tflops"] = v.value
    print("DMMA peak %.2f TF, DFMA peak %.2f TF" % (results["dmma_peak_tflops"], results["dfma_peak_tflops"]))
    a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    best, med = timeit(lambda: torch.matmul(a, b, out=c), reps=5)
    results["cublas_dgemm_8192_tflops"] = 2 * 8192 ** 3 / best / 1e9
    print("cuBLAS dgemm 8192^3: best %.2f ms -> %.2f TF (median %.2f ms)" % (best, results["cublas_dgemm_8192_tflops"], med))


def make_op(rows, K, layout, pad=0):
    """Operand with logical shape rows x K in the requested storage; returns (tensor, logical)."""
    if layout == 0:
        ld = (K + pad + 1) // 2 * 2
        buf = torch.randn(rows, ld, dtype=torch.float64, device=dev)
        return buf, buf[:, :K]
    ld = (rows + pad + 1) // 2 * 2
    buf = torch.randn(K, ld, dtype=torch.float64, device=dev)
    return buf, buf[:, :rows].t()


@section("gemm correctness")
def _():
    torch.manual_seed(0)
    cases = [(128, 128, 16), (128, 128, 256), (256, 384, 512), (200, 150, 100), (130, 70, 33),
             (1000, 900, 1100), (64, 64, 2000), (400, 400, 20000), (17, 1, 5)]
    worst = 0.0
    for (M, N, K) in cases:
        for al in (0, 1):
            for bl in (0, 1):
                Ab, A = make_op(M, K, al, pad=6)
                Bb, B = make_op(N, K, bl, pad=10)
                ldc = (N + 5) // 2 * 2
                C0 = torch.randn(M, ldc, dtype=torch.float64, device=dev)
                for (alpha, beta) in ((1.0, 0.0), (-1.0, 1.0), (0.5, 2.0)):
                    C = C0.clone()
                    dgemm(al, bl, M, N, K, alpha, Ab, Bb, beta, C)
                    ref = alpha * (A @ B.t()) + beta * C0[:, :N]
                    e = relerr(C[:, :N], ref)
                    pad_ok = torch.equal(C[:, N:], C0[:, N:])
                    worst = max(worst, e)
                    if e > 1e-12 or not pad_ok:
                        raise AssertionError("gemm mismatch M%d N%d K%d al%d bl%d a%g b%g err %g pad_ok %s"
                                             % (M, N, K, al, bl, alpha, beta, e, pad_ok))
    results["gemm_worst_relerr"] = worst
    print("gemm worst rel err %.3e" % worst)


@section("gemm structure flags")
def _():
    torch.manual_seed(1)
    n = 700
    W = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev))
    Wg = W.clone()
    # garbage above the 128-aligned diagonal blocks must never be read
    blk = (torch.arange(n, device=dev) // 128)
    mask_upper_blocks = blk[:, None] < blk[None, :]
    Wg[mask_upper_blocks] = float("nan")
    D = torch.empty(n, n, dtype=torch.float64, device=dev)
    check(lib.munk_gram_lower(WS, n, ptr(Wg), n, ptr(D), n, stream()), "gram")
    ref = W.t() @ W
    e = relerr(D, ref)
    assert e < 1e-13, "gram_lower err %g" % e
    assert torch.equal(D, D.t()), "gram_lower result not symmetric"
    # S = W^T P with KLO_M
    P = torch.randn(n, 500, dtype=torch.float64, device=dev)
    S = torch.empty(n, 500, dtype=torch.float64, device=dev)
    dgemm(1, 1, n, 500, n, 1.0, Wg, P, 0.0, S, flags=2)
    e = relerr(S, W.t() @ P)
    assert e < 1e-13, "KLO_M err %g" % e
    # KHI_M: A lower (KC) times B
    Y = torch.empty(n, 500, dtype=torch.float64, device=dev)
    dgemm(0, 1, n, 500, n, 1.0, Wg, P, 0.0, Y, flags=8)
    e = relerr(Y, W @ P)
    assert e < 1e-13, "KHI_M err %g" % e
    # KLO_N: T = X * W (W lower as B operand, MC)
    X = torch.randn(300, n, dtype=torch.float64, device=dev)
    T = torch.empty(300, n, dtype=torch.float64, device=dev)
    dgemm(0, 1, 300, n, n, 1.0, X, Wg, 0.0, T, flags=4)
    e = relerr(T, X @ W)
    assert e < 1e-13, "KLO_N err %g" % e
    # KHI_N: X * W^T (W lower as B operand, KC)
    dgemm(0, 0, 300, n, n, 1.0, X, Wg, 0.0, T, flags=16)
    e = relerr(T, X @ W.t())
    assert e < 1e-13, "KHI_N err %g" % e
    # TILE_LOWER syrk
    C0 = torch.randn(n, n, dtype=torch.float64, device=dev)
    C = C0.clone()
    dgemm(0, 0, n, n, 300, -1.0, X.t().contiguous(), X.t().contiguous(), 1.0, C, flags=1)
    ref = C0 - X.t() @ X
    lowmask = blk[:, None] >= blk[None, :]
    e = relerr(torch.where(lowmask, C, torch.zeros_like(C)), torch.where(lowmask, ref, torch.zeros_like(C)))
    assert e < 1e-13, "TILE_LOWER err %g" % e
    assert torch.equal(C[~lowmask], C0[~lowmask]), "TILE_LOWER touched upper tiles"


def spd(n, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    X = torch.randn(n, n, dtype=torch.float64, generator=g).to(dev)
    return X @ X.t() / n + torch.eye(n, dtype=torch.float64, device=dev) * 2.0


def potrf_trtri(A):
    n = A.shape[0]
    ld = A.stride(0)
    linv = torch.empty(lib.munk_linv_bytes(n) // 8, dtype=torch.float64, device=dev)
    check(lib.munk_potrf(WS, n, ptr(A), ld, ptr(linv), stream()), "potrf")
    return linv


@section("potrf/trtri correctness")
def _():
    for n in (64, 128, 129, 300, 1000, 2000, 2500):
        A0 = spd(n, n)
        ld = (n + 15) // 16 * 16
        A = torch.zeros(n, ld, dtype=torch.float64, device=dev)
        A[:, :n] = A0
        linv = potrf_trtri(A)
        info = ctypes.c_int(-1)
        check(lib.munk_potrf_info(WS, ctypes.byref(info), stream()), "info")
        assert info.value == 0, "info %d at n=%d" % (info.value, n)
        L = torch.tril(A[:, :n])
        Lref = torch.linalg.cholesky(A0)
        e = relerr(L, Lref)
        assert e < 1e-12, "potrf err %g at n=%d" % (e, n)
        check(lib.munk_trtri(WS, n, ptr(A), ld, ptr(linv), stream()), "trtri")
        W = torch.tril(A[:, :n])
        e2 = relerr(W @ Lref, torch.eye(n, dtype=torch.float64, device=dev))
        assert e2 < 1e-11, "trtri err %g at n=%d" % (e2, n)
        D = torch.empty(n, ld, dtype=torch.float64, device=dev)
        check(lib.munk_gram_lower(WS, n, ptr(A), ld, ptr(D), ld, stream()), "gram")
        e3 = relerr(D[:, :n] @ A0, torch.eye(n, dtype=torch.float64, device=dev))
        assert e3 < 1e-11, "inverse err %g at n=%d" % (e3, n)
        print("  n=%d potrf %.2e trtri %.2e inv %.2e" % (n, e, e2, e3), flush=True)
    # not-SPD detection
    A = -torch.eye(300, 304, dtype=torch.float64, device=dev)
    potrf_trtri(A)
    info = ctypes.c_int(-1)
    check(lib.munk_potrf_info(WS, ctypes.byref(info), stream()), "info")
    assert info.value == 1, "expected info=1 for -I, got %d" % info.value


@section("assemble / gather / transpose")
def _():
    import numpy as np
    import networkx as nx
    n = 777
    G = nx.barabasi_albert_graph(n, 5, seed=3)
    e = np.array(G.edges(), dtype=np.int32)
    u = torch.from_numpy(np.ascontiguousarray(e[:, 0])).to(dev)
    v = torch.from_numpy(np.ascontiguousarray(e[:, 1])).to(dev)
    lam = 0.05
    ld = (n + 15) // 16 * 16
    A = torch.full((n, ld), 7.0, dtype=torch.float64, device=dev)
    check(lib.munk_assemble_reglap(WS, n, len(e), ptr(u), ptr(v), lam, ptr(A), ld, stream()), "assemble")
    L = np.array(nx.laplacian_matrix(G, nodelist=list(range(n))).todense())
    ref = np.eye(n) + lam * L
    got = A.cpu().numpy()
    assert np.array_equal(got[:, :n], ref), "assembly not bit-exact"
    assert np.all(got[:, n:] == 0)
    W = torch.tril(torch.randn(n, n, dtype=torch.float64, device=dev))
    idx = torch.tensor([5, 700, 0, 5, 776], dtype=torch.int32, device=dev)
    Q = torch.full((n, 16), 3.0, dtype=torch.float64, device=dev)
    check(lib.munk_gather_cols_lower(n, ptr(W), n, 5, ptr(idx), ptr(Q), 16, stream()), "gather")
    assert torch.equal(Q[:, :5], W[:, idx.long()]) and float(Q[:, 5:].abs().max()) == 0.0
    Wd = W.clone()
    C = torch.empty(n, n, dtype=torch.float64, device=dev)
    check(lib.munk_transpose_lower(n, ptr(Wd), n, ptr(C), n, stream()), "transpose")
    assert torch.equal(C, W.t())


@section("timing")
def _():
    out = {}
    for n in ((4096, 8192, 16384) if not os.environ.get("QUICK") else ()):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        c = torch.empty(n, n, dtype=torch.float64, device=dev)
        for (al, bl, name) in ((0, 0, "NT"), (0, 1, "NN"), (1, 1, "TN")):
            best, med = timeit(lambda: dgemm(al, bl, n, n, n, 1.0, a, b, 0.0, c), reps=3, warm=1)
            out["dgemm_%s_%d" % (name, n)] = 2 * n ** 3 / best / 1e9
            print("  dgemm %s n=%d: %.2f ms  %.2f TF" % (name, n, best, out["dgemm_%s_%d" % (name, n)]), flush=True)
        best, med = timeit(lambda: torch.matmul(a, b, out=c), reps=3, warm=1)
        out["cublas_%d" % n] = 2 * n ** 3 / best / 1e9
        print("  cublas n=%d: %.2f ms %.2f TF" % (n, best, out["cublas_%d" % n]), flush=True)
        del a, b, c
    for n in (128, 256, 512, 1024, 2048, 4096, 8192, 16000, 20000):
        A0 = spd(n, 1)
        A = torch.empty_like(A0)
        linv = torch.empty(lib.munk_linv_bytes(n) // 8, dtype=torch.float64, device=dev)

        def run_potrf():
            A.copy_(A0)
            check(lib.munk_potrf(WS, n, ptr(A), n, ptr(linv), stream()), "potrf")
        tcopy, _m = timeit(lambda: A.copy_(A0), reps=3, warm=1)
        best, med = timeit(run_potrf, reps=3, warm=1)
        t = best - tcopy
        out["potrf_%d_ms" % n] = t
        out["potrf_%d_tflops" % n] = n ** 3 / 3 / t / 1e9
        print("  potrf n=%d: %.2f ms %.2f TF" % (n, t, out["potrf_%d_tflops" % n]), flush=True)
        L0 = A.clone()

        def run_trtri():
            A.copy_(L0)
            check(lib.munk_trtri(WS, n, ptr(A), n, ptr(linv), stream()), "trtri")
        best, med = timeit(run_trtri, reps=3, warm=1)
        t = best - tcopy
        out["trtri_%d_ms" % n] = t
        out["trtri_%d_tflops" % n] = n ** 3 / 3 / t / 1e9
        print("  trtri n=%d: %.2f ms %.2f TF" % (n, t, out["trtri_%d_tflops" % n]), flush=True)
        D = A0  # reuse
        best, med = timeit(lambda: check(lib.munk_gram_lower(WS, n, ptr(A), n, ptr(D), n, stream()), "gram"), reps=3, warm=1)
        out["gram_%d_ms" % n] = best
        out["gram_%d_tflops" % n] = n ** 3 / 3 / best / 1e9
        print("  gram n=%d: %.2f ms %.2f TF" % (n, best, out["gram_%d_tflops" % n]), flush=True)
        t0 = time.time()
        torch.linalg.cholesky(spd(n, 2))
        torch.cuda.synchronize()
        A1 = spd(n, 2)
        best, med = timeit(lambda: torch.linalg.cholesky(A1), reps=2, warm=1)
        out["cusolver_potrf_%d_ms" % n] = best
        print("  cusolver potrf n=%d: %.2f ms" % (n, best), flush=True)
        del A0, A, L0, D, A1, linv
    results["timing"] = out


os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gpu_check.json", "w") as f:
    json.dump(results, f, indent=1)
print(json.dumps(results))
sys.exit(1 if results.get("failures") else 0)
