#!/usr/bin/env python
"""Benchmark of the MUNK hot path on B200 (driver contract: one JSON line on stdout).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C3] [--impl reference]

One "step" = the whole dense pipeline for one synthetic scale-free graph pair of the named
configuration: Laplacian assembly of both graphs, Cholesky + triangular
inverse of the source (C1 = L1^-T), Cholesky + k-column solves of the target (D2[Lt,:]), landmark
projection C2hat, and the n1 x n2 score product S = C1 C2hat^T.  At N=1 the workload is
BASELINE.json's 20k -> 16k configuration (C3); N>1 shards that one pipeline over the GPUs.

  value  = useful FP64 flops of the step (formulae of SURVEY 8d, triangular products counted once)
           x graph pairs processed by all ranks / device time (CUDA events, max over ranks), inputs
           already resident in HBM.
  e2e    = the same through the host-buffer API: edge/index arrays copied H2D every step and
           C1, C2hat and S copied back into pinned host memory inside the timed region.
  roofline     = the dominant kernel (the DMMA score GEMM launch) against the measured FP64 DMMA peak.
  parity       = after the timed regions: the whole score matrix the timed steps produced (all ranks'
                 shards at N>1) against an independent CPU computation (oracle.scores_by_cholesky).
  api_e2e      = (N=1) munk.embed_networks + munk.scores from networkx graphs to host numpy arrays.
  cpu_baseline = (N=1) the reference arm below on a bounded quarter-size sample, in a subprocess; plus
                 the full-configuration time if `--impl reference` ran on this host before.

`--impl reference` times the UNMODIFIED reference (oracle/_ref, installed by oracle/build_ref.py:
munk.embed_networks + np.dot, i.e. inv + eigh + pinv + dot via numpy/scipy) on the same configuration
with every host core, one step, without importing munk_b200.

Diagnostics (stderr only): MUNK_BENCH_TRACE=1 prints per-step phase times and device-idle gaps of
the multi-GPU run; MUNK_BENCH_CLOCK_PERIOD=<s> sets the clock sampling period.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MUNK pipeline FP64 TFLOP/s (useful flops / pipeline s), 20k->16k nodes"
UNIT = "TFLOP/s"
# DRAM bytes (read + write) of the C3 score launch, ncu --set full, profiles/r02_ncu_summary.md
NCU_SCORE_TRAFFIC_BYTES = 38.0e9


def pipeline_flops(n1, n2, k):
    """Useful flops of the build's algorithm: source chol n1^3/3 + trtri n1^3/3;
    target chol n2^3/3 + the k-column solve 2*n2^2*k for D2[Lt,:] (no full target inverse);
    Gram 2*k^2*n1, small solves, projection 2*n1*n2*k, triangular score product n1^2*n2."""
    src = 2.0 * n1 ** 3 / 3.0
    tgt = 1.0 * n2 ** 3 / 3.0
    rows = 2.0 * k * n2 ** 2
    gram = 2.0 * k * k * n1
    solve = 2.0 * k * k * n2 + 2.0 * k ** 3 / 3.0
    proj = 2.0 * n1 * n2 * k
    score = 1.0 * n1 * n1 * n2
    return dict(source=src, target=tgt, rows=rows, gram=gram, solve=solve, proj=proj, score=score,
                total=src + tgt + rows + gram + solve + proj + score)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (B200_PROFILING.md clocks line), read
    in-process through NVML every 50 ms (a spawned nvidia-smi per sample costs ~100 ms of driver
    time each and perturbs host-driven steps); nvidia-smi is the fallback if NVML is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self._halt = gpu_index, [], threading.Event()
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if visible and all(x.strip().isdigit() for x in visible.split(",")):
                phys = int(visible.split(",")[gpu_index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        nv, h = self.nvml, self.handle
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        act = lambda bit: "Active" if mask & bit else "Not Active"  # noqa: E731
        return [str(sm), str(mx), "", act(nv.nvmlClocksEventReasonHwSlowdown),
                act(nv.nvmlClocksEventReasonHwThermalSlowdown), act(nv.nvmlClocksEventReasonSwThermalSlowdown),
                act(nv.nvmlClocksEventReasonSwPowerCap)]

    def run(self):
        while not self._halt.is_set():
            try:
                if self.nvml is not None:
                    self.rows.append(self._sample_nvml())
                else:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                         timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(float(os.environ.get("MUNK_BENCH_CLOCK_PERIOD", "0.05" if self.nvml is not None else "0.2")))

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, flag in zip(names, r[3:7]):
                    if flag.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm),
                "via": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_inputs(cfg_name, rank, n1=None, n2=None):
    from munk_b200 import pipeline
    import synth as synthetic
    c = synthetic.CONFIGS[cfg_name]
    n1, n2 = n1 or c[0], n2 or c[1]
    m, H, k = c[2], c[3], c[4]
    G1, G2, homs = synthetic.make_case(n1, n2, m, min(H, n1, n2), seeds=(1 + 2 * rank, 2 + 2 * rank))
    inp = pipeline.graph_inputs(G1, G2, homs, min(k, len(homs)))
    return n1, n2, len(inp["landmark_idxs"]), inp, (G1, G2, homs)


def _settle_gc():
    """The synthetic networkx graphs (millions of Python objects) are scaffolding: drop them and
    park what is left in the permanent generation, so that a full garbage collection does not
    stall the host-driven panel loops for tens of ms in the middle of a timed step."""
    import gc
    gc.collect()
    gc.freeze()


# ------------------------------------------------------------------------------------ CPU arm
# Nothing in this section imports munk_b200 (the reference arm must not map the repo's .so).
def host_cores():
    """Cores this process may run on (cgroup/affinity aware)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def force_host_threads():
    """Every BLAS/OpenMP pool gets all host cores, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1 when N>1).  Must run before numpy is imported."""
    n = str(host_cores())
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = n


def workload_name(cfg_name, n1, n2, k):
    import synth
    return "%s: source %d / target %d nodes, %d landmarks, lam 0.05, Barabasi-Albert m=%d (seeds 1/2)" % (
        cfg_name, n1, n2, k, synth.CONFIGS[cfg_name][2])


def cpu_info():
    import numpy, scipy
    info = {"cores": host_cores(), "os_cpu_count": os.cpu_count(), "numpy": numpy.__version__,
            "scipy": scipy.__version__}
    try:
        from threadpoolctl import threadpool_info
        pools = threadpool_info()
        info["blas"] = ["%s %s x%d" % (p.get("internal_api"), p.get("version"), p.get("num_threads")) for p in pools]
        info["threads"] = max([p.get("num_threads", 1) for p in pools] + [1])
    except Exception:
        info["threads"] = host_cores()
    return info


def reference_callable():
    """(kind, fn): fn(G1, G2, homs, k) -> (seconds, runtimes dict) runs the reference's public path
    — munk.embed_networks (munk/munk/munk.py:114-176) followed by the score product
    np.dot(source_X, target_X.T) (scripts/compute_embeddings.py:81).  kind "reference" = the
    unmodified package pip-installed under oracle/_ref (oracle/build_ref.py); "port" = the oracle
    restatement, used only if oracle/_ref is absent."""
    import numpy as np
    from oracle import build_ref
    ref = build_ref.load()
    if ref is not None:
        def run_ref(G1, G2, homs, k):
            t0 = time.time()
            (source_X, _), (target_X, _), _, runtimes = ref.embed_networks(G1, G2, homs, k)
            t_embed = time.time() - t0
            S = np.dot(source_X, target_X.T)                   # compute_embeddings.py:81
            t = time.time() - t0
            runtimes = dict(runtimes, score_product=t - t_embed, shape=list(S.shape))
            return t, runtimes
        return "reference", run_ref
    from oracle import munk_oracle as oracle

    def run_port(G1, G2, homs, k):
        t0 = time.time()
        out = oracle.embed_networks(G1, G2, homs, k)
        t_embed = time.time() - t0
        S = oracle.scores(out["C1"], out["C2"])
        t = time.time() - t0
        return t, {"embed_networks": t_embed, "score_product": t - t_embed, "shape": list(S.shape)}
    return "port", run_port


def ref_cache_path(cfg_name):
    return os.path.join(ROOT, ".bench_cache", "reference_%s.json" % cfg_name)


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation on the SAME configuration as the
    product arm (C3 by default: one step of ~5-8 minutes on 16-32 host cores, so steps = 1 and
    warm-up = 0 whatever --steps/--warmup say; LAPACK has no JIT to warm).  A calibration run on a
    quarter-size sample (n^3 model) guards the driver's time limit: if the full configuration is
    projected to exceed MUNK_REF_BUDGET_S (default 700 s) the largest same-shape sample that fits is
    timed instead and the line says so (`config.same_config` false).  --ref-n1/--ref-n2 force a
    sample (the bounded cpu_baseline leg of the product arm)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import synth
    from threadpoolctl import threadpool_limits
    cores = host_cores()
    n1c, n2c, m, H, k = synth.CONFIGS[args.config]
    kind, run = reference_callable()
    assert "munk_b200" not in sys.modules
    budget = float(os.environ.get("MUNK_REF_BUDGET_S", "700"))
    calibration = None
    with threadpool_limits(limits=cores):
        info = cpu_info()
        if args.ref_n1:
            n1, n2 = min(args.ref_n1, n1c), min(args.ref_n2 or args.ref_n1, n2c)
        else:
            G1, G2, homs = synth.make_case(240, 200, m, 60, seeds=(1, 2))
            run(G1, G2, homs, 30)                       # imports, BLAS thread start-up
            q1, q2 = max(200, n1c // 4), max(150, n2c // 4)
            G1, G2, homs = synth.make_case(q1, q2, m, min(H, q1, q2), seeds=(1, 2))
            t_cal, _ = run(G1, G2, homs, min(k, len(homs)))
            est = t_cal * (n1c / q1) ** 3 * 0.75      # LAPACK runs ~25 % closer to peak at 4x the order (measured)
            calibration = {"n1": q1, "n2": q2, "seconds": t_cal, "projected_full_s": est, "budget_s": budget}
            n1, n2 = n1c, n2c
            if est > budget:
                f = (budget / est) ** (1.0 / 3.0)
                n1, n2 = max(q1, int(n1c * f) // 500 * 500), max(q2, int(n2c * f) // 500 * 500)
        G1, G2, homs = synth.make_case(n1, n2, m, min(H, n1, n2), seeds=(1, 2))
        k = min(k, len(homs))
        t, runtimes = run(G1, G2, homs, k)
        info = cpu_info()
    same = (n1, n2) == (n1c, n2c)
    fl = pipeline_flops(n1, n2, k)["total"]
    val = fl / t / 1e12
    what = ("unmodified reference (oracle/_ref: munk.embed_networks + np.dot, munk.py:114-176, "
            "compute_embeddings.py:81)" if kind == "reference" else "oracle port of the reference algorithm")
    sample = "%s on %s, %d threads: %.1f s" % (
        what, ("the full configuration" if same else "a same-shape sample") + " n1=%d n2=%d k=%d" % (n1, n2, k),
        info.get("threads", cores), t)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": 1, "warmup": 0, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, n1, n2, k), "same_config": same,
                   "parallelism": "CPU, %d BLAS threads" % info.get("threads", cores),
                   "useful_tflop_per_step": fl / 1e12, "seconds_per_pipeline": t,
                   "note": "value = the build's useful-flop count / reference seconds, so value ratios are time "
                           "ratios; the reference's own algorithm executes ~88 TFLOP at C3 (inv + eigh + dense "
                           "diag product + pinv + dot)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": info.get("threads", cores), "kind": kind,
                         "sample": sample, "host": info},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "reference_runtimes_s": runtimes, "calibration": calibration,
    }
    if same and not args.no_cache:
        try:
            import socket
            os.makedirs(os.path.dirname(ref_cache_path(args.config)), exist_ok=True)
            with open(ref_cache_path(args.config), "w") as f:
                json.dump({"seconds": t, "value": val, "kind": kind, "cores": info.get("threads", cores),
                           "host": socket.gethostname(), "when": time.time(), "n1": n1, "n2": n2, "k": k,
                           "runtimes_s": runtimes}, f)
        except OSError:
            pass
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(cfg_name, n1, n2):
    """The product arm's `cpu_baseline`: the reference arm on a bounded same-shape sample, run in a
    subprocess (its own BLAS thread settings, no munk_b200 in that process), plus — when
    `--impl reference` already ran the full configuration on this host (the driver runs it first) —
    that full-size time under `full_config`."""
    res = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", cfg_name,
                          "--ref-n1", str(n1), "--ref-n2", str(n2), "--no-cache"],
                         capture_output=True, text=True, timeout=1500,
                         env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    out = None
    for ln in res.stdout.splitlines():
        if ln.startswith("{"):
            out = json.loads(ln)
    if out is None:
        return {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port",
                "sample": "cpu baseline leg failed: %s" % res.stderr[-300:]}
    base = out["cpu_baseline"]
    base["ms_per_step"] = out["ms_per_step"]
    base["runtimes_s"] = out.get("reference_runtimes_s")
    try:
        import socket
        full = json.load(open(ref_cache_path(cfg_name)))
        if full.get("host") == socket.gethostname() and time.time() - full.get("when", 0) < 6 * 3600:
            base["full_config"] = {"seconds": full["seconds"], "value": full["value"], "kind": full["kind"],
                                   "cores": full["cores"], "runtimes_s": full.get("runtimes_s"),
                                   "source": "bench.py --impl reference on this host, %.0f s ago" % (
                                       time.time() - full["when"])}
    except (OSError, ValueError, KeyError):
        pass
    return base


def cpu_checker(n1, edges1, n2, edges2, landmark_idxs):
    """Independent CPU scores for the `parity` object: oracle.scores_by_cholesky (LAPACK Cholesky
    route, pinned to the reference's S in tests/test_oracle.py), all host cores."""
    import numpy as np
    from threadpoolctl import threadpool_limits
    from oracle import munk_oracle as oracle
    t0 = time.time()
    with threadpool_limits(limits=host_cores()):
        out = oracle.scores_by_cholesky(n1, np.stack(edges1, axis=1).astype(np.int64), n2,
                                        np.stack(edges2, axis=1).astype(np.int64), landmark_idxs)
    out["seconds"] = time.time() - t0
    return out


# ------------------------------------------------------------------------------------ GPU arm
_JSON_FD = None


def _quiet_stdout():
    """Reserve the real stdout for the one JSON line: everything else that writes to fd 1
    (NCCL's version banner, library chatter) is redirected to stderr."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def parity_vs_checker(S_host, chk, landmark_idxs, what):
    """max|S - S_cpu| / max|S_cpu| over the whole matrix (north_star tolerance 1e-8) plus the
    interpolation property S[Ls,:] = D2[Lt,:] against the CPU's D2 rows, in row chunks."""
    import numpy as np
    S_ref = chk["S"]
    scale = float(np.abs(S_ref).max())
    err = 0.0
    for r0 in range(0, S_ref.shape[0], 2048):
        err = max(err, float(np.abs(S_host[r0:r0 + 2048] - S_ref[r0:r0 + 2048]).max()))
    ls = [int(a) for a, _ in landmark_idxs]
    interp = float(np.abs(S_host[ls, :] - chk["D2_rows"]).max())
    return {"S_max_rel_err": err / scale, "S_landmark_rows_vs_cpu_D2_rows": interp / scale, "tolerance": 1e-8,
            "ok": bool(err / scale <= 1e-8 and interp / scale <= 1e-8), "cells_compared": int(S_ref.size),
            "what": what,
            "checker": "oracle.scores_by_cholesky: CPU LAPACK Cholesky route, no GPU data; pinned to the "
                       "reference's S in tests/test_oracle.py; %.1f s on %d host threads" % (chk["seconds"], host_cores())}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        _quiet_stdout()       # NCCL prints its version banner on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from munk_b200 import pipeline
    from munk_b200.device import Device, landmark_groups
    from munk_b200._lib import lib
    import ctypes

    dev = Device(local)
    if world > 1 and args.mode == "strong":
        return run_gpu_dist(args, dev, rank, world, local)
    n1, n2, k, inp, graphs = make_inputs(args.config, rank)
    graphs = None
    _settle_gc()
    fl = pipeline_flops(n1, n2, k)
    tdev = dev.torch_device

    # device-resident inputs for the `value` region
    e1 = (dev.ints(inp["edges1"][0]), dev.ints(inp["edges1"][1]))
    e2 = (dev.ints(inp["edges2"][0]), dev.ints(inp["edges2"][1]))
    s_idx = [a for a, _ in inp["landmark_idxs"]]
    t_idx = [b for _, b in inp["landmark_idxs"]]
    uniq, gptr, members = landmark_groups(s_idx)
    uniq_d, tidx_d = dev.ints(uniq), dev.ints(t_idx)

    # persistent buffers (sized for 180 GB HBM: ~14 GB at C3)
    A1, A2 = dev.empty(n1, n1), dev.empty(n2, n2)
    linv1, linv2 = dev.new_linv(n1), dev.new_linv(n2)
    S_buf = dev.empty(n1, n2)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    phase_ms = {}
    from munk_b200.device import side_device
    dev2, side, tgt_stream = side_device(dev)
    verdicts = []
    use_pair = os.environ.get("MUNK_PAIR", "1") != "0"

    def step(edges1, edges2, uniq_idx, tgt_idx, record=False):
        marks = []

        def mark(name):
            if record:
                e = ev(); e.record(); marks.append((name, e))
        mark("start")
        # Source chain on the high-priority stream, the independent target chain on a second
        # stream with its own workspace: it fills the SMs the source's latency-bound block-column
        # chain leaves idle.  Both factorisations replay cached CUDA graphs, so the host enqueues
        # each chain in well under a millisecond.
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        tgt_stream.wait_stream(main)
        if use_pair:
            # both Cholesky factorisations as ONE interleaved schedule (munk_potrf_pair), then the
            # source's triangular inverse and the target's landmark solves side by side
            with torch.cuda.stream(tgt_stream):
                dev2.assemble(n2, edges2[0], edges2[1], 0.05, out=A2)
                a2_ready = ev(); a2_ready.record()
            with torch.cuda.stream(side):
                h0 = ev(); h0.record()
                dev.assemble(n1, edges1[0], edges1[1], 0.05, out=A1)
                h1 = ev(); h1.record()
                side.wait_event(a2_ready)
                dev.potrf_pair(A1, n1, linv1, dev2, A2, n2, linv2)
                h2 = ev(); h2.record()
            # the dense triangular inverse on the normal-priority stream, the latency-bound landmark
            # solves on the high-priority one: their small kernels slot in at tile boundaries
            with torch.cuda.stream(tgt_stream):
                tgt_stream.wait_event(h2)
                dev.trtri(A1, n1, linv1)
                h3 = ev(); h3.record()
            with torch.cuda.stream(side):
                s0 = s1 = h2
                B = dev2.target_landmark_rows_from_factor(A2, linv2, n2, tgt_idx)
                s2 = ev(); s2.record()
        else:
            with torch.cuda.stream(side):
                h0 = ev(); h0.record()
                dev.assemble(n1, edges1[0], edges1[1], 0.05, out=A1)
                h1 = ev(); h1.record()
                dev.potrf(A1, n1, linv=linv1, check_info=False)
                h2 = ev(); h2.record()
                dev.trtri(A1, n1, linv1)
                h3 = ev(); h3.record()
            with torch.cuda.stream(tgt_stream):
                s0 = ev(); s0.record()
                dev2.assemble(n2, edges2[0], edges2[1], 0.05, out=A2)
                dev2.potrf(A2, n2, linv=linv2, check_info=False)
                s1 = ev(); s1.record()
                B = dev2.target_landmark_rows_from_factor(A2, linv2, n2, tgt_idx)
                s2 = ev(); s2.record()
        main.wait_stream(side)
        main.wait_stream(tgt_stream)
        B.record_stream(main)
        mark("factor_chains")
        Q1 = dev.gather_cols_lower(A1, n1, uniq_idx)
        P, verdict = dev.project(Q1, n1, B, n2, (gptr, members), defer=True); mark("embed")
        verdicts.append(verdict.bad)
        dev.score(A1, n1, P, n2, out=S_buf); mark("score")
        if record:
            torch.cuda.current_stream().synchronize()
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                phase_ms.setdefault(name, []).append(a.elapsed_time(b))
            for name, a, b in (("assemble_src", h0, h1), ("potrf_pair" if use_pair else "potrf_src", h1, h2),
                               ("trtri_src", h2, h3), ("potrf_tgt_concurrent", s0, s1), ("solve_tgt_concurrent", s1, s2)):
                if a is not b:
                    phase_ms.setdefault(name, []).append(a.elapsed_time(b))
        return P

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step(e1, e2, uniq_d, tidx_d)
    barrier()
    assert dev.potrf_info() == 0

    # ---- timed: value (inputs resident)
    sampler = ClockSampler(local); sampler.start()
    dev.stats(reset=True); dev2.stats(reset=True)
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step(e1, e2, uniq_d, tidx_d, record=True)
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches, counted_flops = dev.stats(reset=True)
    l2_, f2_ = dev2.stats(reset=True)
    launches, counted_flops = launches + l2_, counted_flops + f2_
    clocks = sampler.stop()
    assert dev.potrf_info() == 0 and dev2.potrf_info() == 0
    assert not bool(torch.stack(verdicts).any()), "the fast landmark solve was flagged unsafe"

    # ---- timed: e2e (host buffers in, pinned host results out, every step)
    h_e1 = [torch.from_numpy(x).pin_memory() for x in inp["edges1"]]
    h_e2 = [torch.from_numpy(x).pin_memory() for x in inp["edges2"]]
    # The public device API (what scripts/compute_embeddings.py calls): host edge/index arrays in,
    # C1, C2hat^T and S streamed to the pinned buffers of a HostSink while later stages compute.
    sink = pipeline.HostSink(dev, n1, n2)
    h2d = sum(t.numel() * t.element_size() for t in h_e1 + h_e2) + 8 * len(t_idx)
    d2h = sink.d2h_bytes()

    def e2e_step():
        emb = pipeline.embed_device(n1, h_e1, n2, h_e2, inp["landmark_idxs"], dev=dev,
                                    keep_target_factor=False, sink=sink)
        emb.scores_streamed(sink)
        sink.wait()
        emb.finalize()                      # factorisation flags + projection verdict (16 bytes D2H)
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    q0, q1 = ev(), ev()
    q0.record()
    for _ in range(e2e_steps):
        e2e_step()
    q1.record()
    barrier()
    e2e_ms = q0.elapsed_time(q1)

    # ---- parity (outside every timed region): the scores the timed steps left in S_buf, the host
    # copy the e2e steps produced, and D1 1 = 1 through the factor, against an independent CPU run
    parity = api = None
    if rank == 0 and world == 1 and not args.no_parity:
        import numpy as np
        chk = cpu_checker(n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"])
        parity = parity_vs_checker(dev.download(S_buf, n1, n2), chk, inp["landmark_idxs"],
                                   "S of the last device-resident timed step, whole matrix")
        e2e_par = parity_vs_checker(sink.S.numpy(), chk, inp["landmark_idxs"], "S in the e2e HostSink")
        parity["e2e_S_max_rel_err"] = e2e_par["S_max_rel_err"]
        W = torch.tril(A1[:, :n1])
        ones = torch.ones(n1, dtype=torch.float64, device=tdev)
        parity["D1_row_sums_minus_1_max"] = float((W.t().mv(W.mv(ones)) - 1.0).abs().max())
        C1h = sink.C1.numpy()
        ls_ = [int(a) for a, _ in inp["landmark_idxs"]]
        # C1 C1[Ls,:]^T = D1[:,Ls] from the host copy of the factor (64 sampled landmark columns)
        pick = ls_[:: max(1, len(ls_) // 64)]
        parity["C1_C1t_landmark_cols_max_rel_err"] = float(
            np.abs(C1h @ C1h[pick, :].T - chk["D1_cols"][:, [ls_.index(c) for c in pick]]).max()
            / np.abs(chk["D1_cols"]).max())
        parity["ok"] = bool(parity["ok"] and e2e_par["ok"] and parity["D1_row_sums_minus_1_max"] < 1e-8
                            and parity["C1_C1t_landmark_cols_max_rel_err"] < 1e-8)
        del W, ones
        # ---- the reference-signature API, from networkx graphs to host numpy arrays
        if not args.no_api:
            import munk_b200 as munk
            import synth
            c = synth.CONFIGS[args.config]
            G1, G2, homs = synth.make_case(n1, n2, c[2], min(c[3], n1, n2), seeds=(1, 2))
            del sink
            times = []
            for _ in range(2):
                ta = time.time()
                (C1a, _), (C2a, _), _, rts = munk.embed_networks(G1, G2, homs, k)
                tb = time.time()
                Sa = munk.scores(C1a, C2a)
                times.append((time.time() - ta, tb - ta))
            apar = parity_vs_checker(Sa, chk, inp["landmark_idxs"], "munk.scores(*munk.embed_networks(...))")
            best = min(times)
            api = {"seconds": best[0], "embed_networks_s": best[1], "scores_s": best[0] - best[1],
                   "value": fl["total"] / best[0] / 1e12, "unit": UNIT, "runs_s": [round(t[0], 3) for t in times],
                   "runtimes": rts, "S_max_rel_err": apar["S_max_rel_err"],
                   "call": "munk.embed_networks(G1, G2, homologs, %d) + munk.scores(C1, C2hat): networkx graphs in, "
                           "host numpy C1 / C2hat / S out (pageable memory)" % k}
            del C1a, C2a, Sa, G1, G2
        del chk

    # ---- max over ranks
    if world > 1:
        t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=tdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        ms_step = ms / args.steps
        value = fl["total"] * world / (ms_step * 1e-3) / 1e12
        e2e_val = fl["total"] * world / (e2e_ms / e2e_steps * 1e-3) / 1e12
        # FP64 ceilings measured on this box, in this run
        v = ctypes.c_double()
        lib.munk_microbench_dmma(local, 20000, ctypes.byref(v)); dmma_peak = v.value
        a = torch.randn(8192, 8192, dtype=torch.float64, device=tdev); b = torch.randn_like(a); c = torch.empty_like(a)
        torch.matmul(a, b, out=c); torch.cuda.synchronize()
        c0, c1 = ev(), ev(); c0.record(); torch.matmul(a, b, out=c); c1.record(); torch.cuda.synchronize()
        cublas_tf = 2 * 8192 ** 3 / c0.elapsed_time(c1) / 1e9
        del a, b, c
        med = {kk: statistics.median(vv) for kk, vv in phase_ms.items()}
        score_ms = med["score"]
        achieved = fl["score"] / (score_ms * 1e-3) / 1e12
        phases = {
            "assemble_src": {"ms": med["assemble_src"], "GBps": n1 * A1.stride(0) * 8 / med["assemble_src"] / 1e6},
            "trtri_src": {"ms": med["trtri_src"], "tflops": n1 ** 3 / 3 / med["trtri_src"] / 1e9},
            "factor_chains": {"ms": med["factor_chains"],
                              "tflops": (fl["source"] + fl["target"] + fl["rows"]) / med["factor_chains"] / 1e9,
                              "note": "wall time of both factor chains (assembly, Cholesky, source triangular inverse, "
                                      "target k-column solves) from the start of the step to the join"},
            "solve_tgt_concurrent": {"ms": med["solve_tgt_concurrent"],
                                     "note": "target landmark solves, concurrent with the source's triangular inverse"},
            "embed": {"ms": med["embed"], "tflops": (fl["gram"] + fl["solve"] + fl["proj"]) / med["embed"] / 1e9},
            "score": {"ms": score_ms, "tflops": achieved},
        }
        if "potrf_pair" in med:
            phases["potrf_pair"] = {"ms": med["potrf_pair"],
                                    "tflops": (n1 ** 3 + n2 ** 3) / 3 / med["potrf_pair"] / 1e9,
                                    "note": "munk_potrf_pair: source and target Cholesky, block columns interleaved"}
        else:
            phases["potrf_src"] = {"ms": med["potrf_src"], "tflops": n1 ** 3 / 3 / med["potrf_src"] / 1e9}
            phases["potrf_tgt_concurrent"] = {"ms": med["potrf_tgt_concurrent"]}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world == 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n1, n2, k),
                       "parallelism": "1 GPU" if world == 1 else
                                      "replicas: one independent graph pair per GPU (seeds 1+2r/2+2r), no collective",
                       "l2": "inputs larger than L2: %.1f GB of FP64 matrices per step vs 126 MB L2" % (
                           (n1 * n1 + n2 * n2 + 2 * n1 * n2) * 8 / 1e9),
                       "useful_tflop_per_step": fl["total"] / 1e12, "seconds_per_pipeline": ms_step * 1e-3},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "note": "edge/index arrays H2D from pinned memory; C1, C2hat^T and S D2H into pinned memory"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": achieved / dmma_peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this launch at C3 from the
                         # ncu --set full capture in profiles/ (null for other configs: not captured)
                         "traffic": NCU_SCORE_TRAFFIC_BYTES if args.config == "C3" else None,
                         "traffic_source": "profiles/r02_ncu_summary.md (ncu --set full, one launch; algorithmic "
                                           "bytes of the launch: %.2f GB)" % ((n1 * n1 / 2 + 2 * n1 * n2) * 8 / 1e9),
                         "kernel": "dgemm_dmma_tma_kernel<MC,MC> score launch: S = W1^T P, %d x %d x %d, "
                                   "k >= row only (%.2f TFLOP)" % (n1, n2, n1, fl["score"] / 1e12),
                         "peak_source": "measured in this run: FP64 DMMA issue-rate microbenchmark "
                                        "(MEASURED_PEAKS.json has no FP64 entry); cuBLAS dgemm 8192^3 = %.2f TFLOP/s; "
                                        "vendor FP64 peak 37-40 TFLOP/s" % cublas_tf,
                         "cublas_dgemm_tflops": cublas_tf,
                         "hbm_peak_gbs": peaks.get("hbm_gbs")},
            "phases": phases,
            "counted_tile_flops_per_step": counted_flops / args.steps,
        }
        line["parity"] = parity
        line["api_e2e"] = api
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_leg(args.config, max(200, n1 // 4), max(150, n2 // 4))
        else:
            line["cpu_baseline"] = None
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def property_parity(out, dev, inp, n1, n2, rank, world):
    """Parity of a configuration too large for a dense CPU reference (C5: 50 000 / 40 000 nodes;
    the dense checker alone would take minutes of an 8-GPU lease): size-independent properties,
    each evaluated with torch / cuBLAS operations on the sparse Laplacians and on random probe
    vectors - none of this repo's kernels is involved in the checks.
      factor      A1 (C1 C1^T v) = v for 4 random v: the row shards of C1 on all ranks really factor D1 = A1^-1
      target_rows A2 B^T = E_Lt: B is D2[Lt,:] (checked on rank 0 from the projection's own input)
      projection  C1[Ls,:] P = B: P = C2hat^T solves the landmark system (and lies in the row space by construction)
      interpolation S[Ls,:] = B on the ranks that own landmark rows
      product     S_r w = C1_r (P w) for 4 random w: every rank's score rows are C1_r C2hat^T
    All residuals are relative to the largest entry of the right-hand side; tolerance 1e-8."""
    import numpy as np
    import torch
    import torch.distributed as dist
    tdev = dev.torch_device
    rows = torch.as_tensor(np.asarray(out["rows"], dtype=np.int64), device=tdev)
    C1, P, S = out["C1"], out["P"][:, :n2], out["S"][:, :n2]

    def apply_A(edges, n, X, lam=0.05):
        """(I + lam L) X from the edge list (sparse, torch index_add)."""
        u = torch.as_tensor(np.asarray(edges[0], dtype=np.int64), device=tdev)
        v = torch.as_tensor(np.asarray(edges[1], dtype=np.int64), device=tdev)
        deg = torch.zeros(n, dtype=torch.float64, device=tdev)
        deg.index_add_(0, u, torch.ones_like(u, dtype=torch.float64))
        deg.index_add_(0, v, torch.ones_like(v, dtype=torch.float64))
        Y = X * (1.0 + lam * deg)[:, None]
        Y.index_add_(0, u, X[v], alpha=-lam)
        Y.index_add_(0, v, X[u], alpha=-lam)
        return Y

    g = torch.Generator(device="cpu").manual_seed(7)
    V = torch.randn(n1, 4, dtype=torch.float64, generator=g).to(tdev)
    Wv = torch.randn(n2, 4, dtype=torch.float64, generator=g).to(tdev)
    res = {}
    # factor: t = C1^T v (sum over ranks of C1_r^T v_r), D1 v = C1 t (rows gathered), A1 (D1 v) = v
    t = C1.t().contiguous() @ V[rows] if len(rows) else torch.zeros(n1, 4, dtype=torch.float64, device=tdev)
    dist.all_reduce(t)
    Dv = torch.zeros(n1, 4, dtype=torch.float64, device=tdev)
    if len(rows):
        Dv[rows] = C1 @ t
    dist.all_reduce(Dv)
    res["factor"] = float((apply_A(inp["edges1"], n1, Dv) - V).abs().max() / V.abs().max())
    # target rows, projection: replicated inputs, checked on every rank (max over ranks below)
    ls = torch.as_tensor([int(a) for a, _ in inp["landmark_idxs"]], device=tdev)
    lt = torch.as_tensor([int(b) for _, b in inp["landmark_idxs"]], device=tdev)
    # B = C1[Ls,:] P needs the landmark rows of C1: gather them from their owners
    k = len(ls)
    M = torch.zeros(k, n1, dtype=torch.float64, device=tdev)
    pos = {int(r): i for i, r in enumerate(out["rows"])}
    mine = [(i, pos[int(a)]) for i, a in enumerate(ls.tolist()) if int(a) in pos]
    if mine:
        M[torch.as_tensor([i for i, _ in mine], device=tdev)] = C1[torch.as_tensor([j for _, j in mine], device=tdev)]
    dist.all_reduce(M)
    B = M @ P                                                   # k x n2: what the projection reproduces
    E = torch.zeros(n2, k, dtype=torch.float64, device=tdev)
    E[lt, torch.arange(k, device=tdev)] = 1.0
    res["target_rows_and_projection"] = float((apply_A(inp["edges2"], n2, B.t().contiguous()) - E).abs().max())
    # interpolation on the owners of landmark rows; product on every rank
    interp = torch.zeros(1, dtype=torch.float64, device=tdev)
    if mine:
        li = torch.as_tensor([i for i, _ in mine], device=tdev)
        lj = torch.as_tensor([j for _, j in mine], device=tdev)
        interp[0] = (S[lj] - B[li]).abs().max() / B.abs().max()
    prod = torch.zeros(1, dtype=torch.float64, device=tdev)
    if len(rows):
        ref = C1 @ (P @ Wv)
        prod[0] = (S @ Wv - ref).abs().max() / ref.abs().max()
    both = torch.cat([interp, prod])
    dist.all_reduce(both, op=dist.ReduceOp.MAX)
    res["interpolation"], res["product"] = float(both[0]), float(both[1])
    worst = torch.tensor([max(res.values())], dtype=torch.float64, device=tdev)
    dist.all_reduce(worst, op=dist.ReduceOp.MAX)
    return {"kind": "properties", "residuals": res, "max_residual": float(worst[0]), "tolerance": 1e-8,
            "ok": bool(float(worst[0]) <= 1e-8),
            "what": "size-independent properties over all %d ranks (see bench.py property_parity): the dense CPU "
                    "checker is not run above 30 000 nodes" % world}


def run_gpu_dist(args, dev, rank, world, local):
    """N > 1: ONE graph pair of the configuration, the pipeline sharded over the ranks
    (munk_b200/dist.py: block-column-cyclic Cholesky with NCCL panel broadcasts, per-rank column
    solves, row-sharded score product).  Strong scaling: total work is fixed."""
    import ctypes
    import numpy as np
    import torch
    import torch.distributed as dist
    from munk_b200 import dist as mdist
    from munk_b200._lib import lib

    n1, n2, k, inp, _ = make_inputs(args.config, 0)            # same graph pair on every rank
    _ = None
    _settle_gc()
    fl = pipeline_flops(n1, n2, k)
    tdev = dev.torch_device
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    e1 = (dev.ints(inp["edges1"][0]), dev.ints(inp["edges1"][1]))
    e2 = (dev.ints(inp["edges2"][0]), dev.ints(inp["edges2"][1]))
    phase_ms = {}

    def step(edges1, edges2, record=False, sink=None):
        marks = [] if record else None
        out = mdist.dist_embed_scores(dev, n1, edges1, n2, edges2, inp["landmark_idxs"], rank, world,
                                      marks=marks, sink=sink)
        if record:
            torch.cuda.current_stream().synchronize()
            for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
                phase_ms.setdefault(name, []).append(a.elapsed_time(b))
            te = out["target_chain_events"]
            phase_ms.setdefault("target_chain_concurrent", []).append(te[0].elapsed_time(te[1]))
            step_ends.append((marks[0][1], marks[-1][1], time.perf_counter()))
        return out

    step_ends = []

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(e1, e2)
    barrier()
    sampler = ClockSampler(local); sampler.start()
    dev.stats(reset=True)
    barrier()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.steps):
        step(e1, e2, record=True)          # result dropped at once: keeping it alive would make the
        #                                    next step allocate a second set of n1 x n2 buffers
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches, counted_flops = dev.stats(reset=True)
    clocks = sampler.stop()
    print("rank %d: %.1f ms/step, phases %s" % (rank, ms / args.steps, {
        kk: round(statistics.median(vv), 1) for kk, vv in phase_ms.items()}), file=sys.stderr, flush=True)
    if os.environ.get("MUNK_BENCH_TRACE"):
        gaps = [round(t0.elapsed_time(step_ends[0][0]), 1)] + [
            round(a[1].elapsed_time(b[0]), 1) for a, b in zip(step_ends[:-1], step_ends[1:])] + [
            round(step_ends[-1][1].elapsed_time(t1), 1)]
        walls = [round((b[2] - a[2]) * 1e3, 1) for a, b in zip(step_ends[:-1], step_ends[1:])]
        print("rank %d: device-idle gaps before/between/after steps (ms) %s, host wall per step %s, phases per step %s" % (
            rank, gaps, walls, {kk: [round(x, 1) for x in vv] for kk, vv in phase_ms.items()}),
            file=sys.stderr, flush=True)

    # e2e: host edge arrays in; every rank's rows of C1, S and C2hat out to pinned memory through a
    # ShardSink (copies on a second stream, overlapping the remaining stages)
    h_e1 = [torch.from_numpy(x).pin_memory() for x in inp["edges1"]]
    h_e2 = [torch.from_numpy(x).pin_memory() for x in inp["edges2"]]
    sink = mdist.ShardSink(dev, n1, n2, rank, world)
    h2d = sum(t.numel() * t.element_size() for t in h_e1 + h_e2)
    d2h = sink.d2h_bytes()

    trace = {}

    def e2e_step():
        trace["start"] = ev(); trace["start"].record()
        d1 = [t.to(tdev, non_blocking=True) for t in h_e1]
        d2 = [t.to(tdev, non_blocking=True) for t in h_e2]
        step(d1, d2, sink=sink)
        trace["compute"] = ev(); trace["compute"].record()
        sink.wait()
        torch.cuda.current_stream().synchronize()

    e2e_step()
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    q0, q1 = ev(), ev()
    q0.record()
    for _ in range(e2e_steps):
        e2e_step()
    q1.record()
    barrier()
    e2e_ms = q0.elapsed_time(q1)
    print("rank %d e2e timeline (last step, ms from start): compute done %.1f, copies done %s" % (
        rank, trace["start"].elapsed_time(trace["compute"]),
        {kk: round(trace["start"].elapsed_time(e), 1) for kk, e in sink.done.items()}), file=sys.stderr, flush=True)

    # ---- parity (outside the timed regions): every rank's score rows — of one more device step and
    # of the host copies the e2e steps left in the ShardSink — against the CPU checker's S, which
    # rank 0 computes (no GPU data) and broadcasts
    parity = None
    if not args.no_parity and (n1 > 30000 or args.parity_properties):
        parity = property_parity(step(e1, e2), dev, inp, n1, n2, rank, world)
    elif not args.no_parity:
        out = step(e1, e2)
        rows_t = torch.as_tensor(np.asarray(out["rows"], dtype=np.int64), device=tdev)
        S_ref = torch.empty((n1, n2), dtype=torch.float64, device=tdev)
        meta = torch.zeros(3, dtype=torch.float64, device=tdev)
        chk = None
        if rank == 0:
            chk = cpu_checker(n1, inp["edges1"], n2, inp["edges2"], inp["landmark_idxs"])
            S_ref.copy_(torch.from_numpy(chk["S"]))
            meta[0], meta[1] = float(np.abs(chk["S"]).max()), chk["seconds"]
        dist.broadcast(S_ref, src=0)
        dist.broadcast(meta, src=0)
        errs = torch.zeros(3, dtype=torch.float64, device=tdev)
        if len(out["rows"]):
            mine = S_ref[rows_t]
            errs[0] = (out["S"][:, :n2] - mine).abs().max()
            errs[1] = (sink.S.to(tdev) - mine).abs().max()
            errs[2] = float(len(out["rows"]))
            del mine
        emax = errs.clone(); dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        esum = errs.clone(); dist.all_reduce(esum, op=dist.ReduceOp.SUM)
        scale = float(meta[0])
        parity = {"S_max_rel_err": float(emax[0]) / scale, "e2e_S_max_rel_err": float(emax[1]) / scale,
                  "tolerance": 1e-8, "ok": bool(float(emax[0]) / scale <= 1e-8 and float(emax[1]) / scale <= 1e-8
                                                and int(esum[2]) == n1),
                  "rows_compared": int(esum[2]), "cells_compared": int(esum[2]) * n2,
                  "what": "the row shards of S on all %d ranks (device result of one more step, and the pinned host "
                          "copies of the e2e steps), max over ranks" % world,
                  "checker": "oracle.scores_by_cholesky on rank 0: CPU LAPACK Cholesky route, no GPU data; pinned to "
                             "the reference's S in tests/test_oracle.py; %.1f s on %d host threads" % (
                                 float(meta[1]), host_cores())}
        del out, S_ref, chk

    t = torch.tensor([ms, e2e_ms, float(h2d), float(d2h), float(launches)], dtype=torch.float64, device=tdev)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms, e2e_ms = float(tmax[0]), float(tmax[1])
    if rank == 0:
        ms_step = ms / args.steps
        value = fl["total"] / (ms_step * 1e-3) / 1e12
        e2e_val = fl["total"] / (e2e_ms / e2e_steps * 1e-3) / 1e12
        v = ctypes.c_double()
        lib.munk_microbench_dmma(local, 20000, ctypes.byref(v)); dmma_peak = v.value
        med = {kk: statistics.median(vv) for kk, vv in phase_ms.items()}
        rows0 = len(mdist.ColumnCyclic(n1, world).rows_of(0))
        # rank 0's score launch: its rows x n2, k >= first row of each 128-row tile
        r0 = np.asarray(mdist.ColumnCyclic(n1, world).rows_of(0))
        score_flops0 = float(sum(2.0 * n2 * (n1 - (int(r) // 128) * 128) for r in r0))
        achieved = score_flops0 / (med["score"] * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n1, n2, k),
                       "parallelism": "ONE graph pair sharded over %d GPUs: 1 x %d block-column-cyclic Cholesky (NCCL "
                                      "panel broadcasts), per-rank column solves, row-sharded score product" % (world, world),
                       "l2": "inputs larger than L2: %.1f GB of FP64 matrices per step vs 126 MB L2" % (
                           (n1 * n1 + n2 * n2 + 2 * n1 * n2) * 8 / 1e9),
                       "useful_tflop_per_step": fl["total"] / 1e12, "seconds_per_pipeline": ms_step * 1e-3},
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(tsum[2]), "d2h_bytes_per_step": int(tsum[3]),
                    "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps,
                    "note": "edge arrays H2D on every rank; each rank's rows of C1, S and C2hat D2H into pinned memory (sum over ranks = the three full matrices)"},
            "gpu_launches": int(tsum[4]),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": dmma_peak, "unit": "TFLOP/s",
                         "frac": achieved / dmma_peak, "traffic": None,
                         "kernel": "dgemm_dmma_tma_kernel<MC,MC> score launch on rank 0: %d rows of S = X_own^T P "
                                   "(staircase k-range, %.2f TFLOP)" % (rows0, score_flops0 / 1e12),
                         "peak_source": "measured in this run on rank 0: FP64 DMMA issue-rate microbenchmark"},
            "phases": {kk: {"ms": vv} for kk, vv in med.items()},
            "parity": parity,
            "cpu_baseline": None,
        }
        emit(line)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="strong", choices=["strong", "replicas"],
                    help="N>1: shard one pipeline over the GPUs (strong) or one graph pair per GPU (replicas)")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C3")
    ap.add_argument("--impl", default="munk_b200", choices=["munk_b200", "reference"])
    ap.add_argument("--ref-n1", dest="ref_n1", type=int, default=0,
                    help="reference arm: time this same-shape sample instead of the full configuration")
    ap.add_argument("--ref-n2", dest="ref_n2", type=int, default=0)
    ap.add_argument("--no-cache", dest="no_cache", action="store_true")
    ap.add_argument("--no-cpu", dest="no_cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-parity", dest="no_parity", action="store_true", help="skip the CPU parity check")
    ap.add_argument("--no-api", dest="no_api", action="store_true", help="skip the api_e2e measurement")
    ap.add_argument("--parity-properties", dest="parity_properties", action="store_true",
                    help="N>1: use the size-independent property checks (default above 30 000 nodes) at any size")
    args = ap.parse_args()
    if args.impl == "reference":
        force_host_threads()
        run_reference(args)
    else:
        if int(os.environ.get("RANK", "0")) == 0:
            force_host_threads()          # rank 0 runs the CPU parity checker after the timed regions
        run_gpu(args)


if __name__ == "__main__":
    main()
