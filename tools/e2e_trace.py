"""Timeline of one end-to-end pipeline call (host edge arrays in, pinned host results out) at a
BASELINE.json configuration: when the host enqueued each stage and when the device finished it,
for the compute stream and the HostSink copy stream.  Usage: python tools/e2e_trace.py [C3] [reps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from munk_b200 import pipeline
import synth as synthetic  # noqa: E402
from munk_b200.device import default_device  # noqa: E402


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n1, n2, m, n_homs, k = synthetic.CONFIGS[cfg]
    G1, G2, homs = synthetic.make_case(n1, n2, m, n_homs)
    inp = pipeline.graph_inputs(G1, G2, homs, k)
    dev = default_device()
    h_e1 = [torch.from_numpy(x).pin_memory() for x in inp["edges1"]]
    h_e2 = [torch.from_numpy(x).pin_memory() for x in inp["edges2"]]
    sink = pipeline.HostSink(dev, n1, n2)
    marks = []

    def stamp(name, stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream or torch.cuda.current_stream())
        marks.append((name, time.perf_counter(), e))

    push, push_src, gemm_stair, project = sink.push, sink.push_source_factor, dev.gemm_stair, dev.project
    gather = dev.gather_cols_lower

    def traced_gather(*a, **kw):
        stamp("source chain + target chain done (compute stream)")
        return gather(*a, **kw)

    def traced_push(dst, src):
        push(dst, src)
        stamp("copy %.2f GB done" % (src.numel() * 8 / 1e9), sink.stream)

    def traced_push_src(W1, n):
        push_src(W1, n)
        stamp("copy C1 done", sink.stream)

    def traced_gemm_stair(*a, **kw):
        out = gemm_stair(*a, **kw)
        stamp("score block done")
        return out

    def traced_project(*a, **kw):
        out = project(*a, **kw)
        stamp("projection done")
        return out

    sink.push, sink.push_source_factor, dev.gemm_stair, dev.project = (traced_push, traced_push_src,
                                                                      traced_gemm_stair, traced_project)
    dev.gather_cols_lower = traced_gather
    for rep in range(reps):
        torch.cuda.synchronize()
        del marks[:]
        stamp("start")
        emb = pipeline.embed_device(n1, h_e1, n2, h_e2, inp["landmark_idxs"], dev=dev,
                                    keep_target_factor=False, sink=sink)
        stamp("embed_device returned")
        emb.scores_streamed(sink)
        stamp("scores enqueued")
        sink.wait()
        torch.cuda.current_stream().synchronize()
        t_end = time.perf_counter()
        _, h0, e0 = marks[0]
        print("--- rep %d: host wall %.1f ms" % (rep, (t_end - h0) * 1e3))
        for name, h, e in marks:
            print("  %-52s host enqueue %7.1f ms   device done %7.1f ms" % (name, (h - h0) * 1e3, e0.elapsed_time(e)))
        del emb


if __name__ == "__main__":
    main()
