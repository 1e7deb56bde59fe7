"""Device -> host bandwidth of the result path on this box: one stream, two concurrent streams,
and the triangular C1 download (munk_download_upper) against a full copy.  Pinned host buffers,
CUDA events, warm.  Usage: python tools/d2h_probe.py [n]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from munk_b200.device import default_device  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
    dev = default_device()
    d = torch.randn(n, n, dtype=torch.float64, device=dev.torch_device)
    h = torch.empty(n, n, dtype=torch.float64, pin_memory=True)
    gb = n * n * 8 / 1e9
    ms = timed(lambda: h.copy_(d, non_blocking=True))
    print("D2H one stream   %.2f GB: %.1f ms  %.1f GB/s" % (gb, ms, gb / ms * 1e3))
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    print("H2D one stream   %.2f GB: %.1f ms  %.1f GB/s" % (gb, ms, gb / ms * 1e3))

    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    half = n // 2

    def two():
        cur = torch.cuda.current_stream()
        s1.wait_stream(cur); s2.wait_stream(cur)
        with torch.cuda.stream(s1):
            h[:half].copy_(d[:half], non_blocking=True)
        with torch.cuda.stream(s2):
            h[half:].copy_(d[half:], non_blocking=True)
        cur.wait_stream(s1); cur.wait_stream(s2)

    ms = timed(two)
    print("D2H two streams  %.2f GB: %.1f ms  %.1f GB/s" % (gb, ms, gb / ms * 1e3))

    C = torch.triu(d)
    h.zero_()
    for rb in (256, 1024, 4096):
        ms = timed(lambda: dev.download_upper(C, n, h, row_block=rb))
        moved = sum((n - r0) * min(rb, n - r0) for r0 in range(0, n, rb)) * 8 / 1e9
        print("download_upper row_block %4d: %.2f GB moved, %.1f ms  %.1f GB/s (full copy would move %.2f GB)"
              % (rb, moved, ms, moved / ms * 1e3, gb))
    assert torch.equal(h, C.cpu())
    print("download_upper result equals triu: OK")


if __name__ == "__main__":
    main()
