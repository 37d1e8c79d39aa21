#!/usr/bin/env python3
"""GEMM mainloop study: time zvb_test_linear over K and N at the benchmark's row count."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zipvoice_b200 import _lib  # noqa: E402


def run(M, K, N, out_mode=0, act=0, resid=False, block_n=0, reps=5):
    lib = _lib.load()
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.float16)
    W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.float16)
    b = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda").to(torch.float16) if resid else None
    out = torch.empty(M, N, dtype=torch.float16 if out_mode == 0 else torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream

    def go():
        _lib.check(lib.zvb_test_linear(A.data_ptr(), M, K, K, W.data_ptr(), b.data_ptr(), N, K, block_n, act,
                                       R.data_ptr() if resid else None, None, None, out.data_ptr(), N, out_mode, s))
    go()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        go()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    tiles_per_sm = math.ceil(M / 128) * math.ceil(N / 256) / 148
    kblocks = math.ceil(K / 64)
    us = best * 1e3
    print(f"M={M} K={K:5d} N={N:5d} mode={out_mode} act={act} resid={int(resid)}: {us:8.1f} us  "
          f"{2.0 * M * N * K / best / 1e9:7.1f} TFLOP/s   {us * 1.9e3 / tiles_per_sm / kblocks:7.0f} clk/k-block/tile "
          f"({tiles_per_sm:.1f} tiles/SM x {kblocks} k-blocks)", flush=True)


if __name__ == "__main__":
    M = 156032
    for K in (64, 512, 2048, 8192):
        run(M, K, 512)
    for N in (256, 1024, 2048):
        run(M, 512, N)
    run(M, 512, 1536, act=1)
    run(M, 1536, 512, resid=True)
    run(M, 512, 512, resid=True)
    run(8192, 8192, 8192)
