#!/usr/bin/env python3
"""Does the 128-byte-wide channel strip bound the depthwise convolution?  Same kernel, same bytes: (N=128, C=512) -- eight
strips of 128 B per 1 KB row, read and written by different blocks -- against (N=1024, C=64): every row is one contiguous
128 B and consecutive rows are adjacent in memory."""
import os, sys, math
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zipvoice_b200 import _lib
lib = _lib.load()
s = torch.cuda.current_stream().cuda_stream
for K in (31, 7):
    for N, C in ((128, 512), (1024, 64), (256, 256), (512, 128)):
        L = 1219
        x = torch.randn(N, L, C, device="cuda").half()
        w = (torch.randn(K, C, device="cuda") / math.sqrt(K)).contiguous()
        b = torch.zeros(C, device="cuda")
        out = torch.empty_like(x)
        for _ in range(3):
            _lib.check(lib.zvb_test_dwconv(x.data_ptr(), out.data_ptr(), w.data_ptr(), b.data_ptr(), N, L, C, K, s))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            _lib.check(lib.zvb_test_dwconv(x.data_ptr(), out.data_ptr(), w.data_ptr(), b.data_ptr(), N, L, C, K, s))
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"K={K:2d} N={N:5d} C={C:4d}: {us:7.1f} us  {2 * x.numel() * 2 / us / 1e3:7.0f} GB/s")
