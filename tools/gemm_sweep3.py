#!/usr/bin/env python3
"""K = 512 study: what bounds the feed-forward input GEMMs (tile width, activation, K depth)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep import run
M = 156032
for bn in (64, 128, 192, 256):
    run(M, 512, 1536, act=1, block_n=bn)
run(M, 512, 1536, act=0, block_n=256)
run(M, 512, 1536, act=0, block_n=256, out_mode=1)
for K in (128, 256, 512, 1024, 2048):
    run(M, K, 1536, act=1, block_n=256)
run(M, 512, 3072, act=1, block_n=256)
run(M // 2, 512, 1536, act=1, block_n=256)
