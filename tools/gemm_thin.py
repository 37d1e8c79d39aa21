#!/usr/bin/env python3
"""Thin projections (K = 512, few output columns): how fast does the A operand stream?  HBM floor = 160 MB / 6.5 TB/s = 25 us."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep import run
M = 156032
for N in (16, 48, 64, 128, 192, 256, 272, 384):
    run(M, 512, N)
for bn in (16, 32, 64):
    run(M, 512, 64, block_n=bn)
run(M, 256, 64)
run(M, 1024, 64)
