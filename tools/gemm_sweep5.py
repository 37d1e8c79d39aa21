#!/usr/bin/env python3
"""Residual-stream GEMMs (N = 512, fp16 residual through the aux ring): tile width and K."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep import run
M = 156032
for bn in (64, 128, 256):
    run(M, 48, 512, resid=True, block_n=bn)
for bn in (128, 256):
    run(M, 512, 512, resid=True, block_n=bn)
run(M, 48, 512, resid=False, block_n=256)
run(M, 48, 48, block_n=48)
