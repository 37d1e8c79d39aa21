#!/usr/bin/env python3
"""One GEMM launch (for ncu): python tools/gemm_one.py M K N [out_mode act resid]"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
from gemm_sweep import run
a = [int(x) for x in sys.argv[1:]]
M, K, N = a[:3]
out_mode, act, resid = (a[3:] + [0, 0, 0])[:3]
run(M, K, N, out_mode=out_mode, act=act, resid=bool(resid), reps=2)
