#!/usr/bin/env python3
"""Does the A operand's access pattern bound the thin projections?  Same 160 MB of A: (M, K=512) -- every 16 KB box is 128
pieces of 128 B at a 1 KB pitch -- against (8M, K=64) -- every box is one contiguous 16 KB -- and (2M, K=256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_sweep import run
M = 156032
run(M, 512, 64)
run(8 * M, 64, 64)
run(4 * M, 128, 64)
run(2 * M, 256, 64)
run(M // 2, 1024, 64)
run(M, 512, 16)
run(8 * M, 64, 16)
