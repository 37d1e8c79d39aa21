import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
from gemm_sweep import run
M=156032
for N in (1152,1536,1920):
    run(M,512,N,act=1)
run(M,512,512); run(M,512,512,resid=True)
run(M,384,512,resid=True)
run(M,512,272); run(M,512,384)
