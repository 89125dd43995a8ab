"""One shape of the TN GEMM (fpt_gemm.cuh) -- target for ncu captures: python tools/run_gemm.py M N K [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fermi_jl_b200 as fb
M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
eng = fb.Engine(0)
print(M, N, K, eng.gemm_bench(M, N, K, reps), "TFLOP/s")
