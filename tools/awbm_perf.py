"""AWBM on the GPU at C4 size: phase times (SPIKE_AWBM_TIMING=1) and total.  usage: awbm_perf.py [n]"""
import os, sys, time
os.environ["SPIKE_AWBM_TIMING"] = "1"
sys.path.insert(0, '.')
import numpy as np
import spike_petsc_b200 as sp
from spike_petsc_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
A, Q, R = synthetic.c4_matrix(n)
ia, ja, a = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.astype(np.float64)
S = sp.Spike()
for rep in range(3):
    t0 = time.perf_counter()
    perm, match, stats = S.awbm(ia, ja, a)
    print(f"total {1e3 * (time.perf_counter() - t0):.1f} ms  stats {stats.tolist()}  recovers scramble {bool((match == R).all())}", flush=True)
