"""Solve time vs number of right-hand sides.  usage: mrhs_perf.py n k P tip nrhs [nrhs ...]"""
import sys
sys.path.insert(0, '.')
import torch
import spike_petsc_b200 as sp
n, k, P, tip = [int(v) for v in sys.argv[1:5]]
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE)
S.keep_original(True)
S.set_band_synthetic(n, k)
S.factor()
B = 8.0 * n * (2 * k + 1)
for R in [int(v) for v in sys.argv[5:]]:
    U = torch.rand(R, n, dtype=torch.float64, device='cuda')
    Bv = torch.empty_like(U); X = torch.empty_like(U)
    for r in range(R):
        S.mult(U[r].data_ptr(), Bv[r].data_ptr())
    best = 1e9
    for rep in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve(Bv.data_ptr(), X.data_ptr(), nrhs=R); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    err = ((X - U).norm() / U.norm()).item()
    st = S.view()['stage_ms']
    print(f"n={n} k={k} nrhs={R}: solve {best:.3f} ms = {best/R:.3f} ms/rhs, {(B + 32.0*n*R)/best/1e6/6555.2*100:.1f}% of HBM roofline (B+32NR), sweeps stage {st[3]:.3f} ms, err {err:.1e}", flush=True)
    del U, Bv, X
