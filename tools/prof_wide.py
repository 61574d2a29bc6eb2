"""One wide-band factor (+ solve) of a synthetic band (used under ncu).  usage: prof_wide.py n k P tip nrhs"""
import sys
sys.path.insert(0, '.')
import torch
import spike_petsc_b200 as sp
n, k, P, tip, nrhs = [int(v) for v in sys.argv[1:6]]
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE)
S.set_band_synthetic(n, k)
U = torch.rand(nrhs, n, dtype=torch.float64, device='cuda'); B = torch.empty_like(U); X = torch.empty_like(U)
for r in range(nrhs):
    S.mult(U[r].data_ptr(), B[r].data_ptr())
S.factor()
S.solve(B.data_ptr(), X.data_ptr(), nrhs=nrhs)
torch.cuda.synchronize()
info = S.view()
print("err", ((X - U).norm() / U.norm()).item(), info)
