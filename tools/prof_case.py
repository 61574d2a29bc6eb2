"""One factor + solve of a synthetic band (used under ncu).  usage: prof_case.py n k P tip"""
import sys
sys.path.insert(0, '.')
import torch
import spike_petsc_b200 as sp
n, k, P, tip = [int(v) for v in sys.argv[1:5]]
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE)
S.keep_original(True)
S.set_band_synthetic(n, k)
u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u); x = torch.empty_like(u)
S.mult(u.data_ptr(), b.data_ptr())
S.factor()
S.solve(b.data_ptr(), x.data_ptr())
torch.cuda.synchronize()
info = S.view()
print("err", (x - u).abs().max().item(), info)
