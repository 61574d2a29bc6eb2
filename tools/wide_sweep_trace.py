"""Cycle accounting of one CTA of the wide partition sweeps (debug accumulators).  usage: wide_sweep_trace.py n k P tip nrhs"""
import sys, ctypes as C
sys.path.insert(0, '.')
import torch, spike_petsc_b200 as sp
n, k, P, tip, nrhs = [int(v) for v in sys.argv[1:6]]
L = sp.lib(); L.spk_debug_set_lu_trace.argtypes = [C.c_void_p, C.c_void_p]
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE)
S.set_band_synthetic(n, k)
U = torch.rand(nrhs, n, dtype=torch.float64, device='cuda'); B = torch.empty_like(U); X = torch.empty_like(U)
for r in range(nrhs):
    S.mult(U[r].data_ptr(), B[r].data_ptr())
S.factor(); S.solve(B.data_ptr(), X.data_ptr(), nrhs=nrhs); torch.cuda.synchronize()
tr = torch.zeros(256, dtype=torch.int64, device='cuda')
L.spk_debug_set_lu_trace(S._h, C.c_void_p(tr.data_ptr()))
S.solve(B.data_ptr(), X.data_ptr(), nrhs=nrhs); torch.cuda.synchronize()
info = S.view(); t = tr.cpu()[128:136].tolist()
steps = (info['n_padded'] // 64) // info['partitions']
print('sweeps ms', info['stage_ms'][3], 'steps per direction', steps)
names = ['fwd phase1 (hand over c, enter rows, request D^-1)', 'fwd barriers', 'fwd phase2 (D^-1 product)', 'fwd phase3 (updates)', 'bwd phase1', 'bwd barrier', 'bwd updates']
for nm, v in zip(names, t):
    print(f'{nm:55s} {v:12d} cycles  {v/steps:9.0f} per step')
