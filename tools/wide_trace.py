"""Where the wide LU's CTAs of partition group 0 spend their cycles (clock64 accumulators, debug build of the kernel).
usage: wide_trace.py n k P tip"""
import sys, ctypes as C
sys.path.insert(0, '.')
import torch, spike_petsc_b200 as sp
n, k, P, tip = [int(v) for v in sys.argv[1:5]]
L = sp.lib(); L.spk_debug_set_lu_trace.argtypes = [C.c_void_p, C.c_void_p]
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE)
S.set_band_synthetic(n, k)
tr = torch.zeros(16 * 8, dtype=torch.int64, device='cuda')
S.factor(); torch.cuda.synchronize()          # warm
S.set_band_synthetic(n, k)
L.spk_debug_set_lu_trace(S._h, C.c_void_p(tr.data_ptr()))
S.factor(); torch.cuda.synchronize()
info = S.view(); t = tr.cpu().view(16, 8)
kb = info['k_padded'] // 64
print(info['stage_ms'][:3], 'partitions', info['partitions'])
print("column CTA r: [0] pre-U (loads, Bt, barrier) [1] wait dinv [2] U gemm+stores+barrier [3] wait L row 1 [4] C rows [5] blocked on L flags [7] between columns")
for r in range(kb):
    print('col', r, [int(v) for v in t[r]])
print("inverter: [0] wait D ready [1] load [2] Gauss-Jordan [3] store+publish")
print('inv', [int(v) for v in t[kb]])
