"""Per-phase clock64 breakdown of the LU kernel (CTA 0, steps 100..163).  usage: lu_trace.py n k P"""
import sys, ctypes as C
sys.path.insert(0, '.')
import numpy as np, torch
import spike_petsc_b200 as sp
n, k, P = [int(v) for v in sys.argv[1:4]]
S = sp.Spike(partitions=P, tip_tiles=104, mem=sp.MEM_DEVICE)
S.set_band_synthetic(n, k)
tr = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
L = sp.lib(); L.spk_debug_set_lu_trace.argtypes = [C.c_void_p, C.c_void_p]
L.spk_debug_set_lu_trace(S._h, tr.data_ptr())
S.factor(); torch.cuda.synchronize()
t = tr.cpu().numpy().reshape(64, 16)
names = ["c:S1 enter", "c:S1 exit", "c:phaseA end", "c:S2 exit", "c:update end", "c:publish end", "-", "-",
         "s:loop top", "s:D exit", "-", "s:inv done", "s:arrived"]
step = np.diff(t[:, 0]).astype(float)
print("cycles/step (col warp 0 S1-enter to S1-enter): mean %.0f min %.0f max %.0f" % (step.mean(), step.min(), step.max()))
def seg(a, b): return (t[:, b] - t[:, a]).astype(float)
print("col warp0:  wait S1 %.0f | phase A %.0f | wait S2 %.0f | update %.0f | publish %.0f" % (
    seg(0, 1).mean(), seg(1, 2).mean(), seg(2, 3).mean(), seg(3, 4).mean(), seg(4, 5).mean()))
print("service:    wait D %.0f | Gauss-Jordan inverse + publish %.0f | arrive %.0f | (arrive -> next loop top) %.0f" % (
    seg(8, 9).mean(), seg(9, 11).mean(), seg(11, 12).mean(), (t[1:, 8] - t[:-1, 12]).astype(float).mean()))
print("service arrive relative to col S1 enter: %.0f ; D exit relative to col S2 exit: %.0f" % ((t[:, 12] - t[:, 0]).astype(float).mean(), (t[1:, 9] - t[:-1, 3]).astype(float).mean()))
