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
step = np.diff(t[:, 0]).astype(float)
print("cycles/step (col warp 0 S1-enter to S1-enter): mean %.0f min %.0f max %.0f" % (step.mean(), step.min(), step.max()))
def seg(a, b): return (t[:, b] - t[:, a]).astype(float)
print("col warp0:  wait full %.0f | Ub = X U %.0f | update %.0f | publish/load (1 step in KT) %.0f | loop top %.0f" % (
    seg(0, 1).mean(), seg(1, 2).mean(), seg(2, 4).mean(), seg(4, 5).mean(), (t[1:, 0] - t[:-1, 5]).astype(float).mean()))
print("inverter:   Jacobi start %.0f | Newton-Schulz %.0f" % (seg(10, 14).mean(), seg(14, 13).mean()))
print("inverter:   wait for D_s %.0f | - %.0f | inverse %.0f | publish %.0f | arrive+store %.0f | iteration %.0f" % (
    seg(8, 9).mean(), seg(9, 10).mean(), seg(10, 13).mean(), seg(13, 11).mean(), seg(11, 12).mean(), np.diff(t[:, 8]).astype(float).mean()))
print("D_s^-1 published %.0f cycles before column warp 0 starts waiting for it" % ((t[:, 0] - t[:, 11]).astype(float).mean()))
if len(sys.argv) > 4:
    print("step | col warp0: wait  Ub  update  pub  top || inverter: wait - jacobi ns pub rest")
    for i in range(0, 40):
        r = t[i]
        nxt = t[i + 1]
        print("%3d | %5d %5d %5d %5d %5d || %5d %5d %5d %5d %5d %5d" % (i, r[1] - r[0], r[2] - r[1], r[4] - r[2], r[5] - r[4], nxt[0] - r[5],
              r[9] - r[8], r[10] - r[9], r[14] - r[10], r[13] - r[14], r[11] - r[13], nxt[8] - r[11]))
if len(sys.argv) > 4 and sys.argv[4] == "roles":
    # column warp 0 owns column 0 at step 0: (column - s) mod 13 = (-s) mod 13; it is the next-column owner when that is 1.
    # trace rows are steps 100..163
    print("role(jrel) | ready-before-X (wait) | Ub | update | end-of-step work | X_s published relative to warp-0 ready | D_{s+1} at inverter rel. to X_s published")
    for i in range(0, 60):
        s = 100 + i
        jr = (-s) % 13
        r = t[i]
        print("%3d jrel=%2d | %5d %5d %5d %5d | %6d | %6d" % (s, jr, r[1] - r[0], r[2] - r[1], r[4] - r[2], r[5] - r[4], r[11] - r[0], t[i + 1][9] - r[11]))
