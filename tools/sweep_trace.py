"""Cycle stamps of the narrow sweep kernel (trace build: tools/build_variant.sh swtrace "-DSW_TRACE").
usage: SPIKE_B200_LIB=build/var/swtrace/libspike_b200.so python tools/sweep_trace.py [n k P tip]
Prints, for forward iterations 32..63 of CTA 0 of the LAST k_sweep launch (the corrections when P > 1; P = 1: the
partition sweep), the near warp's phases, the far warp's and the copy warp's, in cycles relative to the near warp's
loop top."""
import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch, spike_petsc_b200 as sp
n, k, P, tip = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (1_250_000, 100, 296, 78)
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(n, k)
u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u); x = torch.empty_like(u)
S.mult(u.data_ptr(), b.data_ptr())
for _ in range(3):
    S.factor(); S.solve(b.data_ptr(), x.data_ptr())
torch.cuda.synchronize()
L = sp.lib()
out = np.zeros((32, 16), dtype=np.int64)
assert L.spk_debug_sweep_trace(out.ctypes.data_as(C.c_void_p)) == 0
names = ["top", "cg ready", "y pre-Dinv", "y final", "stored+shfl", "after sync", "after sink", "-", "far top", "far mbar ok", "far done", "far after sync", "copy top", "copy issued", "copy after sync"]
t0 = out[:, 0:1]
rel = out - t0
print("iteration period (near loop top to next):", np.diff(out[:, 0]).tolist())
print("near top -> after sync per iteration:", rel[:, 5].tolist())
print("copy top -> issued per iteration:", (out[:, 13] - out[:, 12]).tolist())
for j in (1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12, 13, 14):
    print(f"{names[j]:>16}: median {int(np.median(rel[1:-1, j]))}  min {int(rel[1:-1, j].min())} max {int(rel[1:-1, j].max())}")
