"""Near-warp chain of the narrow sweep, operation by operation (build: tools/build_variant.sh swfine "-DSW_TRACE -DSW_TRACE_FINE")."""
import sys, ctypes as C; sys.path.insert(0, '.')
import numpy as np, torch, spike_petsc_b200 as sp
n, k, P, tip = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (1_250_000, 100, 296, 78)
S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(n, k)
u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u); x = torch.empty_like(u)
S.mult(u.data_ptr(), b.data_ptr())
for _ in range(3):
    S.factor(); S.solve(b.data_ptr(), x.data_ptr())
torch.cuda.synchronize()
out = np.zeros((32, 16), dtype=np.int64)
assert sp.lib().spk_debug_sweep_trace(out.ctypes.data_as(C.c_void_p)) == 0
order = [(0, "loop top"), (1, "cg ready (rhs, 3 far partials: LDS + 3 DADD)"), (7, "adjacent tile in registers (LDS.128)"), (8, "DMUL + DFMA"),
         (9, "SHFL.BFLY 1 + DADD"), (10, "SHFL.BFLY 2 + DADD"), (2, "cg - part (DADD)"), (3, "D^-1 product: 2 SHFL.IDX, DMUL+DFMA, 2x(SHFL+DADD)"),
         (4, "STS y, 2 SHFL.IDX"), (5, "__syncthreads"), (6, "sink (STG)")]
prev = None
for slot, name in order:
    col = out[1:-1, slot]
    if prev is not None:
        d = col - prev
        print(f"{name:>60}: +{int(np.median(d)):4d} cycles (min {int(d.min())}, max {int(d.max())})")
    prev = col
print("period:", int(np.median(np.diff(out[:, 0]))))
