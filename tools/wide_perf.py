"""Stage timings of the wide-band path.  usage: wide_perf.py n,k,P,tip,nrhs[,G] ..."""
import sys; sys.path.insert(0, '.')
import os, torch, spike_petsc_b200 as sp
cases = [(1_000_000, 512, 0, 0, 32)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for case in cases:
    n, k, P, tip, nrhs = case[:5]
    S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE)
    S.set_band_synthetic(n, k)
    U = torch.rand(nrhs, n, dtype=torch.float64, device='cuda'); B = torch.empty_like(U); X = torch.empty_like(U)
    for r in range(nrhs):
        S.mult(U[r].data_ptr(), B[r].data_ptr())
    torch.cuda.synchronize()
    S.factor(); S.solve(B.data_ptr(), X.data_ptr(), nrhs=nrhs); torch.cuda.synchronize()
    info = S.view(); err = ((X - U).norm() / U.norm()).item()
    st = info['stage_ms']
    flops = n * (2.0 * k * k + k)
    sflops = 2.0 * n * (2 * k + 1) * nrhs
    print(f"n={n} k={k} P={info['partitions']} tip={info['tip_tiles']} nrhs={nrhs}: factor {info['factor_ms']:.3f} ms "
          f"solve {info['solve_ms']:.3f} ms err {err:.2e} boosted {info['boosted_pivots']} | windows {st[0]:.3f} lu {st[1]:.3f} "
          f"({flops/st[1]/1e9:.2f} TFLOP/s alg) tips {st[2]:.3f} sweeps {st[3]:.3f} ({sflops/st[3]/1e9:.2f} TFLOP/s) red {st[4]:.3f} corr {st[5]:.3f}", flush=True)
    S.close(); del U, B, X
