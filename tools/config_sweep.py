"""Stage timings for the BASELINE configs that fit one GPU.  usage: config_sweep.py [n,k,P,tip ...]"""
import sys; sys.path.insert(0, '.')
import torch, spike_petsc_b200 as sp
cases = [(100_000, 10, 0, 0), (1_000_000, 50, 592, 0), (1_000_000, 50, 1184, 0), (10_000_000, 100, 296, 78)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
for n, k, P, tip in cases:
    best = None
    for rep in range(3):
        S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(n, k)
        u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u); x = torch.empty_like(u)
        S.mult(u.data_ptr(), b.data_ptr()); S.factor(); S.solve(b.data_ptr(), x.data_ptr()); torch.cuda.synchronize()
        info = S.view(); err = ((x - u).norm() / u.norm()).item()
        if best is None or info['factor_ms'] + info['solve_ms'] < best[0]:
            best = (info['factor_ms'] + info['solve_ms'], info, err)
        S.close(); del u, b, x
    tot, info, err = best
    B = 8.0 * n * (2 * k + 1)
    st = info['stage_ms']
    print(f"n={n} k={k} P={info['partitions']} tip={info['tip_tiles']}: factor {info['factor_ms']:.3f} ms ({2*B/info['factor_ms']/1e6/6555.2*100:.1f}% HBM) "
          f"solve {info['solve_ms']:.3f} ms ({(B+32*n)/info['solve_ms']/1e6/6555.2*100:.1f}% HBM) err {err:.1e} | windows {st[0]:.3f} lu {st[1]:.3f} tips {st[2]:.3f} sweeps {st[3]:.3f} red {st[4]:.3f} corr {st[5]:.3f}", flush=True)
