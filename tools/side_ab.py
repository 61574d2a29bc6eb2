"""A/B of the side stream (spike tips + reduced blocks next to the first solve's sweeps, capi.cu SideScope):
whole factor+solve step time with SPIKE_B200_SIDE_STREAM=0 / 1, CUDA events around K back-to-back steps,
solution compared bit for bit between the two modes.  usage: side_ab.py [n,k,P,tip ...]
(SIDE_AB_MODES=1: only the product mode, as a whole-step partition / window sweep; SIDE_AB_TIMING=0: the engine's
event timers off, as in the library default -- the stage columns are then zero)"""
import os, sys; sys.path.insert(0, '.')
import torch, spike_petsc_b200 as sp
cases = [(10_000_000, 100, 296, 78), (1_250_000, 100, 296, 78), (1_000_000, 50, 592, 48), (100_000, 10, 296, 24)]
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
K = 10
for n, k, P, tip in cases:
    xs = {}
    for mode in os.environ.get("SIDE_AB_MODES", "0,1,0,1").split(","):
        os.environ["SPIKE_B200_SIDE_STREAM"] = mode
        S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE, timing=os.environ.get('SIDE_AB_TIMING', '1') != '0'); S.keep_original(True); S.set_band_synthetic(n, k)
        u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u); x = torch.empty_like(u)
        S.mult(u.data_ptr(), b.data_ptr())
        for _ in range(3):
            S.factor(); S.solve(b.data_ptr(), x.data_ptr())
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        ev[0].record()
        for i in range(K):
            S.factor(); S.solve(b.data_ptr(), x.data_ptr()); ev[i + 1].record()
        torch.cuda.synchronize()
        ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(K))
        info = S.view(); st = info['stage_ms']
        err = ((x - u).norm() / u.norm()).item()
        same = "" if mode not in xs else f" bit-identical to first run of this mode: {bool(torch.equal(xs[mode], x))}"
        if mode == "1" and "0" in xs: same += f", to side=0: {bool(torch.equal(xs['0'], x))}"
        xs.setdefault(mode, x.clone())
        print(f"n={n} k={k} P={info['partitions']} tip={info['tip_tiles']} side={mode}: step median {ms[K//2]:.4f} min {ms[0]:.4f} ms err {err:.1e} | "
              f"windows {st[0]:.3f} lu {st[1]:.3f} tips {st[2]:.3f} sweeps {st[3]:.3f} red {st[4]:.3f} corr {st[5]:.3f}{same}", flush=True)
        S.close(); del u, b, x
