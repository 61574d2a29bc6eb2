import os, sys, time; sys.path.insert(0, '.')
import torch, spike_petsc_b200 as sp
n, k = 100_000, 10
for mode in ("0", "1", "0", "1"):
    os.environ["SPIKE_B200_SIDE_STREAM"] = mode
    for P, tip in ((296, 24), (592, 12)):
        S = sp.Spike(partitions=P, tip_tiles=tip, mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(n, k)
        u = torch.ones(n, dtype=torch.float64, device='cuda'); b = torch.empty_like(u)
        S.mult(u.data_ptr(), b.data_ptr()); S.factor()
        ts = []
        for rep in range(4):
            xk = torch.zeros(n, dtype=torch.float64, device='cuda')
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _, its, rn, conv = S.krylov(b.data_ptr(), method=sp.GMRES, restart=30, rtol=1e-5, maxit=200, x=xk.data_ptr())
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        print(f"side={mode} P={P} tip={tip}: its {its} conv {conv} ms {[round(t,3) for t in ts]}", flush=True)
        S.close()
