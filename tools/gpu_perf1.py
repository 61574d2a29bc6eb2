import sys, time, numpy as np
sys.path.insert(0,'.')
import spike_petsc_b200 as sp
import torch
def run(n,k,P,tip,reps=3):
    S=sp.Spike(partitions=P,tip_tiles=tip,mem=sp.MEM_DEVICE)
    S.keep_original(True)
    S.set_band_synthetic(n,k)
    u=torch.ones(n,dtype=torch.float64,device='cuda'); b=torch.empty_like(u); x=torch.empty_like(u)
    S.mult(u.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    t0=time.time(); S.factor(); torch.cuda.synchronize(); tf=time.time()-t0
    info=S.view()
    ts=[]
    for r in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve(b.data_ptr(), x.data_ptr()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); S.mult(u.data_ptr(), b.data_ptr()); e1.record(); torch.cuda.synchronize(); tm=e0.elapsed_time(e1)
    err=(x-u).abs().max().item()
    B=8*n*(2*k+1)
    print(f"n={n} k={k} P={info['partitions']} tip={info['tip_tiles']} factor_ms={info['factor_ms']:.3f} (wall {tf*1e3:.1f}) solve_ms={min(ts):.3f} mult_ms={tm:.3f} err={err:.2e} | factor %HBM={2*B/info['factor_ms']/1e6/6555.2*100:.1f} solve %HBM={(B+32*n)/min(ts)/1e6/6555.2*100:.1f} mult %HBM={(B+16*n)/tm/1e6/6555.2*100:.1f} factor TF={n*(2*k*k+k)/info['factor_ms']/1e9:.2f}",flush=True)
    S.close()
run(1_000_000,50,0,0)
run(1_000_000,50,296,0)
run(1_000_000,50,592,0)
run(1_000_000,50,592,48)
run(10_000_000,100,148,0)
run(10_000_000,100,296,0)
run(10_000_000,100,296,104)
run(10_000_000,100,592,104)
