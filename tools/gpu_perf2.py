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
    S.factor(); torch.cuda.synchronize()
    info=S.view()
    ts=[]
    for r in range(reps):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); S.solve(b.data_ptr(), x.data_ptr()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    err=((x-u).norm()/u.norm()).item()
    B=8*n*(2*k+1)
    print(f"n={n} k={k} P={info['partitions']} tip={info['tip_tiles']} factor_ms={info['factor_ms']:.3f} solve_ms={min(ts):.3f} err={err:.2e} | factor %HBM={2*B/info['factor_ms']/1e6/6555.2*100:.1f} solve %HBM={(B+32*n)/min(ts)/1e6/6555.2*100:.1f} factor TF={n*(2*k*k+k)/info['factor_ms']/1e9:.2f}",flush=True)
    S.close()
if __name__=="__main__":
    cases=[(400_000,100,148,104),(1_000_000,50,296,0),(10_000_000,100,148,0),(10_000_000,100,296,104)]
    if len(sys.argv)>1: cases=[tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]]
    for c in cases: run(*c)
