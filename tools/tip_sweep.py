import sys; sys.path.insert(0,'.')
import torch, spike_petsc_b200 as sp
n,k=10_000_000,100
for delta in (1.2,1.0):
  for tip in (26,39,52,65,78,104):
    S=sp.Spike(partitions=296,tip_tiles=tip,mem=sp.MEM_DEVICE); S.keep_original(True); S.set_band_synthetic(n,k,delta=delta)
    u=torch.ones(n,dtype=torch.float64,device='cuda'); b=torch.empty_like(u); x=torch.empty_like(u)
    S.mult(u.data_ptr(),b.data_ptr()); S.factor(); S.solve(b.data_ptr(),x.data_ptr()); torch.cuda.synchronize()
    st=S.view()['stage_ms']
    print(f"delta={delta} tip_tiles={tip} ({tip/13:.0f} bandwidths): relerr={(x-u).norm().item()/u.norm().item():.2e} windows={st[0]:.3f} corr={st[5]:.3f}",flush=True)
    S.close(); del u,b,x; torch.cuda.empty_cache()
