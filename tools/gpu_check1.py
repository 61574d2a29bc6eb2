import sys, time, numpy as np
sys.path.insert(0,'.')
from oracle import oracle as O
import spike_petsc_b200 as sp
def check(n,k,P,tip,delta=1.2):
    a=O.gen_band(n,k,delta=delta)
    u=np.ones(n); b=O.band_mult(a,u)
    S=sp.Spike(partitions=P,tip_tiles=tip)
    S.set_band_dense(a,k)
    # band round trip
    rt=S.get_band_rows(); print(f"n={n} k={k} P={P} tip={tip}: roundtrip maxdiff",np.abs(rt-a).max())
    y=S.mult(u); print("  mult err",np.abs(y-b).max())
    S.factor()
    info=S.view(); print("  info",{k_:info[k_] for k_ in ('kt','partitions','tip_tiles','boosted_pivots','factor_ms')})
    # compare factors of partition structure with oracle spike
    x=S.solve(b); print("  solve err vs u",np.abs(x-u).max()/1.0)
    lu,_=O.band_lu(a); xo=O.band_solve(lu,b); print("  vs oracle exact LU",np.abs(x-xo).max())
    if info['partitions']==1:
        wide,nb,kw=O.block_lu(a); f=S.get_band_rows(); print("  factor entries maxdiff vs oracle block LU",np.abs(f-wide[:,kw-k:kw+k+1]).max())
    return S
check(4096,20,1,-1)
check(4096,20,4,-1)
check(4096,100,1,-1)
check(8192,100,2,-1)
check(20000,50,8,-1)
check(20000,50,8,0)
check(20001,37,5,30)
check(4000,10,4,-1)
