#!/usr/bin/env python
"""potts/neighonly Q=10000 on sc/26n from a random start (CPU oracle, 96^3): how many DISTINCT neighbour
values a site sees along a grain-growth run, per site and as the maximum over the 24-32 sites a warp of the
streaming kernel handles together.  The numbers behind DESIGN.md 4.6 (profiles/r02_neigh_distinct_stats.txt)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import oracle_py as O
n=96; q=10000; T=0.5
o=O.Oracle("potts/neighonly","sc/26n",n,n,n,q,T)
rng=np.random.default_rng(1)
spins=rng.integers(1,q+1,size=o.nsites,dtype=np.int32)
def stats(sp):
    a=sp.reshape(n,n,n)  # z,y,x
    nb=[]
    for dz in (-1,0,1):
        for dy in (-1,0,1):
            for dx in (-1,0,1):
                if dx==0 and dy==0 and dz==0: continue
                nb.append(np.roll(a,(-dz,-dy,-dx),(0,1,2)))
    nb=np.stack(nb,-1)  # z,y,x,26
    e0=(nb!=a[...,None]).sum(-1)
    srt=np.sort(np.where(nb==a[...,None],0,nb),-1)
    distinct=(np.diff(srt,axis=-1)!=0).sum(-1)+ (srt[...,0]!=0)
    # warp groups: same colour: x parity fixed, lanes hold x=4*lane+P and +2; 32 lanes => 128 x-sites; n=96 -> use 24 lanes
    d=distinct[:, :, 0::2]  # colour x parity 0: shape z,y,48
    # thread holds (x, x+2) ; site slot h=0: x=0,4,8.. ; over lanes
    w=d[:,:,0::2].max(-1)  # warp max over lanes (24 lanes here), slot h=0
    return (e0==0).mean(), distinct.mean(), np.percentile(distinct,[50,90,99]), w.mean(), np.percentile(w,[50,90]), (d[:,:,0::2]==0).all(-1).mean()
for sw in range(0,201):
    if sw in (0,1,2,3,5,8,10,15,20,25,30,50,75,100,150,200):
        f0,m,pc,wm,wp,allz=stats(spins)
        print(sw,"P(e0=0)=%.3f mean distinct=%.2f pct=%s warpmax mean=%.2f pct=%s warp-all-zero=%.3f"%(f0,m,pc,wm,wp,allz),flush=True)
    o.sweep_philox(spins,12345,sw)
