import sys,time,os
sys.path.insert(0,'.')
import numpy as np
from oracle import cref
from oracle import oracle as O
g='bn254_g1'; G=O.GROUPS[g]
n=1<<20
base=G.encode_affine([G.gen])[0]
t=time.time(); pts=cref.generate_multiples(g,base,1,n,nthreads=64); print('gen',round(time.time()-t,3))
s=cref.random_scalars(g,n,1)
for th,nbt,c in ((1,1,16),(16,0,0),(32,0,0),(64,0,0),(128,0,0),(64,64,0),(128,16,0),(128,0,16),(64,0,16)):
    best=1e9
    for rep in range(2):
        t=time.time(); r=cref.msm(g,pts,s,c=c,nthreads=th,nb_tasks=nbt); best=min(best,time.time()-t)
    print('threads',th,'nb_tasks',nbt,'c',r[2],'leaves',r[3],'time',round(best,3), flush=True)
