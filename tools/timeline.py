"""Debug tool (GPU box): timeline of the three step kernels within one step of the bench mix.  Needs a library built with
-DMSOC_TIMELINE (nvcc ... -DMSOC_TIMELINE -o scratch/ab/libmsoc_tl.so marl_soccer_b200/csrc/msoc.cu) passed through MSOC_LIB."""
import os, sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from marl_soccer_b200 import _capi
from marl_soccer_b200.sim import BatchedSoccerSim, load_default_config
n=int(sys.argv[1]) if len(sys.argv)>1 else 1<<20; dev=torch.device('cuda:0')
sim=BatchedSoccerSim(n, config=load_default_config(), device=dev, seed=0)
sim.reset(_capi.MODE_FULL_RANDOM, seed=0)
flat=torch.rand((16*n*12,),device=dev)*2-1
offs=np.random.default_rng(99).integers(0,15*n,size=4096)
phase=(torch.arange(n,device=dev,dtype=torch.int64)*2654435761)%1000
L=_capi.lib()
out=(C.c_ulonglong*(1<<18))(); cnt=C.c_uint(0)
for k in range(1003):
    sim.reset(_capi.MODE_FULL_RANDOM, mask=(phase==(k%1000)))
    if k==1002:
        torch.cuda.synchronize(); L.msoc_debug_timeline(out, C.byref(cnt))  # clears
    o=int(offs[k])*12; sim.step(flat[o:o+n*12].view(n,4,3))
torch.cuda.synchronize()
L.msoc_debug_timeline(out, C.byref(cnt))
m=min(cnt.value,1<<16)
a=np.array(out[:4*m],dtype=np.uint64).reshape(m,4).astype(np.int64)
t0=a[:,1].min()
print('records',m)
for kid,name in ((0,'fast(1/16 blocks)'),(1,'light'),(2,'heavy'),(3,'pair+multi')):
    r=a[a[:,0]==kid]
    if len(r)==0: continue
    st=(r[:,1]-t0)/1e3; en=(r[:,2]-t0)/1e3
    print(f'{name}: n {len(r)} start min {st.min():.0f} max {st.max():.0f} us | end min {en.min():.0f} p50 {np.median(en):.0f} p90 {np.percentile(en,90):.0f} max {en.max():.0f} us | dur p50 {np.median(en-st):.0f} p90 {np.percentile(en-st,90):.0f} max {(en-st).max():.0f}')
# active warps over time
T=int(((a[:,2].max()-t0)/1e3))+1
for kid,name in ((1,'light'),(2,'heavy'),(3,'pair+multi')):
    r=a[a[:,0]==kid]; act=np.zeros(T+1)
    for s_,e_ in zip(((r[:,1]-t0)/1e3).astype(int),((r[:,2]-t0)/1e3).astype(int)):
        act[s_:e_+1]+=1
    print(name,'active warp-batches sampled:',' '.join(str(int(x)) for x in act[::(20 if T>200 else 5)]))
