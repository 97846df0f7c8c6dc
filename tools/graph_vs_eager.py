"""Why the CUDA-graph figure of bench.py refills its actions (GPU box): step time with i.i.d. action windows (eager), with a
graph of G steps replayed on the SAME G action sets, and with one constant action tensor.  Periodic actions give every env a
net drift that pins the agents against the walls: the contact work, not the graph, is what gets slower.
    python tools/graph_vs_eager.py [envs]"""
import sys, os, time; sys.path.insert(0,'.')
import numpy as np, torch
from marl_soccer_b200 import _capi
from marl_soccer_b200.sim import BatchedSoccerSim, load_default_config
import bench
dev=torch.device('cuda:0'); cfg=load_default_config()
n=int(sys.argv[1]) if len(sys.argv)>1 else 65536
sim=BatchedSoccerSim(n,config=cfg,device=dev,seed=1); sim.reset(_capi.MODE_FULL_RANDOM,seed=1)
pool=bench.make_pool(torch,np,n,dev,7)
bench.preroll(torch,sim,pool,cfg,1000,_capi)
def timeit(fn,reps):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); 
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
k=[0]
def eager20():
    for i in range(20): sim.step(pool[1000+k[0]]); k[0]+=1
print('eager ms/step', timeit(eager20,10)/20)
for G in (1,5,20,100):
    side=torch.cuda.Stream(device=dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for i in range(3): sim.step(pool[2000+i])
    torch.cuda.current_stream(dev).wait_stream(side)
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g,stream=side):
        for i in range(G): sim.step(pool[3000+i])
    ms=timeit(g.replay, max(2,200//G))/G
    print('graph G=%d ms/step %.4f'%(G,ms)); del g
# same action tensor every step, eager
a=pool[5]
def eager_same():
    for i in range(20): sim.step(a)
print('eager same-actions ms/step', timeit(eager_same,10)/20)
