import sys, json, torch
sys.path.insert(0,'.')
import bench
from gcl_b200 import MinkowskiEngine as ME, ops
from gcl_b200.pipeline import PairMatcher
dev=torch.device('cuda:0')
model=bench.seeded_model(ME)
m=PairMatcher(model, device=dev)
host=bench.make_batches(1, 8, seed=0)
x,p=host[0]; x=x.to(dev)
eng=m.engine
cm1,_=ops.voxelize(x, 0.3, p)
maps=eng.build_maps(cm1); cms,km=maps
print({k:(tuple((v[0] if isinstance(v,tuple) else v).shape), round(((v[0] if isinstance(v,tuple) else v)>=0).float().mean().item(),3)) for k,v in km.items()})
recs=[]
orig=ops.spconv_fwd
names=[]
def timed(in0,W,nbr,n_out,in1=None,**kw):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); out=orig(in0,W,nbr,n_out,in1=in1,**kw); e1.record()
    recs.append((e0,e1,tuple(W.shape),n_out,kw.get('algo'))); return out
feats=torch.ones((cm1.n,1),device=dev)
for _ in range(3):
    recs.clear(); ops.spconv_fwd=timed
    eng.forward(cm1,feats,maps); ops.spconv_fwd=orig
torch.cuda.synchronize()
for e0,e1,ws,n,a in recs: print(f"{e0.elapsed_time(e1):8.3f} ms  W={ws} n_out={n} algo={a}")
# stage timings
def t(fn,n=5):
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize(); e0.record()
    for _ in range(n): r=fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
print('voxelize', t(lambda: ops.voxelize(x,0.3,p)))
print('build_maps', t(lambda: eng.build_maps(cm1)))
print('forward', t(lambda: eng.forward(cm1,feats,maps)))
print('match total', t(lambda: m.match(x,p)))
import time
t0=time.perf_counter(); 
for _ in range(5): m.match(x,p)
torch.cuda.synchronize(); print('wall match', (time.perf_counter()-t0)/5*1e3)
