"""Drop-in path timing: the table-driven ResUNetBN2C graph (== the reference's model/resunet.py graph) executed op by op on
gcl_b200.MinkowskiEngine in eval mode (tcgen05 convs, separate BatchNorm / ReLU / cat kernels), vs the fused engine."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gcl_b200 import MinkowskiEngine as ME, ops
from gcl_b200.engine import ResUNetEngine

dev = torch.device("cuda:0")
model = bench.seeded_model(ME).to(dev)
xyz, ptr = bench.make_batches(1, 8, seed=0)[0]
xyz = xyz.to(dev)
eng = ResUNetEngine(model, device=dev)


def dropin():
  cm, _ = ops.voxelize(xyz, bench.VOXEL, ptr)
  st = ME.SparseTensor(torch.ones(cm.n, 1, device=dev), coordinates=cm.coords)
  with torch.no_grad():
    return model(st).F


def fused():
  return eng.extract(xyz, bench.VOXEL, ptr)[0]


def timeit(fn, n=10):
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n):
    out = fn()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n, out


t_d, f_d = timeit(dropin)
t_f, f_f = timeit(fused)
rel = ((f_d - f_f).norm() / f_f.norm()).item()
print(json.dumps({"clouds": 16, "voxels": int(f_f.shape[0]), "dropin_module_path_ms": round(t_d, 2), "fused_engine_ms": round(t_f, 2),
                  "rel_diff_between_paths": rel}))
