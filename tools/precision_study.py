"""Operand-precision study for the conv GEMMs (evidence for the storage/MMA dtype choice).
Runs the seeded ResUNetBN2C on one KITTI-shape pair with the EXACT fp32 CUDA-core kernels, rounding the conv operands
(activations + weights) on the fly to emulate a tensor-core kind with fp32 accumulation, and reports the feature error
against the unrounded run:   tf32 = 10-bit mantissa, truncated activations (what kind::tf32 does to fp32 smem operands);
fp16 = 10-bit mantissa, round-to-nearest (activations stored as fp16 in HBM); bf16 = 7-bit mantissa."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from gcl_b200 import MinkowskiEngine as ME, ops

dev = torch.device("cuda:0")
model = bench.seeded_model(ME).to(dev)
xyz, ptr = bench.make_batches(1, 1, seed=0)[0]
cm, _ = ops.voxelize(xyz.to(dev), bench.VOXEL, ptr)
ME.set_inference_conv_algo("fp32")
orig = ops.spconv_fwd


def trunc_tf32(t):
  return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def rn_tf32(t):
  u = t.view(torch.int32)
  return ((u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF).view(torch.float32)


def make(mode, min_cin):
  def f(in0, W, nbr, n_out, in1=None, **kw):
    cin = in0.shape[1] + (in1.shape[1] if in1 is not None else 0)
    m = mode if (cin % min_cin == 0) else ("tf32" if cin >= 32 else "fp32")
    if m == "tf32":
      r, rw = trunc_tf32, rn_tf32
    elif m == "fp16":
      r = rw = lambda t: t.half().float()
    elif m == "bf16":
      r = rw = lambda t: t.bfloat16().float()
    else:
      r = rw = lambda t: t
    return orig(r(in0.contiguous()), rw(W.contiguous()), nbr, n_out, in1=(r(in1.contiguous()) if in1 is not None else None), **kw)
  return f


def run():
  st = ME.SparseTensor(torch.ones(cm.n, 1, device=dev), coordinates=cm.coords)
  with torch.no_grad():
    return model(st).F


ref = run()
out = {"voxels": int(cm.n)}
for mode, min_cin in [("tf32", 32), ("fp16", 64), ("fp16", 32), ("bf16", 32)]:
  ops.spconv_fwd = make(mode, min_cin)
  got = run()
  ops.spconv_fwd = orig
  rel = ((got - ref).norm() / ref.norm()).item()
  worst = (got - ref).norm(dim=1).max().item()       # rows are unit vectors: this is relative per row
  out[f"{mode}_cin%{min_cin}"] = {"rel_frobenius": rel, "worst_row": worst}
print(json.dumps(out))
