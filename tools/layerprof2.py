"""Per-layer timing of the convolution launches of one 16-pair forward, halo-staging kernel vs direct-gather kernel
(CUDA events around each launch while the GPU is kept busy).  usage: python tools/layerprof2.py [pairs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from gcl_b200 import MinkowskiEngine as ME, ops  # noqa: E402
from gcl_b200.engine import ResUNetEngine  # noqa: E402

dev = torch.device("cuda:0")
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = bench.seeded_model(ME)
x, p = bench.make_batches(1, pairs, seed=0)[0]
x = x.to(dev)
res = {}
for halo in (True, False):
  eng = ResUNetEngine(model, device=dev)
  eng.use_halo = halo
  cm1, _ = ops.voxelize(x, 0.3, p)
  maps = eng.build_maps(cm1)
  feats = torch.ones((cm1.n, 1), device=dev)
  eng.forward(cm1, feats, maps)
  cms, km = maps
  if halo:
    for name in ("k3s1", "k3s2", "k3s4", "k3s8"):
      h = km[name][4]
      ng = h.tile_ngroups.float()
      used = int(h.counter.item()) * 16
      print(f"{name}: rows {h.n_out} tiles {ng.numel()} groups/tile {ng.mean().item():.2f} (max {int(ng.max())}) "
            f"records {used / 1e6:.1f} MB = {used / h.n_out:.1f} B/row; pairs/row {(km[name][0] >= 0).float().sum().item() / h.n_out:.2f}")
  recs = []
  o1, o2 = ops.spconv_fwd, ops.spconv_fwd_halo

  def t1(in0, W, nbr, n_out, in1=None, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = o1(in0, W, nbr, n_out, in1=in1, **kw); e1.record()
    recs.append((e0, e1, "direct", in0.shape[1] + (in1.shape[1] if in1 is not None else 0), out.shape[1], n_out, W.shape[0] if W.dim() == 3 else 1))
    return out

  def t2(in0, W, h, in1=None, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = o2(in0, W, h, in1=in1, **kw); e1.record()
    recs.append((e0, e1, "halo", in0.shape[1], out.shape[1], h.n_out, 27))
    return out

  for _ in range(3):
    recs.clear()
    ops.spconv_fwd, ops.spconv_fwd_halo = t1, t2
    torch.cuda._sleep(int(2e7))
    eng.forward(cm1, feats, maps)
    ops.spconv_fwd, ops.spconv_fwd_halo = o1, o2
  torch.cuda.synchronize()
  res[halo] = [(k, cin, cout, n, K, a.elapsed_time(b) * 1e3) for a, b, k, cin, cout, n, K in recs]
print(f"{'layer':>3} {'cin':>4} {'cout':>4} {'K':>3} {'rows':>8} | {'halo us':>9} {'direct us':>9}  kind")
for i, (a, b) in enumerate(zip(res[True], res[False])):
  print(f"{i:3d} {a[1]:4d} {a[2]:4d} {a[4]:3d} {a[3]:8d} | {a[5]:9.1f} {b[5]:9.1f}  {a[0]}")
print("total", sum(a[5] for a in res[True]), sum(b[5] for b in res[False]))
