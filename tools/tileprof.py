"""per-tile vs per-stage cost of the direct tcgen05 conv: force the populated-offset mask of every tile (results are wrong on
purpose; only the timing matters).  Fit: time = tiles/148 * (t_tile + stages * t_stage)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from gcl_b200 import ops
dev = torch.device("cuda:0")
x, p = bench.make_batches(1, 16, seed=0)[0]
cm1, _ = ops.voxelize(x.to(dev), 0.3, p)
nbr, keys = ops.kernel_map(cm1, cm1, 3, with_keys=True)
srt, perm, mask = ops.kernel_map_sort(nbr, keys, copy=True)
n = cm1.n
def t(fn, reps=10):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e3
order = [13, 12, 14, 10, 16, 4, 22, 9, 11, 15, 17, 3, 5, 21, 23, 1, 7, 19, 25, 0, 2, 6, 8, 18, 20, 24, 26]
for cin, cout in [(64, 64), (256, 256)]:
  xx = torch.randn(n, cin, device=dev).half()
  W = ops.weights_to_tc(torch.randn(27, cin, cout, device=dev) / 30, half=True)
  res = []
  for nbits in (1, 4, 8, 16):
    bits = sum(1 << k for k in order[:nbits])
    m = torch.full_like(mask, bits)
    res.append((nbits, t(lambda: ops.spconv_fwd(xx, W, srt, n, algo=2, row_perm=perm, tile_mask=m, relu=True))))
  real = t(lambda: ops.spconv_fwd(xx, W, srt, n, algo=2, row_perm=perm, tile_mask=mask, relu=True))
  avg = sum(bin(int(v) & 0xffffffff).count("1") for v in mask.tolist()) / mask.numel()
  tiles_per_sm = mask.numel() / 148
  (b1, u1), (b2, u2) = res[1], res[-1]
  slope = (u2 - u1) / (b2 - b1) / tiles_per_sm
  icpt = u1 / tiles_per_sm - slope * b1
  print(f"dbg={os.environ.get('GCLB_TC_DBG','0')} {cin}->{cout}: real mask ({avg:.1f} offsets/tile) {real:.1f} us; forced: " + ", ".join(f"{b}:{u:.0f}" for b, u in res) +
        f"  => per offset {slope * 1e3:.0f} ns, per tile {icpt * 1e3:.0f} ns")
