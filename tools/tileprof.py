"""per-tile vs per-stage cost of the direct tcgen05 conv: force the populated-offset mask of every tile (results are wrong on
purpose; only the timing matters).  Fit: time = tiles/148 * (t_tile + stages * t_stage)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from gcl_b200 import ops
dev = torch.device("cuda:0")
PAIRS = int(sys.argv[1]) if len(sys.argv) > 1 else 16
x, p = bench.make_batches(1, PAIRS, seed=0)[0]
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
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[2:]] or ([(64, 64)] if PAIRS != 16 else [(64, 64), (256, 256)])
for cin, cout in SHAPES:
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
  # the same forced masks over a table whose gathers are sequential rows (nbr[r][k] = r + k): separates the cost of the
  # scattered access pattern from the cost of the bytes
  seq = (torch.arange(n, device=dev, dtype=torch.int32)[:, None] + torch.arange(27, device=dev, dtype=torch.int32)[None, :]).clamp_(max=n - 1).contiguous()
  ident = torch.arange(n, device=dev, dtype=torch.int32)
  rs = []
  for nbits in (4, 16):
    bits = sum(1 << k for k in order[:nbits])
    m = torch.full_like(mask, bits)
    rs.append((nbits, t(lambda: ops.spconv_fwd(xx, W, seq, n, algo=2, row_perm=ident, tile_mask=m, relu=True))))
  ar = torch.arange(n, device=dev, dtype=torch.int64)
  k27 = torch.arange(27, device=dev, dtype=torch.int64)
  variants = {
    "stride-2 rows, re-read per offset": (((ar[:, None] + k27[None, :]) * 2) % n),
    "random contiguous quads, no reuse": (torch.randint(0, n // 4, (n // 4 + 1, 27), device=dev)[ar // 4] * 4 + (ar % 4)[:, None]),
    "random contiguous pairs, no reuse": (torch.randint(0, n // 2, (n // 2 + 1, 27), device=dev)[ar // 2] * 2 + (ar % 2)[:, None]),
    "random rows, same rows for every offset": torch.randperm(n, device=dev)[:, None].expand(n, 27),
    "random rows, no reuse": torch.randint(0, n, (n, 27), device=dev),
  }
  for name, tab in variants.items():
    tab = tab.to(torch.int32).contiguous()
    r2 = []
    for nbits in (4, 16):
      bits = sum(1 << k for k in order[:nbits])
      m = torch.full_like(mask, bits)
      r2.append(t(lambda: ops.spconv_fwd(xx, W, tab, n, algo=2, row_perm=ident, tile_mask=m, relu=True)))
    print(f"   {name}: 4:{r2[0]:.0f}, 16:{r2[1]:.0f} => per offset {(r2[1] - r2[0]) / 12 / tiles_per_sm * 1e3:.0f} ns")
  print(f"   sequential-row table: " + ", ".join(f"{b}:{u:.0f}" for b, u in rs) + f" => per offset {(rs[1][1] - rs[0][1]) / 12 / tiles_per_sm * 1e3:.0f} ns")
  # the real table with its rows' neighbours shuffled across the whole tensor (same bytes, no spatial locality at all)
  shuf = srt.clone(); v = shuf >= 0
  shuf[v] = torch.randint(0, n, (int(v.sum()),), device=dev, dtype=torch.int32)
  nomiss = srt.clone(); v0 = nomiss < 0
  nomiss[v0] = torch.randint(0, n, (int(v0.sum()),), device=dev, dtype=torch.int32)
  r3 = []
  for nbits in (4, 16):
    bits = sum(1 << k for k in order[:nbits])
    m = torch.full_like(mask, bits)
    r3.append(t(lambda: ops.spconv_fwd(xx, W, nomiss, n, algo=2, row_perm=perm, tile_mask=m, relu=True)))
  print(f"   real table, missing neighbours -> random rows (in bounds): real mask {t(lambda: ops.spconv_fwd(xx, W, nomiss, n, algo=2, row_perm=perm, tile_mask=mask, relu=True)):.1f} us; "
        f"forced 4:{r3[0]:.0f}, 16:{r3[1]:.0f} => per offset {(r3[1] - r3[0]) / 12 / tiles_per_sm * 1e3:.0f} ns")
  print(f"   rows {n}: real table {real:.1f} us, same table with random neighbour rows {t(lambda: ops.spconv_fwd(xx, W, shuf, n, algo=2, row_perm=perm, tile_mask=mask, relu=True)):.1f} us")
  print(f"dbg={os.environ.get('GCLB_TC_DBG','0')} {cin}->{cout}: real mask ({avg:.1f} offsets/tile) {real:.1f} us; forced: " + ", ".join(f"{b}:{u:.0f}" for b, u in res) +
        f"  => per offset {slope * 1e3:.0f} ns, per tile {icpt * 1e3:.0f} ns")
