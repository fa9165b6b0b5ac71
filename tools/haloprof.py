"""time the halo conv of one layer shape at full KITTI batch size (ablation via GCLB_HALO_DBG)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from gcl_b200 import ops
dev = torch.device("cuda:0")
x, p = bench.make_batches(1, 16, seed=0)[0]
cm1, _ = ops.voxelize(x.to(dev), 0.3, p)
shapes = [(64, 64), (32, 32), (128, 128)] if len(sys.argv) < 2 else [tuple(map(int, a.split("x"))) for a in sys.argv[1:]]
nbr, keys = ops.kernel_map(cm1, cm1, 3, with_keys=True)
srt, perm, mask = ops.kernel_map_sort(nbr, keys, copy=True)
halo = ops.kernel_map_halo(nbr, perm)
n = cm1.n
for cin, cout in shapes:
  xx = torch.randn(n, cin, device=dev).half()
  W = ops.weights_to_tc(torch.randn(27, cin, cout, device=dev) / 30, half=True)
  def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
  th = t(lambda: ops.spconv_fwd_halo(xx, W, halo, relu=True))
  td = t(lambda: ops.spconv_fwd(xx, W, srt, n, algo=2, row_perm=perm, tile_mask=mask, relu=True))
  print(f"dbg={os.environ.get('GCLB_HALO_DBG', '0')} {cin}->{cout} rows {n}: halo {th:.1f} us  direct {td:.1f} us")
  if int(os.environ.get("GCLB_HALO_DBG", "0")) & 512:
    import ctypes, numpy as np
    from gcl_b200 import _lib
    buf = np.zeros((148, 16), np.uint64)
    ops.spconv_fwd_halo(xx, W, halo, relu=True); torch.cuda.synchronize()
    _lib.load().gclb_debug_halo_prof(buf.ctypes.data)
    m = buf.astype(np.float64).mean(0)
    names = ["copy:wait_meta", "copy:wait_halo", "copy:wait_empty", "copy:work", "copy:total", "mma:wait_meta", "mma:wait_acc_empty",
             "mma:wait_full", "mma:total", "epi:wait_acc", "epi:total", "halo:wait_meta", "halo:wait_empty", "halo:total", "tiles", "stages"]
    print("  per-CTA mean cycles:", ", ".join(f"{n}={v:.0f}" for n, v in zip(names, m)))
