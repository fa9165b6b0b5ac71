"""Halo-staging tcgen05 convolution (gcl_b200/csrc/spconv_halo.cu) against the direct-gather tcgen05 kernel and the fp32
CPU oracle: same-map 3x3x3 convolutions on fp16 activations, all channel widths of the ResUNet's residual blocks."""
import numpy as np
import pytest
import torch

import oracle.me_cpu as OME

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _rel(a, b):
  a, b = a.double().cpu(), b.double().cpu()
  return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _lidar_like(seed, n, extent):
  rng = np.random.RandomState(seed)
  ground = np.concatenate([rng.uniform(-extent, extent, (n, 2)), rng.normal(0, 0.05, (n, 1))], 1)
  wall = np.stack([rng.uniform(-extent, extent, n // 3), np.full(n // 3, 2.0) + rng.normal(0, 0.03, n // 3),
                   rng.uniform(0, 3, n // 3)], 1)
  return np.concatenate([ground, wall]).astype(np.float32)


def _dense_cube(side, seed):
  g = np.stack(np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij"), -1).reshape(-1, 3)
  rng = np.random.RandomState(seed)
  g = g[rng.permutation(len(g))]                    # random row order: every tile touches ~27 x 128 distinct rows
  return torch.from_numpy(np.concatenate([np.zeros((len(g), 1), np.int64), g], 1)).int()


def _maps(ops, coords):
  cm = ops.hash_build(coords.to(DEV))
  nbr, keys = ops.kernel_map(cm, cm, 3, with_keys=True)
  srt, perm, mask = ops.kernel_map_sort(nbr, keys, copy=True)
  halo = ops.kernel_map_halo(nbr, perm)
  return cm, nbr, srt, perm, mask, halo


@pytest.mark.parametrize("cin,cout", [(64, 64), (32, 32), (128, 128), (256, 256), (64, 128), (32, 64)])
def test_halo_conv_vs_direct_gather_and_oracle(cin, cout):
  from gcl_b200 import ops
  torch.manual_seed(cin + cout)
  from gcl_b200 import synth
  xyz = synth.cast(synth.Scene(7, n_boxes=25, extent=40.0), dict(synth.NUSCENES, azimuth_steps=700))
  t = torch.from_numpy(xyz)
  _, sel = OME.utils.sparse_quantize(t / 0.3, return_index=True)
  c = torch.floor(t[sel] / 0.3).int()
  coords, _ = OME.utils.sparse_collate([c], [torch.ones(len(c), 1)])
  cm, nbr, srt, perm, mask, halo = _maps(ops, coords)
  n = cm.n
  assert int(halo.status.item()) == 0
  ng = halo.tile_ngroups.cpu().numpy()
  used = int(halo.counter.item()) * 16
  print(f"\n[halo {cin}->{cout}] n={n} tiles={len(ng)} groups/tile mean {ng.mean():.2f} max {ng.max()} "
        f"record bytes {used} = {used / max(n, 1):.1f} B/row (worst-case buffer {halo.slots.numel()})")
  x = (torch.randn(n, cin) * 0.5)
  W = torch.randn(27, cin, cout) / np.sqrt(27 * cin) * 3
  sc, sh = torch.rand(cout) + 0.5, torch.randn(cout) * 0.1
  res = torch.randn(n, cout) * 0.3
  xd, Wimg = x.half().to(DEV), ops.weights_to_tc(W.to(DEV), half=True)
  kw = dict(scale=sc.to(DEV), shift=sh.to(DEV), relu=True)
  res_d = res.half().to(DEV) if cin == cout else None
  ref = ops.spconv_fwd(xd, Wimg, srt, n, algo=2, row_perm=perm, tile_mask=mask, residual=res_d, out_dtype=torch.float16, **kw)
  got = ops.spconv_fwd_halo(xd, Wimg, halo, residual=res_d, out_dtype=torch.float16, **kw)
  assert got.dtype == torch.float16 and got.shape == (n, cout)
  assert _rel(got, ref) < 2e-3 * 0 + 1e-3          # both round to fp16 at the end; accumulation order differs only for > 1 slab
  if cin <= 64:
    assert torch.equal(got, ref)                   # one slab, same offset order: bit-identical
  # fp32 output against the fp32 CPU oracle on the UNROUNDED operands
  got32 = ops.spconv_fwd_halo(xd, Wimg, halo, residual=res_d, out_dtype=torch.float32, **kw)
  want = OME.sparse_conv_reference(x, W, nbr.cpu().numpy(), n) * sc + sh
  if res_d is not None:
    want = want + res
  want = torch.relu(want)
  assert _rel(got32, want) < 1e-3
  # determinism
  assert torch.equal(got32, ops.spconv_fwd_halo(xd, Wimg, halo, residual=res_d, out_dtype=torch.float32, **kw))


def test_halo_conv_dense_cube_many_groups_and_ragged_tail():
  """a dense cube in random row order: every tile touches thousands of distinct rows => ~10 groups per tile; the last tile
  is ragged (n % 128 != 0)"""
  from gcl_b200 import ops
  torch.manual_seed(3)
  coords = _dense_cube(21, 1)                       # 9261 voxels
  cm, nbr, srt, perm, mask, halo = _maps(ops, coords)
  n = cm.n
  assert n % 128 != 0 and int(halo.status.item()) == 0
  ng = halo.tile_ngroups.cpu().numpy()
  assert ng.max() >= 4 and ng.max() <= 16, ng.max()
  x, W = torch.randn(n, 64) * 0.5, torch.randn(27, 64, 64) / 40
  xd, Wimg = x.half().to(DEV), ops.weights_to_tc(W.to(DEV), half=True)
  got = ops.spconv_fwd_halo(xd, Wimg, halo, out_dtype=torch.float32)
  ref = ops.spconv_fwd(xd, Wimg, srt, n, algo=2, row_perm=perm, tile_mask=mask, out_dtype=torch.float32)
  assert _rel(got, ref) < 1e-5
  want = OME.sparse_conv_reference(x, W, nbr.cpu().numpy(), n)
  assert _rel(got, want) < 1e-3
  # no permutation (identity tile order) and a tiny map (single partial tile)
  halo_id = ops.kernel_map_halo(nbr, None)
  got_id = ops.spconv_fwd_halo(xd, Wimg, halo_id, out_dtype=torch.float32)
  assert _rel(got_id, want) < 1e-3
  small = _dense_cube(3, 2)
  cm_s, nbr_s, _, perm_s, _, halo_s = _maps(ops, small)
  xs = torch.randn(cm_s.n, 64)
  got_s = ops.spconv_fwd_halo(xs.half().to(DEV), Wimg, halo_s, out_dtype=torch.float32)
  assert _rel(got_s, OME.sparse_conv_reference(xs, W, nbr_s.cpu().numpy(), cm_s.n)) < 1e-3
