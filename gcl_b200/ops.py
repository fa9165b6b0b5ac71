"""Tensor-level wrappers over the C ABI (include/gclb200.h).  PyTorch is used for device memory and streams
only; every computation below is a libgclb200 kernel."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import call, ptr, stream, require_cuda


class CoordMap:
  """One coordinate map: int32 rows (b,x,y,z) at a tensor stride + its device hash table."""
  __slots__ = ("coords", "table", "capacity", "n", "tensor_stride", "status")

  def __init__(self, coords, table, capacity, n, tensor_stride, status=None):
    self.coords, self.table, self.capacity, self.n, self.tensor_stride = coords, table, capacity, n, tensor_stride
    self.status = status      # device int32 status word still to be checked (maps built with sync=False)


def _new_table(n_rows: int, device, upper_bound: bool = False) -> Tuple[torch.Tensor, int]:
  """upper_bound: n_rows is a host-side bound (point count / parent rows) far above the rows that will be inserted, so
  half the usual slack already gives a very low load factor (and halves the memset)."""
  lib = _lib.load()
  # 2 quad slots per row when the row count is exact; 1.34 per row (load <= 0.75 even if every row were its own quad, ~0.1
  # in practice) when n_rows is a bound
  cap = int(lib.gclb_hash_capacity((2 * int(n_rows) + 2) // 3 if upper_bound else int(n_rows)))
  return torch.empty(int(lib.gclb_hash_bytes(cap)), dtype=torch.uint8, device=device), cap


def _workspace(nbytes: int, device) -> torch.Tensor:
  return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=device)


def hash_build(coords4: torch.Tensor, tensor_stride: int = 1, check: bool = True) -> CoordMap:
  """SparseTensor construction (scripts/test_kitti.py:143-148): insert unique rows, row order preserved."""
  require_cuda(coords4)
  assert coords4.dtype == torch.int32 and coords4.dim() == 2 and coords4.shape[1] == 4
  coords4 = coords4.contiguous()
  n = coords4.shape[0]
  table, cap = _new_table(n, coords4.device)
  status = torch.zeros(1, dtype=torch.int32, device=coords4.device)
  call("gclb_hash_build", ptr(table), cap, int(tensor_stride), ptr(coords4), n, ptr(status), stream())
  if check:
    _lib.check_status(status, "SparseTensor coordinates")
  return CoordMap(coords4, table, cap, n, tensor_stride)


def hash_query(cm: CoordMap, q4: torch.Tensor) -> torch.Tensor:
  require_cuda(q4)
  q4 = q4.to(torch.int32).contiguous()
  out = torch.empty(q4.shape[0], dtype=torch.int32, device=q4.device)
  call("gclb_hash_query", ptr(cm.table), cm.capacity, int(cm.tensor_stride), ptr(q4), q4.shape[0], ptr(out), stream())
  return out


def voxelize(xyz: torch.Tensor, voxel: float, cloud_ptr: Optional[torch.Tensor] = None, return_inverse=False,
             check: bool = True, sync: bool = True):
  """K1: sparse_quantize(xyz / voxel) + floor().int() + sparse_collate in one pass
  (lib/complement_data_loader.py:788-789,809-812,1310-1311).
  xyz float32 [P,3] (all clouds concatenated), cloud_ptr int64 [n_clouds+1] (host or device).
  Returns (CoordMap of the V unique voxels in first-occurrence order, unique_map int64 [V] [, inverse int32 [P]])."""
  require_cuda(xyz)
  assert xyz.dtype == torch.float32 and xyz.dim() == 2 and xyz.shape[1] == 3
  xyz = xyz.contiguous()
  P, dev = xyz.shape[0], xyz.device
  if cloud_ptr is None:
    cloud_ptr = torch.tensor([0, P], dtype=torch.int64)
  n_clouds = cloud_ptr.numel() - 1
  if not cloud_ptr.is_cuda:   # tiny host array: stage through pinned memory so the copy never blocks the host
    cloud_ptr = cloud_ptr.to(torch.int64).pin_memory().to(dev, non_blocking=True)
  cloud_ptr = cloud_ptr.to(dtype=torch.int64).contiguous()
  lib = _lib.load()
  table, cap = _new_table(P, dev, upper_bound=True)
  coords = torch.empty((P, 4), dtype=torch.int32, device=dev)
  umap = torch.empty(P, dtype=torch.int64, device=dev)
  inv = torch.empty(P, dtype=torch.int32, device=dev) if return_inverse else None
  n_out = torch.zeros(1, dtype=torch.int64, device=dev)
  status = torch.zeros(1, dtype=torch.int32, device=dev)
  ws = _workspace(lib.gclb_compact_workspace_bytes(P), dev)
  call("gclb_voxelize", ptr(xyz), P, ptr(cloud_ptr), n_clouds, float(voxel), ptr(table), cap, ptr(coords), ptr(umap),
       ptr(inv), ptr(n_out), ptr(status), ptr(ws), stream())
  if not sync:   # row count and status stay on the device: finish with `finish_maps` (one host read for many maps)
    cm = CoordMap(coords, table, cap, n_out, 1, status)
    return (cm, umap, inv) if return_inverse else (cm, umap)
  if check:
    _lib.check_status(status, "voxelize")
  V = int(n_out.item())
  cm = CoordMap(coords[:V], table, cap, V, 1)
  return (cm, umap[:V], inv) if return_inverse else (cm, umap[:V])


def quantize_rows(rows: torch.Tensor, return_inverse=False):
  """first-occurrence de-duplication of int32 rows of width 3 or 4 (ME.utils.sparse_quantize on discrete input)."""
  require_cuda(rows)
  rows = rows.to(torch.int32).contiguous()
  P, width, dev = rows.shape[0], rows.shape[1], rows.device
  lib = _lib.load()
  table, cap = _new_table(P, dev)
  coords = torch.empty((P, 4), dtype=torch.int32, device=dev)
  umap = torch.empty(P, dtype=torch.int64, device=dev)
  inv = torch.empty(P, dtype=torch.int32, device=dev) if return_inverse else None
  n_out = torch.zeros(1, dtype=torch.int64, device=dev)
  status = torch.zeros(1, dtype=torch.int32, device=dev)
  ws = _workspace(lib.gclb_compact_workspace_bytes(P), dev)
  call("gclb_quantize_rows", ptr(rows), P, width, ptr(table), cap, ptr(coords), ptr(umap), ptr(inv), ptr(n_out),
       ptr(status), ptr(ws), stream())
  _lib.check_status(status, "sparse_quantize")
  V = int(n_out.item())
  cm = CoordMap(coords[:V], table, cap, V, 1)
  return (cm, umap[:V], inv) if return_inverse else (cm, umap[:V])


def stride_map(cm: CoordMap, stride: int, return_parent_rows=False, sync: bool = True):
  """Strided coordinate map: unique(floor(c / S) * S), S = tensor_stride * stride, first-appearance order.
  With sync=False the row count stays on the device (CoordMap.n is the device int64 tensor) and `coords`
  keeps its worst-case length; finalize with `finish_stride_maps`."""
  new_ts = cm.tensor_stride * stride
  dev = cm.coords.device
  lib = _lib.load()
  n_in_dev = cm.n if isinstance(cm.n, torch.Tensor) else None     # parent not finished yet: chain on the device
  n_in = cm.coords.shape[0] if n_in_dev is not None else cm.n     # host-side upper bound
  table, cap = _new_table(n_in, dev, upper_bound=True)
  coords = torch.empty((n_in, 4), dtype=torch.int32, device=dev)
  parent = torch.empty(n_in, dtype=torch.int32, device=dev) if return_parent_rows else None
  n_out = torch.zeros(1, dtype=torch.int64, device=dev)
  status = torch.zeros(1, dtype=torch.int32, device=dev)
  ws = _workspace(lib.gclb_compact_workspace_bytes(n_in), dev)
  call("gclb_stride_map", ptr(cm.coords), n_in, ptr(n_in_dev), new_ts, ptr(table), cap, ptr(coords), ptr(parent),
       ptr(n_out), ptr(status), ptr(ws), stream())
  out = CoordMap(coords, table, cap, n_out, new_ts, status)
  if sync:
    finish_maps([out])
  return (out, parent) if return_parent_rows else out


def finish_maps(maps: Sequence[CoordMap]):
  """ONE host read for the row counts and status words of several maps built with sync=False."""
  pend = [m for m in maps if isinstance(m.n, torch.Tensor)]
  if not pend:
    return
  vals = torch.cat([m.n for m in pend] + [m.status.to(torch.int64) for m in pend]).tolist()
  k = len(pend)
  for i, m in enumerate(pend):
    st = int(vals[k + i])
    if st:
      _lib.check_status(torch.tensor([st]), f"coordinate map (stride {m.tensor_stride})")
    m.n = int(vals[i])
    m.coords = m.coords[:m.n]
    m.status = None


finish_stride_maps = finish_maps


def kernel_map(in_cm: CoordMap, out_cm: CoordMap, ksize: int, dilation: int = 1, transposed: bool = False,
               count_pairs: bool = False, with_keys: bool = False):
  """K2 neighbour table nbr int32 [n_out, ksize^3] (see include/gclb200.h).  Forward: offsets scale with the
  input tensor stride; transposed (A7): out map is the finer one and offsets scale with ITS stride, negated."""
  K = ksize ** 3
  dev = out_cm.coords.device
  nbr = torch.empty((out_cm.n, K), dtype=torch.int32, device=dev)
  counts = torch.zeros(K, dtype=torch.int32, device=dev) if count_pairs else None
  keys = torch.empty(out_cm.n, dtype=torch.uint8, device=dev) if with_keys else None
  masks = torch.empty(out_cm.n, dtype=torch.int32, device=dev) if (with_keys and K <= 32) else None
  hist = torch.zeros((64, (out_cm.n + 1023) // 1024), dtype=torch.int32, device=dev) if with_keys else None
  step = out_cm.tensor_stride if transposed else in_cm.tensor_stride
  call("gclb_kmap_build", ptr(in_cm.table), in_cm.capacity, ptr(out_cm.coords), out_cm.n, ksize, step, dilation,
       -1 if transposed else 1, in_cm.tensor_stride, ptr(nbr), ptr(counts), ptr(keys), ptr(masks), ptr(hist), stream())
  if with_keys:
    return nbr, (keys, masks, hist)
  return (nbr, counts) if count_pairs else nbr


def kernel_map_pairs(nbr: torch.Tensor):
  """ME-style per-offset (in_idx, out_idx) lists, canonical order; returns (in_idx, out_idx, offset_ptr[K+1])."""
  n_out, K = nbr.shape
  dev = nbr.device
  lib = _lib.load()
  in_idx = torch.empty(max(n_out * K, 1), dtype=torch.int32, device=dev)
  out_idx = torch.empty(max(n_out * K, 1), dtype=torch.int32, device=dev)
  off = torch.zeros(K + 1, dtype=torch.int64, device=dev)
  ws = _workspace(lib.gclb_compact_workspace_bytes(n_out * K), dev)
  call("gclb_kmap_pairs", ptr(nbr), n_out, K, ptr(in_idx), ptr(out_idx), ptr(off), ptr(ws), stream())
  total = int(off[-1].item())
  return in_idx[:total], out_idx[:total], off


def kernel_map_sort(nbr: torch.Tensor, keys=None, copy: bool = True):
  """Group table rows by neighbour-direction pattern for the tcgen05 kernel: returns (nbr_sorted, perm, tile_mask) with
  nbr_sorted[t] = nbr[perm[t]] and tile_mask[t // 128] = bit mask of the offsets populated in that 128-row tile.
  keys: the (row_keys, row_masks) pair from kernel_map(with_keys=True).  copy=False skips the physical copy
  (nbr_sorted is None): the conv kernel then reads rows of the original table through `perm` (needs row_masks)."""
  n_out, K = nbr.shape
  ksize = round(K ** (1 / 3))
  assert ksize ** 3 == K
  lib = _lib.load()
  perm = torch.empty(n_out, dtype=torch.int32, device=nbr.device)
  row_keys, row_masks, key_hist = keys if keys is not None else (None, None, None)
  if not copy:
    assert row_masks is not None, "copy=False needs the row masks from kernel_map(with_keys=True)"
  out = torch.empty_like(nbr) if copy else None
  mask = torch.empty((n_out + 127) // 128, dtype=torch.int32, device=nbr.device) if K <= 32 else None
  ws = _workspace(lib.gclb_kmap_sort_workspace_bytes(n_out), nbr.device)
  call("gclb_kmap_sort_rows", ptr(nbr), n_out, ksize, ptr(row_keys), ptr(row_masks), ptr(key_hist), ptr(perm), ptr(out),
       ptr(mask), ptr(ws), stream())
  return out, perm, mask


class HaloMap:
  """per-tile distinct-row lists + local-index tables of a bucket-sorted same-map 3x3x3 kernel map (gclb_kmap_halo_build)"""
  __slots__ = ("slots", "tile_groups", "tile_ngroups", "counter", "status", "perm", "n_out")

  def __init__(self, slots, tile_groups, tile_ngroups, counter, status, perm, n_out):
    self.slots, self.tile_groups, self.tile_ngroups = slots, tile_groups, tile_ngroups
    self.counter, self.status, self.perm, self.n_out = counter, status, perm, n_out


def kernel_map_halo(nbr: torch.Tensor, perm: Optional[torch.Tensor]) -> HaloMap:
  """Halo staging metadata for the tcgen05 halo kernel.  nbr is the ORIGINAL table [n_out, 27]; perm the bucket order from
  kernel_map_sort (tile row t = output row perm[t]).  The record buffer is sized for the worst case (22 KB per tile; a
  LiDAR map touches ~1/6 of it), so there is no overflow path and no host synchronisation."""
  n_out, K = nbr.shape
  assert K == 27
  lib = _lib.load()
  dev = nbr.device
  tiles = (n_out + 127) // 128
  nbytes, maxg = int(lib.gclb_kmap_halo_bytes(n_out)), int(lib.gclb_kmap_halo_max_groups())
  slots = torch.empty(nbytes, dtype=torch.uint8, device=dev)
  tile_groups = torch.empty((max(tiles, 1), maxg, 2), dtype=torch.int32, device=dev)
  tile_ngroups = torch.empty(max(tiles, 1), dtype=torch.int32, device=dev)
  cs = torch.zeros(2, dtype=torch.int64, device=dev)          # [0] granules used, [1] status (low word)
  call("gclb_kmap_halo_build", ptr(nbr), n_out, ptr(perm), ptr(slots), nbytes, ptr(tile_groups), ptr(tile_ngroups),
       cs.data_ptr(), cs.data_ptr() + 8, stream())
  return HaloMap(slots, tile_groups, tile_ngroups, cs[0:1], cs[1:2], perm, n_out)


def spconv_fwd_halo(in0: torch.Tensor, Wimg: torch.Tensor, halo: HaloMap, in1: Optional[torch.Tensor] = None, scale=None,
                    shift=None, residual=None, relu=False, normalize: bool = False, out_dtype=None) -> torch.Tensor:
  """same-map 3x3x3 convolution on fp16 activations through the halo-staging tcgen05 kernel; Wimg from
  weights_to_tc(W, half=True)."""
  require_cuda(in0, Wimg, in1, scale, shift, residual)
  assert in0.dtype == torch.float16 and Wimg.dtype == torch.float16
  K, cout, cin = Wimg.shape
  c0, c1 = in0.shape[1], (in1.shape[1] if in1 is not None else 0)
  assert K == 27 and c0 + c1 == cin
  out = torch.empty((halo.n_out, cout), dtype=out_dtype or in0.dtype, device=in0.device)
  flags = int(bool(relu)) | (2 if normalize else 0) | 8 | (16 if out.dtype == torch.float16 else 0)
  flags |= 32 if f16_slab(c0, c1) == 32 else 0
  call("gclb_spconv_fwd_halo", ptr(in0), c0, ptr(in1), c1, in0.shape[0], ptr(Wimg), cout, ptr(halo.slots),
       ptr(halo.tile_groups), ptr(halo.tile_ngroups), ptr(halo.perm), ptr(scale), ptr(shift), ptr(residual), flags, ptr(out),
       halo.n_out, stream())
  return out


def spconv_fwd(in0: torch.Tensor, W: torch.Tensor, nbr: Optional[torch.Tensor], n_out: int,
               in1: Optional[torch.Tensor] = None, scale=None, shift=None, residual=None, relu=False,
               out: Optional[torch.Tensor] = None, algo: int = 0, normalize: bool = False,
               row_perm: Optional[torch.Tensor] = None, tile_mask: Optional[torch.Tensor] = None,
               nbr_is_sorted: bool = True, out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
  """K3 forward.  W is [K, Cin, Cout] (or [Cin, Cout] for the K == 1 `mm` path); with algo=2 (tcgen05) W is the
  tensor-core layout [K, Cout, Cin] from `weights_to_tc`.  fp16 activations (algo=2): pass float16 in0 / in1 / residual
  with a float16 weight image (`weights_to_tc(W, half=True)`); out_dtype picks the output storage (default: in0's)."""
  require_cuda(in0, W, nbr, in1, scale, shift, residual)
  if W.dim() == 2:
    W = W.unsqueeze(0)
  if algo == 2:
    K, cout, cin = W.shape
  else:
    K, cin, cout = W.shape
  c0 = in0.shape[1]
  c1 = in1.shape[1] if in1 is not None else 0
  assert c0 + c1 == cin, f"channel mismatch: {c0}+{c1} vs {cin}"
  half = in0.dtype == torch.float16
  assert in0.dtype in (torch.float32, torch.float16) and W.dtype == in0.dtype, (in0.dtype, W.dtype)
  assert (in1 is None or in1.dtype == in0.dtype) and (residual is None or residual.dtype == in0.dtype)
  assert not half or algo == 2, "fp16 activations exist on the tcgen05 path only"
  if nbr is not None:
    assert nbr.dtype == torch.int32 and tuple(nbr.shape) == (n_out, K)
  if out is None:
    out = torch.empty((n_out, cout), dtype=out_dtype or in0.dtype, device=in0.device)
  assert out.dtype in (torch.float32, torch.float16) and (algo == 2 or out.dtype == torch.float32)
  flags = int(bool(relu)) | (2 if normalize else 0) | (4 if (row_perm is not None and nbr_is_sorted) else 0)
  flags |= (8 if half else 0) | (16 if out.dtype == torch.float16 else 0)
  flags |= 32 if (half and f16_slab(c0, c1) == 32) else 0
  call("gclb_spconv_fwd", ptr(in0), c0, ptr(in1), c1, in0.shape[0], ptr(W.contiguous()), K, cout, ptr(nbr),
       ptr(row_perm), ptr(tile_mask), ptr(scale), ptr(shift), ptr(residual), flags, ptr(out), n_out, algo, stream())
  return out


def spconv_fwd_probe(x: torch.Tensor, W: torch.Tensor, cm: CoordMap, ksize: int, dilation: int = 1, scale=None,
                     shift=None, residual=None, relu=False, emit_k3: bool = False, out_dtype=torch.float32):
  """stride-1 convolution of a narrow input (cin <= 4) with the kernel map fused into the kernel (hash probes instead
  of a neighbour table): conv1 of the ResUNet.  emit_k3: also return the stride-1 3x3x3 kernel map of `cm` that the
  inner probes amount to, as (nbr [n, 27], (row_keys, row_masks, key_hist)) -- the same objects kernel_map(cm, cm, 3,
  with_keys=True) builds with a separate pass."""
  require_cuda(x, W)
  K, cin, cout = W.shape
  assert K == ksize ** 3 and x.shape[1] == cin and x.shape[0] == cm.n
  dev = x.device
  out = torch.empty((cm.n, cout), dtype=out_dtype, device=dev)
  nbr = keys = masks = hist = None
  if emit_k3:
    assert ksize % 2 == 1 and ksize >= 3 and dilation == 1
    nbr = torch.empty((cm.n, 27), dtype=torch.int32, device=dev)
    keys = torch.empty(cm.n, dtype=torch.uint8, device=dev)
    masks = torch.empty(cm.n, dtype=torch.int32, device=dev)
    hist = torch.zeros((64, (cm.n + 1023) // 1024), dtype=torch.int32, device=dev)
  call("gclb_spconv_fwd_probe", ptr(x.contiguous()), cin, ptr(W.contiguous()), ksize, cout, ptr(cm.table), cm.capacity,
       ptr(cm.coords), cm.n, cm.tensor_stride, dilation, ptr(scale), ptr(shift), ptr(residual),
       int(bool(relu)) | (16 if out_dtype == torch.float16 else 0), ptr(out), ptr(nbr), ptr(keys), ptr(masks), ptr(hist),
       stream())
  if emit_k3:
    return out, (nbr, (keys, masks, hist))
  return out


def f16_slab(c0: int, c1: int = 0) -> int:
  """channels per gathered fp16 row: 64 (128-byte rows) when every source width is a multiple of 64, else 32 (64-byte rows)"""
  return 64 if (c0 % 64 == 0 and c1 % 64 == 0) else 32


def weights_to_tc(W: torch.Tensor, half: bool = False, c0: Optional[int] = None) -> torch.Tensor:
  """[K, Cin, Cout] (or [Cin, Cout]) -> tensor-core image (shape [K, Cout, Cin], pre-swizzled per 32-channel slab,
  rounded to tf32; half=True: per 64- or 32-channel slab (see f16_slab; c0 = width of the first source of a two-source
  layer), rounded to fp16); done once per layer."""
  require_cuda(W)
  W3 = (W if W.dim() == 3 else W.unsqueeze(0)).contiguous().float()
  K, cin, cout = W3.shape
  Wt = torch.empty((K, cout, cin), dtype=torch.float16 if half else torch.float32, device=W.device)
  if half:
    c0 = cin if c0 is None else c0
    call("gclb_weights_to_tc_f16", ptr(W3), K, cin, cout, f16_slab(c0, cin - c0), ptr(Wt), stream())
  else:
    call("gclb_weights_to_tc", ptr(W3), K, cin, cout, ptr(Wt), stream())
  return Wt


def weights_to_tc_dgrad(W: torch.Tensor, flip: bool) -> torch.Tensor:
  """tf32 tensor-core image of the data-gradient convolution's weights, in one launch from the forward weights
  [K, Cin, Cout] (or [Cin, Cout]): equals weights_to_tc((W.flip(0) if flip else W).transpose(1, 2))."""
  require_cuda(W)
  W3 = (W if W.dim() == 3 else W.unsqueeze(0)).contiguous().float()
  K, cin, cout = W3.shape
  Wt = torch.empty((K, cin, cout), dtype=torch.float32, device=W.device)     # image of [K, Cout -> Cin]
  call("gclb_weights_to_tc_dgrad", ptr(W3), K, cin, cout, int(bool(flip)), ptr(Wt), stream())
  return Wt


def tc_supported(c0: int, c1: int, cout: int, K: int, half: bool = False) -> bool:
  kch = 32
  return (_lib.load().gclb_has_tcgen05() == 1 and c0 % kch == 0 and c1 % kch == 0 and c0 >= kch
          and cout in (32, 64, 128, 256) and K in (1, 27))


def spconv_wgrad(x: torch.Tensor, gout: torch.Tensor, nbr: Optional[torch.Tensor], K: int) -> torch.Tensor:
  cin, cout = x.shape[1], gout.shape[1]
  gW = torch.zeros((K, cin, cout), dtype=torch.float32, device=x.device)
  call("gclb_spconv_wgrad", ptr(x.contiguous()), cin, x.shape[0], ptr(gout.contiguous()), cout, gout.shape[0],
       ptr(nbr), K, ptr(gW), stream())
  return gW


def wgrad_tc_supported(cin: int, cout: int) -> bool:
  return cin % 32 == 0 and cout % 32 == 0 and 32 <= cin <= 256 and 32 <= cout <= 256


def spconv_wgrad_tc(x: torch.Tensor, gout: torch.Tensor, nbr_sorted: Optional[torch.Tensor], K: int,
                    row_perm: Optional[torch.Tensor] = None, tile_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
  """Weight gradient on tcgen05 (tf32 operands) over the row-bucketed forward table; see gclb_spconv_wgrad_tc."""
  require_cuda(x, gout, nbr_sorted, row_perm, tile_mask)
  cin, cout = x.shape[1], gout.shape[1]
  if nbr_sorted is not None:
    assert nbr_sorted.dtype == torch.int32 and tuple(nbr_sorted.shape) == (gout.shape[0], K)
  gW = torch.zeros((K, cin, cout), dtype=torch.float32, device=x.device)
  call("gclb_spconv_wgrad_tc", ptr(x.contiguous()), cin, x.shape[0], ptr(gout.contiguous()), cout, gout.shape[0],
       ptr(nbr_sorted), ptr(row_perm), ptr(tile_mask), K, ptr(gW), stream())
  return gW


def pointwise_tail(in0, in1, W1, W2, bias, normalize=True):
  n = in0.shape[0]
  c0, c1 = in0.shape[1], (in1.shape[1] if in1 is not None else 0)
  cmid, cout = W1.shape[-1], W2.shape[-1]
  out = torch.empty((n, cout), dtype=torch.float32, device=in0.device)
  call("gclb_pointwise_tail", ptr(in0), c0, ptr(in1), c1, n, ptr(W1.contiguous()), cmid, ptr(W2.contiguous()),
       ptr(bias), cout, int(bool(normalize)), ptr(out), stream())
  return out


def affine_act(x, scale=None, shift=None, residual=None, relu=False, out=None):
  x = x.contiguous()
  if out is None:
    out = torch.empty_like(x)
  call("gclb_affine_act", ptr(x), x.shape[0], x.shape[1], ptr(scale), ptr(shift), ptr(residual), int(bool(relu)),
       ptr(out), stream())
  return out


def subsample(cm: CoordMap, n_clouds: int, S: int, groups: int = 2, seed: int = 0):
  """Per-cloud random subsample (without replacement) of a batched map's rows, on the device, no host sync.
  Returns (cloud_ptr int64 [n_clouds+1], sel_ptr int64 [groups, n_clouds/groups+1], sel int64 [groups, n_seg*cap], cap)."""
  dev = cm.coords.device
  n_dev = cm.n if isinstance(cm.n, torch.Tensor) else None
  n_rows = cm.coords.shape[0]
  cap = S if (S > 0 and S < n_rows) else n_rows
  n_seg = n_clouds // groups
  cloud_ptr = torch.empty(n_clouds + 1, dtype=torch.int64, device=dev)
  sel_ptr = torch.empty((groups, n_seg + 1), dtype=torch.int64, device=dev)
  sel = torch.empty((groups, max(n_seg * cap, 1)), dtype=torch.int64, device=dev)
  call("gclb_subsample", ptr(cm.coords), ptr(n_dev), n_rows, n_clouds, S, groups, seed & 0xFFFFFFFFFFFFFFFF,
       ptr(cloud_ptr), ptr(sel_ptr), ptr(sel), stream())
  return cloud_ptr, sel_ptr, sel, cap


def nn_search(A: torch.Tensor, B: torch.Tensor, a_ptr=None, b_ptr=None, both=True, algo: int = 0, a_rows=None,
              b_rows=None, max_n: Optional[int] = None, max_m: Optional[int] = None):
  """K4.  Returns (idx01 int64, d01 float32, idx10, d10, a_ptr_dev, b_ptr_dev, workspace): squared-L2 nearest neighbours
  in both directions; indices are segment-local when a_ptr/b_ptr (int64 [n_pairs+1]) are given.  With device-resident
  a_ptr/b_ptr the caller passes max_n/max_m (upper bounds of the segment lengths) and nothing is read back.
  a_rows/b_rows: optional int64 row indirection (segment row i reads A[a_rows[i]])."""
  require_cuda(A, B, a_rows, b_rows)
  A, B = A.contiguous(), B.contiguous()
  assert A.dtype == torch.float32 and B.dtype == torch.float32 and A.shape[1] == B.shape[1]
  dev = A.device
  Cd = A.shape[1]
  if isinstance(a_ptr, torch.Tensor) and a_ptr.is_cuda:
    assert max_n is not None and max_m is not None, "device-resident segment pointers need max_n / max_m"
    a_dev, b_dev = a_ptr.contiguous(), b_ptr.contiguous()
    n_pairs = a_dev.numel() - 1
    N, M = n_pairs * max_n, n_pairs * max_m
  else:
    N = A.shape[0] if a_rows is None else a_rows.numel()
    M = B.shape[0] if b_rows is None else b_rows.numel()
    if a_ptr is None:
      a_host, b_host = [0, N], [0, M]
    else:
      a_host = a_ptr.tolist() if isinstance(a_ptr, torch.Tensor) else list(a_ptr)
      b_host = b_ptr.tolist() if isinstance(b_ptr, torch.Tensor) else list(b_ptr)
    n_pairs = len(a_host) - 1
    max_n = max(a_host[i + 1] - a_host[i] for i in range(n_pairs))
    max_m = max(b_host[i + 1] - b_host[i] for i in range(n_pairs))
    a_dev = torch.tensor(a_host, dtype=torch.int64, device=dev)
    b_dev = torch.tensor(b_host, dtype=torch.int64, device=dev)
  lib = _lib.load()
  idx01 = torch.empty(N, dtype=torch.int64, device=dev)
  d01 = torch.empty(N, dtype=torch.float32, device=dev)
  idx10 = torch.empty(M, dtype=torch.int64, device=dev) if both else None
  d10 = torch.empty(M, dtype=torch.float32, device=dev) if both else None
  ws = _workspace(lib.gclb_nn_workspace_bytes(N, M, n_pairs, max_n, max_m), dev)
  call("gclb_nn", ptr(A), ptr(B), Cd, ptr(a_dev), ptr(b_dev), n_pairs, ptr(a_rows), ptr(b_rows), N, M, max_n, max_m,
       ptr(idx01), ptr(d01), ptr(idx10), ptr(d10), algo, ptr(ws), stream())
  return idx01, d01, idx10, d10, a_dev, b_dev, ws


def mutual_filter(idx01, idx10, a_dev, b_dev, ws=None):
  dev = idx01.device
  n_pairs = a_dev.numel() - 1
  N = idx01.shape[0]
  lib = _lib.load()
  pairs = torch.empty((max(N, 1), 2), dtype=torch.int64, device=dev)
  pair_ptr = torch.zeros(n_pairs + 1, dtype=torch.int64, device=dev)
  if ws is None:
    ws = _workspace(lib.gclb_compact_workspace_bytes(N), dev)
  call("gclb_mutual_filter", ptr(idx01), ptr(idx10), ptr(a_dev), ptr(b_dev), n_pairs, N, ptr(pairs), ptr(pair_ptr),
       ptr(ws), stream())
  return pairs, pair_ptr
