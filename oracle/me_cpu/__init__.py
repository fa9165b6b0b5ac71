"""CPU ORACLE (test infrastructure, NOT product code) -- a MinkowskiEngine-shaped restatement.

PARITY STATUS: **parity unpinned at the MinkowskiEngine boundary**.  The reference (liuQuan98/GCL)
depends on an un-vendored, un-pinned `MinkowskiEngine` (requirements.txt:8, README.md:77 "v0.5 or
higher", README.md:130 installs git HEAD == v0.5.4) whose sources are absent from /root/reference and
which ships no golden vectors.  This module restates ME 0.5.x's *documented* operator semantics
 in numpy/torch-CPU.  It is anchored two ways instead:
  (1) independent known-answer tests against dense `torch.nn.functional.conv3d` / `conv_transpose3d`
      (tests/test_oracle_dense_equiv.py), and
  (2) the reference's own call sites: `/root/reference/model/resunet.py` imports and runs unmodified on
      top of this module (tests/test_reference_live.py; fixtures in tests/golden/).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this package.  The product (`gcl_b200`) never does.

Operator call sites this module satisfies (reference file:line):
  ME.MinkowskiNetwork                model/resunet.py:10,31
  ME.MinkowskiConvolution            model/resunet.py:38-45,62-95,153-171; model/residual_block.py:23-33
  ME.MinkowskiConvolutionTranspose   model/resunet.py:101-134
  ME.MinkowskiBatchNorm / InstanceNorm   model/common.py:6,8
  MEF.relu                           model/resunet.py:177-223; model/residual_block.py:42,51
  ME.cat                             model/resunet.py:203,210,217
  ME.SparseTensor                    scripts/test_kitti.py:143-148; model/resunet.py:227-230
  ME.utils.sparse_quantize           lib/complement_data_loader.py:788-789; util/misc.py:118
  ME.utils.sparse_collate / batched_coordinates   lib/complement_data_loader.py:1310-1311; util/misc.py:120
"""
from __future__ import annotations

import itertools
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

__version__ = "0.5.4-oracle"

# --------------------------------------------------------------------------------------------------
# integer side: keys, hashing (as sorted lookup), strided maps, kernel maps        (Appendix A1-A7)
# --------------------------------------------------------------------------------------------------
_B = 1 << 18  # per-axis radix of the oracle's sortable key: ((b*R + x')*R + y')*R + z'
_O = 1 << 17


def pack_rows(c: np.ndarray) -> np.ndarray:
  """[N,4] int (b,x,y,z) -> int64 scalar keys, collision free for |x|,|y|,|z| < 2^17 and 0 <= b < 512."""
  c = np.asarray(c, dtype=np.int64)
  assert c.ndim == 2 and c.shape[1] == 4
  if c.size:
    assert np.all(np.abs(c[:, 1:]) < _O), "oracle key range exceeded"
    assert np.all((c[:, 0] >= 0) & (c[:, 0] < 512)), "oracle batch range exceeded"
  return ((c[:, 0] * _B + (c[:, 1] + _O)) * _B + (c[:, 2] + _O)) * _B + (c[:, 3] + _O)


def first_occurrence_unique(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
  """A1: rows inserted in input order, first occurrence wins.
  Returns (unique_map ascending original indices, inverse_map: row -> position in unique_map)."""
  if len(keys) == 0:
    return np.zeros(0, np.int64), np.zeros(0, np.int64)
  _, first, inv = np.unique(keys, return_index=True, return_inverse=True)
  order = np.argsort(first, kind="stable")          # rank unique keys by first appearance
  rank = np.empty_like(order)
  rank[order] = np.arange(len(order))
  return first[order].astype(np.int64), rank[inv].astype(np.int64)


class _Lookup:
  """coord -> row lookup for one coordinate map (stands in for ME's hash table)."""

  def __init__(self, coords: np.ndarray):
    keys = pack_rows(coords)
    self.order = np.argsort(keys, kind="stable")
    self.sorted = keys[self.order]
    assert len(np.unique(keys)) == len(keys), "coordinate map rows must be unique"

  def query(self, q: np.ndarray) -> np.ndarray:
    """[Q,4] -> int64 [Q] row index or -1."""
    if len(self.sorted) == 0 or len(q) == 0:
      return np.full(len(q), -1, np.int64)
    k = pack_rows(q)
    pos = np.searchsorted(self.sorted, k)
    pos_c = np.minimum(pos, len(self.sorted) - 1)
    hit = self.sorted[pos_c] == k
    return np.where(hit, self.order[pos_c], -1).astype(np.int64)


def kernel_offsets(kernel_size: int, tensor_stride: int, dilation: int = 1) -> np.ndarray:
  """A5: hyper-cube region, x fastest: k = ix + K*iy + K^2*iz.  odd: (i - K//2)*d*ts ; even: i*d*ts."""
  K = kernel_size
  r = np.arange(K) - (K // 2 if K % 2 == 1 else 0)
  offs = np.zeros((K ** 3, 3), np.int64)
  for k, (iz, iy, ix) in enumerate(itertools.product(range(K), repeat=3)):
    offs[k] = (r[ix], r[iy], r[iz])
  return offs * dilation * tensor_stride


def stride_coords(coords: np.ndarray, new_stride: int) -> np.ndarray:
  """A4: unique(floor(c / S) * S) per spatial axis, batch preserved; canonical order = first appearance
  scanning parent rows in order."""
  c = np.asarray(coords, np.int64).copy()
  c[:, 1:] = np.floor_divide(c[:, 1:], new_stride) * new_stride
  um, _ = first_occurrence_unique(pack_rows(c))
  return c[um].astype(np.int32)


def build_neighbor_table(in_coords, out_coords, offsets) -> np.ndarray:
  """nbr[o, k] = row i of in_coords with coord_in[i] == coord_out[o] + off_k, else -1   (A6)."""
  lk = _Lookup(in_coords)
  out = np.asarray(out_coords, np.int64)
  nbr = np.full((len(out), len(offsets)), -1, np.int64)
  for k, off in enumerate(offsets):
    q = out.copy()
    q[:, 1:] += off
    nbr[:, k] = lk.query(q)
  return nbr


def neighbor_table_to_pairs(nbr: np.ndarray) -> List[Tuple[np.ndarray, np.ndarray]]:
  """Per-offset (in_idx, out_idx) lists, canonical order = ascending out row."""
  out = []
  for k in range(nbr.shape[1]):
    o = np.nonzero(nbr[:, k] >= 0)[0]
    out.append((nbr[o, k].astype(np.int64), o.astype(np.int64)))
  return out


# --------------------------------------------------------------------------------------------------
# coordinate manager / keys / sparse tensor                                         (Appendix A3, A9)
# --------------------------------------------------------------------------------------------------
class CoordinateMapKey:
  def __init__(self, tensor_stride: Sequence[int], string_id: str = ""):
    self.tensor_stride = tuple(int(s) for s in tensor_stride)
    self.string_id = string_id

  def get_tensor_stride(self):
    return list(self.tensor_stride)

  def get_key(self):
    return (list(self.tensor_stride), self.string_id)

  def __eq__(self, o):
    return isinstance(o, CoordinateMapKey) and self.get_key() == o.get_key()

  def __hash__(self):
    return hash((self.tensor_stride, self.string_id))

  def __repr__(self):
    return f"CoordinateMapKey(stride={list(self.tensor_stride)}, id='{self.string_id}')"


class CoordinateManager:
  def __init__(self, D: int = 3):
    self.D = D
    self.coords: Dict[CoordinateMapKey, np.ndarray] = {}
    self.kmaps: Dict[tuple, np.ndarray] = {}
    self.stats = {"kmap_builds": 0, "stride_builds": 0}

  def insert(self, coords: np.ndarray, tensor_stride=(1, 1, 1)) -> CoordinateMapKey:
    key = CoordinateMapKey(tensor_stride, "")
    assert key not in self.coords
    self.coords[key] = np.ascontiguousarray(coords, dtype=np.int32)
    return key

  def get_coordinates(self, key) -> torch.Tensor:
    return torch.from_numpy(self.coords[key].copy())

  def size(self, key) -> int:
    return len(self.coords[key])

  def stride(self, key: CoordinateMapKey, stride: int) -> CoordinateMapKey:
    new = CoordinateMapKey([s * stride for s in key.tensor_stride], "")
    if new not in self.coords:
      self.coords[new] = stride_coords(self.coords[key], new.tensor_stride[0])
      self.stats["stride_builds"] += 1
    return new

  def neighbor_table(self, in_key, out_key, kernel_size, dilation, transposed=False) -> np.ndarray:
    """Forward: nbr[o,k] over in_key rows with offsets scaled by in tensor stride.
    Transposed (A7): out map is the finer one; pairs are the forward strided map (fine->coarse) swapped,
    i.e. out_fine[f] += in_coarse[c] W[k]  iff  coord[f] == coord[c] + off_k(fine stride)."""
    ck = (in_key, out_key, kernel_size, dilation, transposed)
    if ck not in self.kmaps:
      if not transposed:
        offs = kernel_offsets(kernel_size, in_key.tensor_stride[0], dilation)
        nbr = build_neighbor_table(self.coords[in_key], self.coords[out_key], offs)
      else:
        offs = kernel_offsets(kernel_size, out_key.tensor_stride[0], dilation)
        # coarse c pairs with fine f = c + off_k  <=>  c = f - off_k
        nbr = build_neighbor_table(self.coords[in_key], self.coords[out_key], -offs)
      self.kmaps[ck] = nbr
      self.stats["kmap_builds"] += 1
    return self.kmaps[ck]


class SparseTensor:
  def __init__(self, features, coordinates=None, *, tensor_stride=1, coordinate_map_key=None,
               coordinate_manager=None, quantization_mode=None, minkowski_algorithm=None,
               requires_grad=None, device=None):
    assert isinstance(features, torch.Tensor)
    if device is not None:
      features = features.to(device)
    if coordinate_map_key is None:
      assert coordinates is not None
      c = coordinates.detach().cpu().numpy() if isinstance(coordinates, torch.Tensor) else np.asarray(coordinates)
      assert c.ndim == 2 and c.shape[1] == 4 and len(c) == len(features)
      if coordinate_manager is None:
        coordinate_manager = CoordinateManager(D=c.shape[1] - 1)
      ts = (tensor_stride,) * 3 if isinstance(tensor_stride, int) else tuple(tensor_stride)
      # A3: the reference always feeds unique rows (sparse_quantize upstream); duplicates are an error here.
      coordinate_map_key = coordinate_manager.insert(c.astype(np.int32), ts)
    else:
      assert coordinate_manager is not None
    self._F = features
    self.coordinate_map_key = coordinate_map_key
    self._manager = coordinate_manager
    if requires_grad is not None:
      self._F.requires_grad_(requires_grad)

  # -- accessors the reference uses
  @property
  def F(self):
    return self._F

  @property
  def feats(self):
    return self._F

  @property
  def C(self):
    return self._manager.get_coordinates(self.coordinate_map_key)

  @property
  def coordinates(self):
    return self.C

  @property
  def coordinate_manager(self):
    return self._manager

  @property
  def tensor_stride(self):
    return self.coordinate_map_key.get_tensor_stride()

  @property
  def D(self):
    return self._manager.D

  @property
  def device(self):
    return self._F.device

  @property
  def dtype(self):
    return self._F.dtype

  @property
  def shape(self):
    return self._F.shape

  def size(self):
    return self._F.size()

  def __len__(self):
    return len(self._F)

  def _check(self, other):
    if not isinstance(other, SparseTensor):
      raise TypeError("operand must be a SparseTensor")
    if other._manager is not self._manager or other.coordinate_map_key != self.coordinate_map_key:
      raise ValueError("SparseTensors must share coordinate manager and coordinate map key")

  def _like(self, feats):
    return SparseTensor(feats, coordinate_map_key=self.coordinate_map_key, coordinate_manager=self._manager)

  def __iadd__(self, other):
    self._check(other)
    self._F = self._F + other._F   # functional form: autograd-safe, same values as ME's in-place add
    return self

  def __add__(self, other):
    self._check(other)
    return self._like(self._F + other._F)

  def __repr__(self):
    return f"SparseTensor(N={len(self)}, C={self._F.shape[1]}, {self.coordinate_map_key})"


def cat(*tensors):
  if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
    tensors = tuple(tensors[0])
  for t in tensors[1:]:
    tensors[0]._check(t)
  return tensors[0]._like(torch.cat([t.F for t in tensors], dim=1))


# --------------------------------------------------------------------------------------------------
# layers                                                                            (Appendix A6-A10)
# --------------------------------------------------------------------------------------------------
class MinkowskiNetwork(nn.Module):
  def __init__(self, D):
    super().__init__()
    self.D = D


def sparse_conv_reference(x: torch.Tensor, W: torch.Tensor, nbr: np.ndarray, n_out: int) -> torch.Tensor:
  """A6: out[o] = sum_k in[nbr[o,k]] @ W[k]   -- ME's CPU algorithm: per offset gather -> GEMM -> scatter-add."""
  out = x.new_zeros((n_out, W.shape[-1]))
  for k, (i_idx, o_idx) in enumerate(neighbor_table_to_pairs(nbr)):
    if len(i_idx) == 0:
      continue
    out.index_add_(0, torch.from_numpy(o_idx), x.index_select(0, torch.from_numpy(i_idx)) @ W[k])
  return out


class _ConvBase(nn.Module):
  TRANSPOSED = False

  def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
               kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
    super().__init__()
    assert dimension == 3, "oracle covers D=3 only"
    assert kernel_generator is None and not expand_coordinates
    assert isinstance(kernel_size, int) and kernel_size >= 1
    self.in_channels, self.out_channels = in_channels, out_channels
    self.kernel_size, self.stride, self.dilation, self.dimension = kernel_size, stride, dilation, dimension
    self.kernel_volume = kernel_size ** 3
    self.use_mm = (self.kernel_volume == 1 and stride == 1)
    shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
    self.kernel = nn.Parameter(torch.empty(shape))
    self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
    self.reset_parameters()

  def reset_parameters(self):
    with torch.no_grad():
      n = self.in_channels * self.kernel_volume
      stdv = 1.0 / np.sqrt(n)
      self.kernel.uniform_(-stdv, stdv)
      if self.bias is not None:
        self.bias.uniform_(-stdv, stdv)

  def forward(self, x: SparseTensor) -> SparseTensor:
    mgr, in_key = x._manager, x.coordinate_map_key
    if self.use_mm:
      out = x.F @ self.kernel
      out_key = in_key
    else:
      if not self.TRANSPOSED:
        out_key = mgr.stride(in_key, self.stride) if self.stride > 1 else in_key
      else:
        ts = [s // self.stride for s in in_key.tensor_stride]
        out_key = CoordinateMapKey(ts, "")
        if out_key not in mgr.coords:
          raise ValueError("transposed conv needs an existing coordinate map at the output stride (A7)")
      nbr = mgr.neighbor_table(in_key, out_key, self.kernel_size, self.dilation, self.TRANSPOSED)
      out = sparse_conv_reference(x.F, self.kernel, nbr, mgr.size(out_key))
    if self.bias is not None:
      out = out + self.bias
    return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)

  def extra_repr(self):
    return (f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}, "
            f"stride={self.stride}, dilation={self.dilation}")


class MinkowskiConvolution(_ConvBase):
  TRANSPOSED = False


class MinkowskiConvolutionTranspose(_ConvBase):
  TRANSPOSED = True


class MinkowskiBatchNorm(nn.Module):
  """A8: BatchNorm1d on .F (child module name `bn` is part of the checkpoint contract, A10)."""

  def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
    super().__init__()
    self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                             track_running_stats=track_running_stats)

  def forward(self, x: SparseTensor) -> SparseTensor:
    return x._like(self.bn(x.F))


class MinkowskiInstanceNorm(nn.Module):
  def __init__(self, num_features, dimension=-1):
    super().__init__()
    self.num_features = num_features
    self.weight = nn.Parameter(torch.ones(1, num_features))
    self.bias = nn.Parameter(torch.zeros(1, num_features))

  def forward(self, x):
    raise NotImplementedError("InstanceNorm is outside the ResUNetBN2C hot path (SURVEY.md section 2, row 3)")


class MinkowskiReLU(nn.Module):
  def __init__(self, inplace=False):
    super().__init__()

  def forward(self, x):
    return x._like(torch.relu(x.F))


from . import MinkowskiFunctional  # noqa: E402
from . import utils  # noqa: E402
from .utils import sparse_quantize, sparse_collate, batched_coordinates  # noqa: E402,F401
