"""`MinkowskiEngine`-shaped operator module backed by libgclb200 (hand-written sm_100a CUDA).

Drop-in for the subset of MinkowskiEngine 0.5.x that the GCL/FCGF hot path consumes:
/root/reference/model/resunet.py, model/residual_block.py and model/common.py import and run unmodified on
top of this module after `gcl_b200.install_as_minkowski_engine()`.

  ME.MinkowskiNetwork / MinkowskiConvolution / MinkowskiConvolutionTranspose / MinkowskiBatchNorm /
  MinkowskiInstanceNorm (constructible; raises on call) / MinkowskiReLU / MinkowskiFunctional.relu / cat /
  SparseTensor / CoordinateManager / CoordinateMapKey / utils.{sparse_quantize,sparse_collate,batched_coordinates}

Everything sparse (hashing, strided maps, kernel maps, convolution forward/backward) runs in libgclb200; there is
no CPU path: CPU feature tensors raise.  Semantics follow SURVEY.md Appendix A (A3-A10).
"""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .._lib import GclbError

__version__ = "0.5.4+gclb200"


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
  """Owns the device hash tables, strided coordinate maps and the kernel-map cache shared by every
  SparseTensor derived from one input (ME's CoordinateManager; A3, A4, A6)."""

  def __init__(self, D: int = 3):
    self.D = D
    self.maps: Dict[CoordinateMapKey, ops.CoordMap] = {}
    self.kmaps: Dict[tuple, torch.Tensor] = {}
    self.kmaps_sorted: Dict[tuple, tuple] = {}
    self.stats = {"kmap_builds": 0, "stride_builds": 0}

  def insert(self, coords4: torch.Tensor, tensor_stride=(1, 1, 1)) -> CoordinateMapKey:
    key = CoordinateMapKey(tensor_stride, "")
    if key in self.maps:
      raise GclbError("coordinate map for this tensor stride already exists in the manager")
    assert len(set(key.tensor_stride)) == 1, "isotropic tensor strides only"
    self.maps[key] = ops.hash_build(coords4, tensor_stride=key.tensor_stride[0])
    return key

  def adopt(self, cm: ops.CoordMap) -> CoordinateMapKey:
    """Register an already hashed map (e.g. straight from gcl_b200.voxelize) without re-hashing."""
    key = CoordinateMapKey((cm.tensor_stride,) * 3, "")
    self.maps[key] = cm
    return key

  def get_coordinates(self, key) -> torch.Tensor:
    return self.maps[key].coords

  def size(self, key) -> int:
    return self.maps[key].n

  def stride(self, key: CoordinateMapKey, stride: int) -> CoordinateMapKey:
    new = CoordinateMapKey([s * stride for s in key.tensor_stride], "")
    if new not in self.maps:
      self.maps[new] = ops.stride_map(self.maps[key], stride)
      self.stats["stride_builds"] += 1
    return new

  def kernel_map(self, in_key, out_key, kernel_size: int, dilation: int, transposed: bool) -> torch.Tensor:
    ck = (in_key, out_key, kernel_size, dilation, transposed)
    if ck not in self.kmaps:
      self.kmaps[ck] = ops.kernel_map(self.maps[in_key], self.maps[out_key], kernel_size, dilation, transposed)
      self.stats["kmap_builds"] += 1
    return self.kmaps[ck]

  def kernel_map_sorted(self, in_key, out_key, kernel_size: int, dilation: int, transposed: bool):
    """row-bucketed copy (sorted table, perm, tile masks) of a cached kernel map for the tcgen05 kernel"""
    ck = (in_key, out_key, kernel_size, dilation, transposed)
    if ck not in self.kmaps_sorted:
      self.kmaps_sorted[ck] = ops.kernel_map_sort(self.kernel_map(in_key, out_key, kernel_size, dilation, transposed))
    return self.kmaps_sorted[ck]

  def kernel_map_pairs(self, in_key, out_key, kernel_size, dilation=1, transposed=False):
    """ME-style kernel map {k: (in_idx, out_idx)} in canonical order (debug / parity API)."""
    nbr = self.kernel_map(in_key, out_key, kernel_size, dilation, transposed)
    i, o, off = ops.kernel_map_pairs(nbr)
    off = off.tolist()
    return {k: (i[off[k]:off[k + 1]], o[off[k]:off[k + 1]]) for k in range(len(off) - 1)}


class SparseTensor:
  def __init__(self, features, coordinates=None, *, tensor_stride=1, coordinate_map_key=None,
               coordinate_manager=None, quantization_mode=None, minkowski_algorithm=None, requires_grad=None,
               device=None):
    if not isinstance(features, torch.Tensor):
      raise TypeError("features must be a torch.Tensor")
    if device is not None:
      features = features.to(device)
    if not features.is_cuda:
      raise GclbError("gcl_b200.MinkowskiEngine runs on CUDA only (no CPU fallback): move features to a CUDA device")
    if coordinate_map_key is None:
      if coordinates is None:
        raise ValueError("either coordinates or coordinate_map_key must be given")
      c = torch.as_tensor(coordinates)
      if c.dim() != 2 or c.shape[1] != 4 or len(c) != len(features):
        raise ValueError("coordinates must be [N, 4] (batch, x, y, z) with one row per feature row")
      c = c.to(device=features.device, dtype=torch.int32)
      if coordinate_manager is None:
        coordinate_manager = CoordinateManager(D=3)
      ts = (tensor_stride,) * 3 if isinstance(tensor_stride, int) else tuple(tensor_stride)
      coordinate_map_key = coordinate_manager.insert(c, ts)   # raises on duplicate rows (A3)
    elif coordinate_manager is None:
      raise ValueError("coordinate_manager is required together with coordinate_map_key")
    self._F = features
    self.coordinate_map_key = coordinate_map_key
    self._manager = coordinate_manager
    if requires_grad is not None:
      self._F.requires_grad_(requires_grad)

  @property
  def F(self):
    return self._F

  feats = features = F

  @property
  def C(self):
    return self._manager.get_coordinates(self.coordinate_map_key)

  coordinates = C

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
    if torch.is_grad_enabled() and (self._F.requires_grad or other._F.requires_grad):
      self._F = self._F + other._F
    else:
      self._F += other._F
    return self

  def __add__(self, other):
    self._check(other)
    return self._like(self._F + other._F)

  def __repr__(self):
    return f"SparseTensor(N={len(self)}, C={self._F.shape[1]}, {self.coordinate_map_key}, device={self.device})"


def cat(*tensors):
  if len(tensors) == 1 and isinstance(tensors[0], (list, tuple)):
    tensors = tuple(tensors[0])
  for t in tensors[1:]:
    tensors[0]._check(t)
  return tensors[0]._like(torch.cat([t.F for t in tensors], dim=1))


# Which convolution kernel the module path uses.  Training (autograd active) always runs the exact-fp32 kernels so that
# gradients match the reference's fp32 arithmetic; inference uses the tcgen05 kind::tf32 kernel where it applies.
_CONV_ALGO = {"inference": "auto", "training": "fp32"}


def set_training_conv_algo(name: str):
  """'fp32' (default: exact-fp32 forward / dgrad / wgrad, gradients match the reference's arithmetic to ~1e-5) or 'tf32'
  (forward, dgrad and wgrad on the tcgen05 kind::tf32 kernels where the layer shape allows, ~1e-3)."""
  if name not in ("fp32", "tf32"):
    raise ValueError("algo must be 'fp32' or 'tf32'")
  _CONV_ALGO["training"] = name


def set_inference_conv_algo(name: str):
  """'auto' (tcgen05 kind::tf32 where supported, else fp32) or 'fp32' (exact-fp32 CUDA-core kernels everywhere)."""
  if name not in ("auto", "fp32"):
    raise ValueError("algo must be 'auto' or 'fp32'")
  _CONV_ALGO["inference"] = name


class MinkowskiNetwork(nn.Module):
  def __init__(self, D):
    super().__init__()
    self.D = D


class _SparseConvFn(torch.autograd.Function):
  """out = conv(x, W) over a neighbour table; backward = dgrad (same kernel, dual table / transposed weights) + wgrad.
  `tc_fwd` / `tc_bwd` are row-bucketed tables (sorted, perm, masks) when the tcgen05 kernel should be used."""

  @staticmethod
  def forward(ctx, x, W, nbr_fwd, nbr_bwd, n_out, mirror, tc_fwd, tc_bwd):
    ctx.save_for_backward(x, W)
    ctx.nbr_fwd, ctx.nbr_bwd, ctx.mirror, ctx.n_in, ctx.tc_bwd = nbr_fwd, nbr_bwd, mirror, x.shape[0], tc_bwd
    ctx.tc_fwd = tc_fwd
    xc = x.contiguous()
    if tc_fwd is not None:
      Wt = ops.weights_to_tc(W.detach())
      if tc_fwd == "mm":
        return ops.spconv_fwd(xc, Wt, None, n_out, algo=2)
      srt, perm, mask = tc_fwd
      return ops.spconv_fwd(xc, Wt, srt, n_out, algo=2, row_perm=perm, tile_mask=mask)
    return ops.spconv_fwd(xc, W, nbr_fwd, n_out)

  @staticmethod
  def backward(ctx, gout):
    x, W = ctx.saved_tensors
    gout = gout.contiguous()
    gx = gW = None
    W3 = W if W.dim() == 3 else W.unsqueeze(0)
    if ctx.needs_input_grad[0]:
      # mirror: stride-1 odd kernels reuse the forward table with offsets reversed (k -> K-1-k)
      if ctx.tc_bwd is not None:
        Wt = ops.weights_to_tc_dgrad(W3, ctx.mirror)        # flip + transpose + image in one launch
        if ctx.tc_bwd == "mm":
          gx = ops.spconv_fwd(gout, Wt, None, ctx.n_in, algo=2)
        else:
          srt, perm, mask = ctx.tc_bwd
          gx = ops.spconv_fwd(gout, Wt, srt, ctx.n_in, algo=2, row_perm=perm, tile_mask=mask)
      else:
        Wd = (W3.flip(0) if ctx.mirror else W3).transpose(1, 2).contiguous()       # [K, Cout, Cin]
        gx = ops.spconv_fwd(gout, Wd, ctx.nbr_bwd, ctx.n_in)
    if ctx.needs_input_grad[1]:
      if ctx.tc_fwd is not None and ops.wgrad_tc_supported(x.shape[1], gout.shape[1]):
        if ctx.tc_fwd == "mm":
          gW = ops.spconv_wgrad_tc(x, gout, None, 1)
        else:
          srt, perm, mask = ctx.tc_fwd
          gW = ops.spconv_wgrad_tc(x, gout, srt, W3.shape[0], row_perm=perm, tile_mask=mask)
        gW = gW.view_as(W)
      else:
        gW = ops.spconv_wgrad(x, gout, ctx.nbr_fwd, W3.shape[0]).view_as(W)
    return gx, gW, None, None, None, None, None, None


class _ConvBase(nn.Module):
  TRANSPOSED = False

  def __init__(self, in_channels, out_channels, kernel_size=-1, stride=1, dilation=1, bias=False,
               kernel_generator=None, expand_coordinates=False, convolution_mode=None, dimension=None):
    super().__init__()
    if dimension != 3:
      raise GclbError("gcl_b200 covers dimension=3 only")
    if kernel_generator is not None or expand_coordinates:
      raise GclbError("kernel_generator / expand_coordinates are outside the ResUNet hot path")
    if not isinstance(kernel_size, int) or kernel_size < 1 or kernel_size > 7:
      raise GclbError("kernel_size must be an int in 1..7 (hyper-cube region)")
    self.in_channels, self.out_channels = in_channels, out_channels
    self.kernel_size, self.stride, self.dilation, self.dimension = kernel_size, stride, dilation, dimension
    self.kernel_volume = kernel_size ** 3
    self.use_mm = self.kernel_volume == 1 and stride == 1
    shape = (in_channels, out_channels) if self.use_mm else (self.kernel_volume, in_channels, out_channels)
    self.kernel = nn.Parameter(torch.empty(shape))
    self.bias = nn.Parameter(torch.empty(1, out_channels)) if bias else None
    self._tc_cache = None
    self.reset_parameters()

  def reset_parameters(self):
    with torch.no_grad():
      # ME 0.5: fan = (out_channels if transposed else in_channels) * kernel_volume
      fan = (self.out_channels if self.TRANSPOSED else self.in_channels) * self.kernel_volume
      stdv = 1.0 / np.sqrt(fan)
      self.kernel.uniform_(-stdv, stdv)
      if self.bias is not None:
        self.bias.uniform_(-stdv, stdv)

  def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
    # accept ME-0.4 era [1, Cin, Cout] kernels for 1x1x1 convolutions (A10)
    k = prefix + "kernel"
    if k in state_dict and self.use_mm and state_dict[k].dim() == 3 and state_dict[k].shape[0] == 1:
      state_dict[k] = state_dict[k][0]
    super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

  def forward(self, x: SparseTensor) -> SparseTensor:
    mgr, in_key = x._manager, x.coordinate_map_key
    if self.use_mm:
      out_key, nbr_f, nbr_b, mirror = in_key, None, None, False
    elif not self.TRANSPOSED:
      out_key = mgr.stride(in_key, self.stride) if self.stride > 1 else in_key
      nbr_f = mgr.kernel_map(in_key, out_key, self.kernel_size, self.dilation, False)
      if self.stride == 1 and self.kernel_size % 2 == 1:
        nbr_b, mirror = nbr_f, True
      else:
        nbr_b, mirror = None, False
    else:
      out_key = CoordinateMapKey([s // self.stride for s in in_key.tensor_stride], "")
      if out_key not in mgr.maps:
        raise ValueError("transposed convolution needs an existing coordinate map at the output stride (A7)")
      nbr_f = mgr.kernel_map(in_key, out_key, self.kernel_size, self.dilation, True)
      nbr_b, mirror = None, False
    n_out = mgr.size(out_key)
    needs_grad = torch.is_grad_enabled() and (x.F.requires_grad or self.kernel.requires_grad)
    if needs_grad:
      dual = (out_key, in_key, self.kernel_size, self.dilation, not self.TRANSPOSED)
      if nbr_b is None and not self.use_mm and x.F.requires_grad:
        # dgrad table: the same pairs, indexed by input row (strided <-> transposed swap)
        nbr_b = mgr.kernel_map(*dual)
      tc_f = tc_b = None
      if _CONV_ALGO["training"] == "tf32":
        fwd_key = (in_key, out_key, self.kernel_size, self.dilation, self.TRANSPOSED)
        if ops.tc_supported(self.in_channels, 0, self.out_channels, self.kernel_volume):
          tc_f = "mm" if self.use_mm else mgr.kernel_map_sorted(*fwd_key)
        if x.F.requires_grad and ops.tc_supported(self.out_channels, 0, self.in_channels, self.kernel_volume):
          tc_b = "mm" if self.use_mm else mgr.kernel_map_sorted(*(fwd_key if mirror else dual))
      out = _SparseConvFn.apply(x.F, self.kernel, nbr_f, nbr_b, n_out, mirror, tc_f, tc_b)
      if self.bias is not None:
        out = out + self.bias
    else:
      shift = self.bias.detach().reshape(-1) if self.bias is not None else None
      xf = x.F.detach().contiguous()
      if _CONV_ALGO["inference"] != "fp32" and ops.tc_supported(self.in_channels, 0, self.out_channels, self.kernel_volume):
        # inference: tcgen05 kind::tf32 kernel (features within 1e-3 of fp32); weight image cached per parameter version
        ver = (self.kernel._version, self.kernel.data_ptr())
        if self._tc_cache is None or self._tc_cache[0] != ver:
          self._tc_cache = (ver, ops.weights_to_tc(self.kernel.detach()))
        if self.use_mm:
          out = ops.spconv_fwd(xf, self._tc_cache[1], None, n_out, shift=shift, algo=2)
        else:
          srt, perm, mask = mgr.kernel_map_sorted(in_key, out_key, self.kernel_size, self.dilation, self.TRANSPOSED)
          out = ops.spconv_fwd(xf, self._tc_cache[1], srt, n_out, shift=shift, algo=2, row_perm=perm, tile_mask=mask)
      else:
        out = ops.spconv_fwd(xf, self.kernel.detach(), nbr_f, n_out, shift=shift)
    return SparseTensor(out, coordinate_map_key=out_key, coordinate_manager=mgr)

  def extra_repr(self):
    return (f"in={self.in_channels}, out={self.out_channels}, kernel_size={self.kernel_size}, "
            f"stride={self.stride}, dilation={self.dilation}")


class MinkowskiConvolution(_ConvBase):
  TRANSPOSED = False


class MinkowskiConvolutionTranspose(_ConvBase):
  TRANSPOSED = True


class MinkowskiBatchNorm(nn.Module):
  """BatchNorm1d over the rows of .F; child module `bn` keeps the checkpoint keys (A8, A10).
  Eval mode without autograd runs libgclb200's fused affine kernel; training mode runs libgclb200's own statistics /
  normalise / backward kernels (csrc/bn.cu) with BatchNorm1d's semantics: biased batch variance for the normalisation,
  unbiased for running_var, momentum update, num_batches_tracked."""

  def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True):
    super().__init__()
    self.bn = nn.BatchNorm1d(num_features, eps=eps, momentum=momentum, affine=affine,
                             track_running_stats=track_running_stats)

  def folded(self):
    """(scale, shift) of the eval-mode affine map."""
    bn = self.bn
    scale = torch.rsqrt(bn.running_var + bn.eps)
    if bn.weight is not None:
      scale = scale * bn.weight.detach()
    shift = -bn.running_mean * scale
    if bn.bias is not None:
      shift = shift + bn.bias.detach()
    return scale.contiguous(), shift.contiguous()

  def forward(self, x: SparseTensor, relu: bool = False) -> SparseTensor:
    """relu=True fuses the MEF.relu that follows the norm in every residual block (gcl_b200 extension; the reference's
    model files call MEF.relu separately and get the same numbers through two launches)."""
    bn = self.bn
    f = x.F
    batch_stats = bn.training or not bn.track_running_stats
    if not batch_stats:
      if not (torch.is_grad_enabled() and f.requires_grad):
        scale, shift = self.folded()
        return x._like(ops.affine_act(f.detach(), scale, shift, relu=relu))
      out = bn(f)                       # eval-mode BN inside an autograd graph (fine-tuning with frozen statistics): ATen
      return x._like(torch.relu(out) if relu else out)
    if f.shape[1] % 4 != 0 or f.dtype != torch.float32 or f.shape[0] < 1:
      out = bn(f)
      return x._like(torch.relu(out) if relu else out)
    # training mode: libgclb200's fused statistics / normalise(+ReLU) kernels with their own backward
    factor = 0.0
    if bn.track_running_stats:
      bn.num_batches_tracked += 1
      factor = 1.0 / float(bn.num_batches_tracked) if bn.momentum is None else float(bn.momentum)
    out = _BatchNormTrainFn.apply(f, bn.weight, bn.bias, bn.running_mean if bn.track_running_stats else None,
                                  bn.running_var if bn.track_running_stats else None, float(bn.eps), factor, bool(relu))
    return x._like(out)


class _BatchNormTrainFn(torch.autograd.Function):
  """BatchNorm1d(training) [+ ReLU] over the rows of a feature matrix on gclb_bn_train_fwd / gclb_bn_train_bwd."""

  @staticmethod
  def forward(ctx, x, gamma, beta, running_mean, running_var, eps, momentum, relu):
    xc = x.detach().contiguous()
    n, c = xc.shape
    dev = xc.device
    y = torch.empty_like(xc)
    stats = torch.empty((2, c), dtype=torch.float32, device=dev)         # save_mean | save_invstd
    sums = torch.empty((2, c), dtype=torch.float64, device=dev)
    g = gamma.detach().contiguous() if gamma is not None else None
    b = beta.detach().contiguous() if beta is not None else None
    ops.call("gclb_bn_train_fwd", ops.ptr(xc), n, c, ops.ptr(g), ops.ptr(b), eps, momentum, ops.ptr(running_mean),
             ops.ptr(running_var), int(relu), ops.ptr(y), stats[0].data_ptr(), stats[1].data_ptr(), ops.ptr(sums), ops.stream())
    ctx.save_for_backward(xc, y if relu else None, g, stats)
    ctx.relu = relu
    ctx.has_affine = gamma is not None
    return y

  @staticmethod
  def backward(ctx, dy):
    xc, y, g, stats = ctx.saved_tensors
    n, c = xc.shape
    dy = dy.contiguous()
    dx = torch.empty_like(xc)
    dgb = torch.empty((2, c), dtype=torch.float32, device=xc.device)     # dgamma | dbeta
    sums = torch.empty((2, c), dtype=torch.float64, device=xc.device)
    ops.call("gclb_bn_train_bwd", ops.ptr(xc), ops.ptr(dy), ops.ptr(y), n, c, ops.ptr(g), stats[0].data_ptr(), stats[1].data_ptr(),
             int(ctx.relu), ops.ptr(dx), dgb[0].data_ptr(), dgb[1].data_ptr(), ops.ptr(sums), ops.stream())
    return (dx, dgb[0] if ctx.has_affine else None, dgb[1] if ctx.has_affine else None, None, None, None, None, None)


def bn_relu(norm, x: SparseTensor) -> SparseTensor:
  """MEF.relu(norm(x)) in one launch pair (gcl_b200 extension used by gcl_b200.resunet; the oracle module has no such helper)"""
  return norm(x, relu=True)


class MinkowskiInstanceNorm(nn.Module):
  def __init__(self, num_features, dimension=-1):
    super().__init__()
    self.num_features = num_features
    self.weight = nn.Parameter(torch.ones(1, num_features))
    self.bias = nn.Parameter(torch.zeros(1, num_features))

  def forward(self, x):
    raise NotImplementedError("MinkowskiInstanceNorm is outside the ResUNetBN2C hot path (SURVEY.md section 2 row 3)")


class MinkowskiReLU(nn.Module):
  def __init__(self, inplace=False):
    super().__init__()

  def forward(self, x):
    return MinkowskiFunctional.relu(x)


from . import MinkowskiFunctional  # noqa: E402
from . import utils  # noqa: E402
from .utils import sparse_quantize, sparse_collate, batched_coordinates  # noqa: E402,F401
