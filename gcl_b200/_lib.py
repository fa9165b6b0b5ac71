"""ctypes binding of libgclb200.so (include/gclb200.h).  No fallback: if the library is missing or a call
fails, this raises -- the product path never degrades to PyTorch/CPU code."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgclb200.so")

_p, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes): must list every symbol of include/gclb200.h and include/gclb200_debug.h (tests/test_abi.py checks it)
SIGNATURES = {
    "gclb_last_error": (C.c_char_p, []),
    "gclb_version": (C.c_int, []),
    "gclb_kernel_launches": (_i64, []),
    "gclb_has_tcgen05": (C.c_int, []),
    "gclb_hash_capacity": (_i64, [_i64]),
    "gclb_hash_bytes": (_sz, [_i64]),
    "gclb_hash_build": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p]),
    "gclb_hash_query": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p]),
    "gclb_compact_workspace_bytes": (_sz, [_i64]),
    "gclb_voxelize": (C.c_int, [_p, _i64, _p, _i32, _f32, _p, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "gclb_quantize_rows": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _p, _p, _p, _p, _p, _p]),
    "gclb_stride_map": (C.c_int, [_p, _i64, _p, _i32, _p, _i64, _p, _p, _p, _p, _p, _p]),
    "gclb_kmap_build": (C.c_int, [_p, _i64, _p, _i64, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "gclb_kmap_pairs": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p]),
    "gclb_kmap_sort_workspace_bytes": (_sz, [_i64]),
    "gclb_kmap_sort_rows": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gclb_spconv_fwd": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _i32, _p, _i64,
                                  _i32, _p]),
    "gclb_kmap_halo_bytes": (_sz, [_i64]),
    "gclb_debug_halo_prof": (C.c_int, [_p]),
    "gclb_debug_umma_rate": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _p, _p]),
    "gclb_kmap_halo_max_groups": (_i32, []),
    "gclb_kmap_halo_build": (C.c_int, [_p, _i64, _p, _p, _sz, _p, _p, _p, _p, _p]),
    "gclb_spconv_fwd_halo": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _p, _p, _p, _p, _p, _p, _p, _i32, _p, _i64, _p]),
    "gclb_spconv_set_range_monitor": (C.c_int, [_p]),
    "gclb_range_check": (C.c_int, [_p, _i32, _p, _p]),
    "gclb_spconv_fwd_probe": (C.c_int, [_p, _i32, _p, _i32, _i32, _p, _i64, _p, _i64, _i32, _i32, _p, _p, _p, _i32, _p,
                                        _p, _p, _p, _p, _p]),
    "gclb_groups_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "gclb_exhaustive_hash_workspace_bytes": (_sz, [_i64]),
    "gclb_exhaustive_hash": (C.c_int, [_p, _p, _i64, _i64, _p, _p, _p, _p]),
    "gclb_colocation_groups": (C.c_int, [_p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _p, _i32, C.c_float, C.c_double, _i32, _i32,
                                         _p, _p, _p, _p, _p, _p, _p, _p]),
    "gclb_weights_to_tc": (C.c_int, [_p, _i32, _i32, _i32, _p, _p]),
    "gclb_weights_to_tc_dgrad": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "gclb_weights_to_tc_f16": (C.c_int, [_p, _i32, _i32, _i32, _i32, _p, _p]),
    "gclb_spconv_wgrad": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _p, _i32, _p, _p]),
    "gclb_spconv_wgrad_tc": (C.c_int, [_p, _i32, _i64, _p, _i32, _i64, _p, _p, _p, _i32, _p, _p]),
    "gclb_pointwise_tail": (C.c_int, [_p, _i32, _p, _i32, _i64, _p, _i32, _p, _p, _i32, _i32, _p, _p]),
    "gclb_affine_act": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _i32, _p, _p]),
    "gclb_bn_train_fwd": (C.c_int, [_p, _i64, _i32, _p, _p, _f32, _f32, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "gclb_bn_train_bwd": (C.c_int, [_p, _p, _p, _i64, _i32, _p, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "gclb_nn_workspace_bytes": (_sz, [_i64, _i64, _i32, _i64, _i64]),
    "gclb_nn": (C.c_int, [_p, _p, _i32, _p, _p, _i32, _p, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p, _i32, _p, _p]),
    "gclb_subsample": (C.c_int, [_p, _p, _i64, _i32, _i64, _i32, C.c_uint64, _p, _p, _p, _p]),
    "gclb_mutual_filter": (C.c_int, [_p, _p, _p, _p, _i32, _i64, _p, _p, _p, _p]),
    "gclb_ingest_points": (C.c_int, [_p, _i64, _i32, _p, _i32, _p, _p, _p, _p]),
    "gclb_corr_points": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i32, _i64, _p, _p, _p]),
    "gclb_pair_metrics": (C.c_int, [_p, _p, _p, _p, _p, _i32, _f32, _p, _p]),
    "gclb_sc2pcr_workspace_bytes": (_sz, [_i64, _i32, C.c_double]),
    "gclb_sc2pcr": (C.c_int, [_p, _p, _p, _i32, _i64, _f32, _f32, _f32, C.c_double, _i32, _i32, _i32, _i32, _i32, _p, _p, _p, _p]),
    "gclb_loss_workspace_bytes": (_sz, [_i64, _i64]),
    "gclb_debug_tma_gather4": (C.c_int, [_p, _i64, _i32, _i32, _i32, _p, _p, _p]),
    "gclb_group_loss": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _f32, _f32, _f32,
                                  _i32, _p, _p, _p, _p, _p]),
    "gclb_group_loss_bwd": (C.c_int, [_p, _i64, _i32, _p, _p, _p, _p, _i64, _p, _p, _i64, _p, _i64, _f32, _f32, _f32,
                                      _i32, _p, _p, _p, _p, _p]),
}

_lib = None


class GclbError(RuntimeError):
  pass


def load():
  """Load the shared library (building is explicit: `python -m gcl_b200.build` / __graft_entry__.build())."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(f"{LIB_PATH} not found: build it with `python -m gcl_b200.build` (nvcc, sm_100a). "
                      "gcl_b200 has no CPU or PyTorch fallback.")
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SIGNATURES.items():
    fn = getattr(lib, name)
    fn.restype, fn.argtypes = res, args
  _lib = lib
  return lib


def ptr(t):
  """device (or host) pointer of a contiguous tensor, None -> NULL"""
  if t is None:
    return None
  assert t.is_contiguous(), "gcl_b200 kernels need contiguous tensors"
  return t.data_ptr()


def stream():
  # the raw handle of torch's current stream; torch.cuda.current_stream() builds a Python Stream object per call (~15 us,
  # 180 launches per training step), the C accessor is ~0.3 us
  return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def call(name, *args):
  lib = load()
  rc = getattr(lib, name)(*args)
  if rc != 0:
    raise GclbError(f"{name} failed ({rc}): {lib.gclb_last_error().decode()}")


def require_cuda(*tensors):
  for t in tensors:
    if t is not None and not t.is_cuda:
      raise GclbError("gcl_b200 operators run on CUDA tensors only (there is no CPU fallback); "
                      f"got a tensor on {t.device}")


ST_RANGE, ST_FULL, ST_DUPLICATE, ST_FP16_OVERFLOW, ST_FP16_UNDERFLOW = 1, 2, 4, 8, 16


def check_status(status: torch.Tensor, what: str):
  """Read the device status word (one host sync) and raise on any error bit."""
  s = int(status.item())
  if s & ST_RANGE:
    raise GclbError(f"{what}: coordinate outside the supported range (batch < 1023, |xyz| < 2^17 voxels) or non-finite")
  if s & ST_FULL:
    raise GclbError(f"{what}: coordinate hash table full")
  if s & ST_DUPLICATE:
    raise GclbError(f"{what}: duplicate coordinate rows (quantize with ME.utils.sparse_quantize first)")
  if s & ST_FP16_OVERFLOW:
    raise GclbError(f"{what}: an activation left the finite fp16 range (|y| > 65504 or NaN) and was saturated -- these "
                    "weights need fp32 activation storage: ResUNetEngine(model, half=False)")
  if s & ST_FP16_UNDERFLOW:
    raise GclbError(f"{what}: a whole activation tensor is below 2^-11 in magnitude (fp16 subnormal range) -- these "
                    "weights need fp32 activation storage: ResUNetEngine(model, half=False)")
