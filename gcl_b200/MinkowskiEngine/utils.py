"""MinkowskiEngine.utils subset on the GPU.

sparse_quantize accepts CPU or CUDA input like ME's; the de-duplication always runs in libgclb200 on `device`
(default: the input's device if CUDA, else the current CUDA device) and results are returned on the input's
device / container type, so the reference's loaders (`_, sel = ME.utils.sparse_quantize(xyz / voxel,
return_index=True)`, lib/complement_data_loader.py:788-789) keep working.  There is no CPU implementation.
"""
import numpy as np
import torch

from .. import ops
from .._lib import GclbError


def _cuda_device(t, device):
  if isinstance(t, torch.Tensor) and t.is_cuda:
    return t.device
  if device is not None and str(device) != "cpu":
    return torch.device(device)
  if not torch.cuda.is_available():
    raise GclbError("sparse_quantize needs a CUDA device: gcl_b200 has no CPU fallback")
  return torch.device("cuda", torch.cuda.current_device())


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
  if labels is not None:
    raise GclbError("labels are not used on the GCL hot path")
  is_torch = isinstance(coordinates, torch.Tensor)
  src_dev = coordinates.device if is_torch else torch.device("cpu")
  dev = _cuda_device(coordinates, device)
  c = coordinates if is_torch else torch.from_numpy(np.ascontiguousarray(coordinates))
  if quantization_size is not None:
    c = c / quantization_size
  c = c.to(dev)
  rows = torch.floor(c).to(torch.int32) if c.dtype.is_floating_point else c.to(torch.int32)
  if return_inverse:
    cm, umap, inv = ops.quantize_rows(rows, return_inverse=True)
    inv = inv.to(torch.int64)
  else:
    cm, umap = ops.quantize_rows(rows)
    inv = None
  uniq = rows[umap]
  back = (lambda t: t.to(src_dev)) if is_torch else (lambda t: t.cpu().numpy())
  if return_maps_only:
    return (back(umap), back(inv)) if return_inverse else back(umap)
  out = [back(uniq)]
  if features is not None:
    idx = umap.to(features.device) if isinstance(features, torch.Tensor) else umap.cpu().numpy()
    out.append(features[idx])
  if return_index:
    out.append(back(umap))
  if return_inverse:
    out.append(back(inv))
  return out[0] if len(out) == 1 else tuple(out)


def batched_coordinates(coords, dtype=torch.int32, device=None):
  rows = []
  for b, c in enumerate(coords):
    c = torch.as_tensor(c)
    c = torch.floor(c).to(dtype) if c.dtype.is_floating_point else c.to(dtype)
    rows.append(torch.cat([torch.full((len(c), 1), b, dtype=dtype, device=c.device), c], dim=1))
  out = torch.cat(rows, 0) if rows else torch.zeros((0, 4), dtype=dtype)
  return out.to(device) if device is not None else out


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
  bc = batched_coordinates(coords, dtype=dtype, device=device)
  f = torch.cat([torch.as_tensor(x) for x in feats], 0)
  if device is not None:
    f = f.to(device)
  if labels is not None:
    return bc, f, torch.cat([torch.as_tensor(x) for x in labels], 0)
  return bc, f
