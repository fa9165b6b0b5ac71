"""ORACLE (test infrastructure): MinkowskiEngine.utils subset.

Call sites restated: lib/complement_data_loader.py:788-789 (`_, sel = ME.utils.sparse_quantize(xyz / voxel,
return_index=True)`), lib/colocation_data_loader.py:379,388, util/misc.py:118,120,
lib/complement_data_loader.py:1310-1311 (sparse_collate).
"""
import numpy as np
import torch


def _to_numpy(x):
  return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def quantize_rows(discrete: np.ndarray):
  """first-occurrence dedupe over full rows (any width) -> (unique_map ascending, inverse_map)."""
  d = np.ascontiguousarray(discrete)
  if len(d) == 0:
    return np.zeros(0, np.int64), np.zeros(0, np.int64)
  _, first, inv = np.unique(d, axis=0, return_index=True, return_inverse=True)
  inv = np.asarray(inv).reshape(-1)
  order = np.argsort(first, kind="stable")
  rank = np.empty_like(order)
  rank[order] = np.arange(len(order))
  return first[order].astype(np.int64), rank[inv].astype(np.int64)


def sparse_quantize(coordinates, features=None, labels=None, ignore_label=-100, return_index=False,
                    return_inverse=False, return_maps_only=False, quantization_size=None, device="cpu"):
  """A1.  floor -> int32 -> hash-dedupe in input order, first occurrence wins; unique_map ascending.
  If `quantization_size` is given the coordinates are divided by it first (in their own dtype)."""
  assert labels is None, "labels are not used on the reference path"
  is_torch = isinstance(coordinates, torch.Tensor)
  c = coordinates
  if quantization_size is not None:
    c = c / quantization_size
  if is_torch:
    discrete = torch.floor(c).int().cpu().numpy() if c.dtype.is_floating_point else c.int().cpu().numpy()
  else:
    c = np.asarray(c)
    discrete = np.floor(c).astype(np.int32) if np.issubdtype(c.dtype, np.floating) else c.astype(np.int32)
  unique_map, inverse_map = quantize_rows(discrete)
  wrap = (lambda a: torch.from_numpy(a)) if is_torch else (lambda a: a)
  if return_maps_only:
    return (wrap(unique_map), wrap(inverse_map)) if return_inverse else wrap(unique_map)
  out = [wrap(discrete[unique_map])]
  if features is not None:
    out.append(features[unique_map])
  if return_index:
    out.append(wrap(unique_map))
  if return_inverse:
    out.append(wrap(inverse_map))
  return out[0] if len(out) == 1 else tuple(out)


def batched_coordinates(coords, dtype=torch.int32, device=None):
  """A2: int32 [sum N, 4]; column 0 = index of the cloud in the list."""
  rows = []
  for b, c in enumerate(coords):
    c = torch.as_tensor(_to_numpy(c))
    c = torch.floor(c).to(dtype) if c.dtype.is_floating_point else c.to(dtype)
    rows.append(torch.cat([torch.full((len(c), 1), b, dtype=dtype), c], dim=1))
  out = torch.cat(rows, 0) if rows else torch.zeros((0, 4), dtype=dtype)
  return out.to(device) if device is not None else out


def sparse_collate(coords, feats, labels=None, dtype=torch.int32, device=None):
  bc = batched_coordinates(coords, dtype=dtype, device=device)
  f = torch.cat([torch.as_tensor(_to_numpy(x)) if not isinstance(x, torch.Tensor) else x for x in feats], 0)
  if device is not None:
    f = f.to(device)
  if labels is not None:
    l = torch.cat([torch.as_tensor(_to_numpy(x)) for x in labels], 0)
    return bc, f, l
  return bc, f
