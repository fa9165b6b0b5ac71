"""Positive-group construction on the GPU: the reference's util/pointcloud.py:69-132
`get_matching_indices_colocation`, which the colocation loaders call once per sample (lib/colocation_data_loader.py:394,
672) and which loops over every centre point through 1 + J Open3D KD-trees in Python.  Here the clouds' voxel hashes (K1)
answer the radius queries; see csrc/groups.cu."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib, ops
from ._lib import call, ptr, stream


def colocation_groups(center_xyz: torch.Tensor, neighbourhood_xyz: Sequence[torch.Tensor], list_trans, voxel_size: float,
                      search_voxel_size: float, K: Optional[int] = None, kcap: int = 32):
  """center_xyz float32 [Nc,3] and neighbourhood_xyz[j] float32 [Nj,3]: the loader's voxel-downsampled clouds, each in its
  own sensor frame (one point per voxel of size voxel_size: lib/colocation_data_loader.py:379-390); list_trans[j] 4x4 maps
  cloud j into the centre frame.  Returns device tensors (group int64 [G], index int64 [sum group], finest_flag bool)."""
  dev = center_xyz.device
  assert dev.type == "cuda", "gcl_b200 has no CPU path"
  c = center_xyz.to(torch.float32).contiguous()
  nbs = [x.to(device=dev, dtype=torch.float32).contiguous() for x in neighbourhood_xyz]
  J = len(nbs)
  assert J >= 1 and len(list_trans) == J
  nb = torch.cat(nbs, 0)
  nb_ptr = torch.tensor(np.cumsum([0] + [len(x) for x in nbs]), dtype=torch.int64)
  cm_c, _ = ops.voxelize(c, voxel_size)
  cm_n, _ = ops.voxelize(nb, voxel_size, nb_ptr)
  if cm_c.n != c.shape[0] or cm_n.n != nb.shape[0]:
    raise _lib.GclbError("colocation_groups expects clouds with one point per voxel (ME.utils.sparse_quantize'd at voxel_size)")
  T = np.stack([np.asarray(t, dtype=np.float64).reshape(4, 4) for t in list_trans])
  Ti = np.linalg.inv(T)
  T_d = torch.from_numpy(T).to(dev).contiguous()
  Ti_d = torch.from_numpy(Ti).to(dev).contiguous()
  k = int(K) if K is not None else 0
  kc = k if k > 0 else int(kcap)
  n = c.shape[0]
  lib = _lib.load()
  group = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
  index = torch.empty(max(n * (1 + J) * kc, 1), dtype=torch.int64, device=dev)
  finest = torch.empty(max(n * (1 + J) * kc, 1), dtype=torch.uint8, device=dev)
  counts = torch.zeros(2, dtype=torch.int64, device=dev)
  status = torch.zeros(1, dtype=torch.int32, device=dev)
  ws = torch.empty(max(int(lib.gclb_groups_workspace_bytes(n, J, kc)), 8), dtype=torch.uint8, device=dev)
  call("gclb_colocation_groups", ptr(c), n, ptr(cm_c.table), cm_c.capacity, ptr(nb), ptr(nb_ptr.to(dev)), ptr(cm_n.table),
       cm_n.capacity, ptr(T_d), ptr(Ti_d), J, float(voxel_size), float(search_voxel_size), k, kc, ptr(group), ptr(index),
       ptr(finest), ptr(counts[0:1]), ptr(counts[1:2]), ptr(status), ptr(ws), stream())
  _lib.check_status(status, "colocation_groups")
  g, m = (int(v) for v in counts.tolist())
  return group[:g], index[:m], finest[:m].bool()


def get_matching_indices_colocation(center_xyz, neighbourhood_xyz, list_trans, search_voxel_size, voxel_size, K=None):
  """reference return convention (util/pointcloud.py:69-132 with calc_distance_err=False): python lists
  (group, index, finest_flag, central_distance=[])"""
  g, i, f = colocation_groups(center_xyz, neighbourhood_xyz, list_trans, voxel_size, search_voxel_size, K)
  return g.tolist(), i.tolist(), f.float().tolist(), []


def exhaustive_hash(group: torch.Tensor, index: torch.Tensor, M: int) -> torch.Tensor:
  """util/misc.py:29-36 `_exhaustive_hash(torch.split(index, group), M)` on the device: int64 keys of every unordered pair
  inside a group, reference order.  group int64 [G] sizes, index int64 [sum group] (device)."""
  dev = index.device
  assert dev.type == "cuda", "gcl_b200 has no CPU path"
  group = group.to(device=dev, dtype=torch.int64)
  index = index.to(torch.int64).contiguous()
  G = group.numel()
  gptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
  if G:
    gptr[1:] = torch.cumsum(group, 0)
  n_keys = int((group * (group - 1) // 2).sum().item()) if G else 0
  keys = torch.empty(max(n_keys, 1), dtype=torch.int64, device=dev)
  n_out = torch.zeros(1, dtype=torch.int64, device=dev)
  lib = _lib.load()
  ws = torch.empty(max(int(lib.gclb_exhaustive_hash_workspace_bytes(G)), 8), dtype=torch.uint8, device=dev)
  call("gclb_exhaustive_hash", ptr(gptr), ptr(index), G, int(M), ptr(keys), ptr(n_out), ptr(ws), stream())
  return keys[:n_keys]
