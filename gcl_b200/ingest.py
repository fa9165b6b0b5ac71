"""Data ingest: velodyne .bin files -> pinned host memory -> ONE H2D copy -> point matrix on the device.

Reference: lib/complement_data_loader.py:358-361 (`_get_xyz`: np.fromfile(fname, np.float32).reshape(-1, 4)[:, :3]) and the
augmentation of the pair loaders :65-70, :753-781 (random rotation `pts @ R.T + T`, random scale), which the reference runs in
DataLoader worker processes on the CPU before `sparse_quantize`.  Here the file bytes land directly in a pinned staging buffer
(`readinto`, no intermediate numpy array), cross PCIe once per batch, and the slice / transform / scale is one HBM-bound kernel
(gclb_ingest_points) feeding K1 (`gclb_voxelize`).
"""
from __future__ import annotations

import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream


def write_velodyne_bin(path: str, xyz: np.ndarray, reflectance: Optional[np.ndarray] = None):
  """KITTI layout: float32 records (x, y, z, reflectance) -- for synthetic benchmarks and tests"""
  rec = np.zeros((len(xyz), 4), np.float32)
  rec[:, :3] = xyz
  if reflectance is not None:
    rec[:, 3] = reflectance
  rec.tofile(path)


class ScanReader:
  """reads batches of velodyne files into a reusable pinned buffer"""

  def __init__(self, capacity_points: int = 4_500_000, width: int = 4, threads: int = 8):
    self.width = width
    self.buf = torch.empty((capacity_points, width), dtype=torch.float32).pin_memory()
    self._np = self.buf.numpy()
    # file reads release the GIL: a few threads keep several reads in flight (one thread tops out at ~4-5 GB/s from the page cache)
    self._pool = None
    if threads > 1:
      import concurrent.futures as cf
      self._pool = cf.ThreadPoolExecutor(max_workers=threads)

  def read(self, paths: Sequence[str]):
    """-> (records: pinned float32 [P, width] view, cloud_ptr int64 [n+1] host tensor)"""
    off, ptrs, jobs = 0, [0], []
    rec_bytes = 4 * self.width
    for p in paths:
      size = os.path.getsize(p)
      if size % rec_bytes:
        raise _lib.GclbError(f"{p}: size {size} is not a multiple of {rec_bytes} bytes (float32 x {self.width} records)")
      n = size // rec_bytes
      if off + n > self.buf.shape[0]:
        raise _lib.GclbError("ScanReader capacity exceeded: construct it with a larger capacity_points")
      jobs.append((p, off, n, size))
      off += n
      ptrs.append(off)

    def load(job):
      p, o, n, size = job
      with open(p, "rb", buffering=0) as f:
        got = f.readinto(memoryview(self._np[o:o + n]).cast("B"))
      if got != size:
        raise _lib.GclbError(f"{p}: short read ({got} of {size} bytes)")

    if self._pool is not None and len(jobs) > 1:
      list(self._pool.map(load, jobs))
    else:
      for j in jobs:
        load(j)
    return self.buf[:off], torch.tensor(ptrs, dtype=torch.int64)


def points_to_device(records: torch.Tensor, cloud_ptr: torch.Tensor, device, transforms=None, scales=None) -> torch.Tensor:
  """records float32 [P, 4|3] (pinned host or device) -> xyz float32 [P, 3] on `device`; optional per-cloud 4x4 transforms
  (applied as float32 `pts @ R.T + T`, complement_data_loader.py:65-70) and scales (:771-775)"""
  dev = torch.device(device)
  rec = records if records.is_cuda else records.to(dev, non_blocking=True)
  rec = rec.contiguous()
  P, width = rec.shape
  n_clouds = cloud_ptr.numel() - 1
  cp = cloud_ptr.to(torch.int64)
  cp = cp if cp.is_cuda else cp.pin_memory().to(dev, non_blocking=True)
  T = S = None
  if transforms is not None:
    T = torch.as_tensor(np.asarray(transforms, dtype=np.float32).reshape(n_clouds, 4, 4)).pin_memory().to(dev, non_blocking=True).contiguous()
  if scales is not None:
    S = torch.as_tensor(np.asarray(scales, dtype=np.float32).reshape(n_clouds)).pin_memory().to(dev, non_blocking=True).contiguous()
  xyz = torch.empty((P, 3), dtype=torch.float32, device=dev)
  call("gclb_ingest_points", ptr(rec), P, width, ptr(cp), n_clouds, ptr(T), ptr(S), ptr(xyz), stream())
  return xyz
