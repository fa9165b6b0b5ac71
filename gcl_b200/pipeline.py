"""The inference hot path as one call: scan pairs in -> mutual feature correspondences out
(scripts/test_kitti.py:141-154 per pair: 2x ResUNet forward, 5000-point subsample, feature NN; with the mutual
filter of generalization_ETH/evaluate.py:63-77).  Batches several pairs per launch so that the ~60 kernel
launches of a forward are amortised."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .engine import ResUNetEngine


class PairMatcher:
  def __init__(self, model, voxel: float = 0.3, subsample: int = 5000, device="cuda", seed: int = 0, algo: int = 0):
    self.engine = model if isinstance(model, ResUNetEngine) else ResUNetEngine(model, device=device, algo=algo)
    self.voxel, self.subsample = float(voxel), int(subsample)
    self.device = torch.device(device)
    self.gen = torch.Generator(device=self.device)
    self.gen.manual_seed(seed)

  @torch.no_grad()
  def match(self, xyz: torch.Tensor, cloud_ptr: torch.Tensor):
    """xyz float32 [P,3] (device, or pinned host: copied asynchronously), cloud_ptr int64 [2*n_pairs+1] (host);
    clouds are ordered (pair0.scan0, pair0.scan1, pair1.scan0, ...).
    Returns dict: pairs int64 [*,2] rows (i, j) = indices into the SUBSAMPLED rows of scan0 / scan1,
    pair_ptr int64 [n_pairs+1] (device), sel0 / sel1 (voxel rows of the subsample), unique_map (voxel row ->
    input point), n_voxels (host list per cloud)."""
    if not xyz.is_cuda:
      xyz = xyz.to(self.device, non_blocking=True)
    n_clouds = cloud_ptr.numel() - 1
    assert n_clouds % 2 == 0, "clouds come in pairs"
    feats, cm, umap = self.engine.extract(xyz, self.voxel, cloud_ptr)
    # voxel rows are grouped by cloud in order: row range of each cloud from the batch column
    counts = torch.bincount(cm.coords[:, 0].long(), minlength=n_clouds).tolist()
    starts = [0]
    for c in counts:
      starts.append(starts[-1] + c)
    S = self.subsample
    sel = []
    for c in range(n_clouds):
      v = counts[c]
      if S > 0 and v > S:
        sel.append(torch.randperm(v, device=self.device, generator=self.gen)[:S] + starts[c])
      else:
        sel.append(torch.arange(starts[c], starts[c + 1], device=self.device))
    sel0, sel1 = torch.cat(sel[0::2]), torch.cat(sel[1::2])
    a_ptr, b_ptr = [0], [0]
    for p in range(n_clouds // 2):
      a_ptr.append(a_ptr[-1] + sel[2 * p].numel())
      b_ptr.append(b_ptr[-1] + sel[2 * p + 1].numel())
    F0, F1 = feats.index_select(0, sel0), feats.index_select(0, sel1)
    idx01, d01, idx10, d10, a_dev, b_dev, ws = ops.nn_search(F0, F1, a_ptr, b_ptr, both=True)
    pairs, pair_ptr = ops.mutual_filter(idx01, idx10, a_dev, b_dev, ws)
    return dict(pairs=pairs, pair_ptr=pair_ptr, sel0=sel0, sel1=sel1, a_ptr=a_ptr, b_ptr=b_ptr, idx01=idx01,
                unique_map=umap, n_voxels=counts, feats=feats, coords=cm.coords)
