"""SC2-PCR registration on libgclb200: same class, constructor arguments and method names as
/root/reference/scripts/SC2_PCR/SC2_PCR.py `Matcher`, so `scripts/test_kitti.py:180-182` reads unchanged:

    matcher = Matcher(inlier_threshold=0.6, num_node=8000, use_mutual=False, d_thre=0.1, num_iterations=20, ratio=0.2,
                      nms_radius=0.6, max_points=8000, k1=30, k2=20)
    T, labels, src_corr, tgt_corr = matcher.estimator(xyz0[None], xyz1[None], F0[None], F1[None])

plus a batched entry (`sc2_pcr_batch`) that registers many scan pairs in one set of kernel launches.  CUDA tensors only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, matching
from ._lib import call, ptr, stream, GclbError


def sc2_pcr_batch(src: torch.Tensor, tgt: torch.Tensor, seg_ptr: torch.Tensor, n_max: int, *, d_thre=0.1,
                  inlier_threshold=0.6, nms_radius=0.6, ratio=0.2, num_iterations=20, k1=30, k2=20, max_points=8000,
                  refine_iters=20):
  """src, tgt float32 [sum n, 3] (row i <-> row i), seg_ptr int64 [P+1] on the device, n_max = host bound of the segment
  lengths.  Returns (trans float32 [P,4,4], info int32 [P,4] = rows used, seeds, best-seed inliers, refined inliers)."""
  _lib.require_cuda(src, tgt, seg_ptr)
  src, tgt = src.float().contiguous(), tgt.float().contiguous()
  seg_ptr = seg_ptr.to(torch.int64).contiguous()
  P = seg_ptr.numel() - 1
  dev = src.device
  lib = _lib.load()
  n_max = int(min(max(n_max, 1), max_points))
  ws = torch.empty(int(lib.gclb_sc2pcr_workspace_bytes(n_max, P, float(ratio))), dtype=torch.uint8, device=dev)
  trans = torch.empty((P, 4, 4), dtype=torch.float32, device=dev)
  info = torch.empty((P, 4), dtype=torch.int32, device=dev)
  call("gclb_sc2pcr", ptr(src), ptr(tgt), ptr(seg_ptr), P, n_max, float(d_thre), float(inlier_threshold), float(nms_radius),
       float(ratio), int(num_iterations), int(k1), int(k2), int(max_points), int(refine_iters), ptr(trans), ptr(info), ptr(ws),
       stream())
  return trans, info


class Matcher:
  """scripts/SC2_PCR/SC2_PCR.py:7-31 (same defaults)."""

  def __init__(self, inlier_threshold=0.10, num_node="all", use_mutual=True, d_thre=0.1, num_iterations=10, ratio=0.2,
               nms_radius=0.1, max_points=8000, k1=30, k2=20, select_scene=None):
    self.inlier_threshold, self.num_node, self.use_mutual = inlier_threshold, num_node, use_mutual
    self.d_thre, self.num_iterations, self.ratio = d_thre, num_iterations, ratio
    self.max_points, self.nms_radius, self.k1, self.k2 = max_points, nms_radius, k1, k2

  def match_pair(self, src_keypts, tgt_keypts, src_features, tgt_features):
    return matching.match_pair(src_keypts, tgt_keypts, src_features, tgt_features, num_node=self.num_node)

  def SC2_PCR(self, src_keypts, tgt_keypts):
    """[1, n, 3] x 2 -> [1, 4, 4]  (:304-381; batch size 1 like the reference's own asserts)"""
    if src_keypts.dim() != 3 or src_keypts.shape[0] != 1:
      raise GclbError("SC2_PCR expects [1, num_corr, 3] tensors (the reference asserts bs == 1)")
    n = src_keypts.shape[1]
    seg = torch.tensor([0, n], dtype=torch.int64, device=src_keypts.device)
    trans, _ = sc2_pcr_batch(src_keypts[0], tgt_keypts[0], seg, n, d_thre=self.d_thre, inlier_threshold=self.inlier_threshold,
                             nms_radius=self.nms_radius, ratio=self.ratio, num_iterations=self.num_iterations, k1=self.k1,
                             k2=self.k2, max_points=self.max_points)
    return trans

  def estimator(self, src_keypts, tgt_keypts, src_features, tgt_features):
    """:383-411 -> (pred_trans [1,4,4], pred_labels [1,n], src_keypts_corr, tgt_keypts_corr)"""
    src_corr, tgt_corr = self.match_pair(src_keypts, tgt_keypts, src_features, tgt_features)
    pred_trans = self.SC2_PCR(src_corr, tgt_corr)
    warped = (pred_trans[:, :3, :3] @ src_corr.permute(0, 2, 1) + pred_trans[:, :3, 3:4]).permute(0, 2, 1)
    distance = torch.sum((warped - tgt_corr) ** 2, dim=-1) ** 0.5
    pred_labels = (distance < self.inlier_threshold).float()
    return pred_trans, pred_labels, src_corr, tgt_corr
