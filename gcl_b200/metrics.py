"""Evaluation metrics and the alternative (pair-wise) loss head on libgclb200 primitives.

  pair_metrics / registration_recall   scripts/test_kitti.py:188-217 (RTE, RRE, success = RTE < 2 m and RRE < 5 deg) and
                                       lib/trainer.py:406-409 (evaluate_hit_ratio), :357-359 (feat_match_ratio = hit_ratio > 0.05)
  HardestContrastiveLoss               lib/trainer.py:412-462 (HardestContrastiveLossTrainer.contrastive_hardest_negative_loss):
                                       FCGF's pair-wise loss, the baseline head GCL's group loss replaces
CUDA tensors only.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from ._lib import call, ptr, stream


def pair_metrics(trans_est: torch.Tensor, trans_gt: torch.Tensor, src=None, tgt=None, seg_ptr=None, hit_thresh: float = 0.1):
  """trans_* float32 [P,4,4] (device); optional correspondences src/tgt float32 [sum n,3] with seg_ptr int64 [P+1].
  Returns float32 [P,4] on the device: RTE (m), RRE (deg), hit ratio, #correspondences."""
  _lib.require_cuda(trans_est, trans_gt, src, tgt, seg_ptr)
  E, Gt = trans_est.float().contiguous(), trans_gt.to(trans_est.device).float().contiguous()
  P = E.shape[0]
  out = torch.empty((P, 4), dtype=torch.float32, device=E.device)
  s = src.float().contiguous() if src is not None else None
  t = tgt.float().contiguous() if tgt is not None else None
  sp = seg_ptr.to(torch.int64).contiguous() if seg_ptr is not None else None
  call("gclb_pair_metrics", ptr(E), ptr(Gt), ptr(s), ptr(t), ptr(sp), P, float(hit_thresh), ptr(out), stream())
  return out


def registration_recall(metrics: torch.Tensor, rte_thresh: float = 2.0, rre_thresh: float = 5.0, feat_match_thresh: float = 0.05):
  """aggregates like scripts/test_kitti.py:197-217 / lib/trainer.py:357-375: success rate, mean RTE / RRE over the successful
  pairs (RRE NaN counts as failure), mean hit ratio and feature-match recall.  One host read."""
  m = metrics.cpu()
  rte, rre, hit = m[:, 0], m[:, 1], m[:, 2]
  ok_rte = rte < rte_thresh
  ok_rre = (~torch.isnan(rre)) & (rre < rre_thresh)
  ok = ok_rte & ok_rre
  mean = lambda v, k: float(v[k].mean()) if bool(k.any()) else float("nan")
  return {"success_rate": float(ok.float().mean()), "rte": mean(rte, ok_rte), "rre": mean(rre, ok_rre),
          "hit_ratio": float(hit.mean()), "feat_match_ratio": float((hit > feat_match_thresh).float().mean()), "n": int(len(m))}


class HardestContrastiveLoss:
  """loss = HardestContrastiveLoss(pos_thresh=0.1, neg_thresh=1.4)
  pos, neg = loss(F0, F1, positive_pairs, num_pos=5192, num_hn_samples=2048)      # lib/trainer.py:412-462
  The hardest negatives (two row-wise arg-mins over 5192 x 2048 distance matrices) come from the fused K4 kernel (no N x M matrix,
  row-indirect operands); the loss values and their gradients are then formed on the selected pairs only, which is exactly what
  autograd propagates through the reference's `.min(1)`.  Host-side random selections use the same np.random calls in the same
  order as the reference (sel0, sel1, then pos_sel)."""

  def __init__(self, pos_thresh=0.1, neg_thresh=1.4, rng=np.random):
    self.pos_thresh, self.neg_thresh, self.rng = pos_thresh, neg_thresh, rng

  def __call__(self, F0, F1, positive_pairs, num_pos=5192, num_hn_samples=2048, selections=None):
    _lib.require_cuda(F0, F1)
    dev = F0.device
    N0, N1 = len(F0), len(F1)
    pp = torch.as_tensor(np.asarray(positive_pairs.cpu() if isinstance(positive_pairs, torch.Tensor) else positive_pairs), dtype=torch.int64)
    n_pairs = len(pp)
    hash_seed = max(N0, N1)
    if selections is None:
      sel0 = self.rng.choice(N0, min(N0, num_hn_samples), replace=False)
      sel1 = self.rng.choice(N1, min(N1, num_hn_samples), replace=False)
      pos_sel = self.rng.choice(n_pairs, num_pos, replace=False) if n_pairs > num_pos else None
    else:
      sel0, sel1, pos_sel = selections
    sample = pp[torch.as_tensor(pos_sel)] if pos_sel is not None else pp
    to_d = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64).to(dev)
    sel0_d, sel1_d = to_d(sel0), to_d(sel1)
    i0, i1 = sample[:, 0].to(dev), sample[:, 1].to(dev)
    with torch.no_grad():     # hardest negatives: arg-min over the sub-sampled other cloud (K4, row-indirect, squared L2 ordering)
      F0d, F1d = F0.detach().float().contiguous(), F1.detach().float().contiguous()
      j01 = ops.nn_search(F0d, F1d, both=False, a_rows=i0.contiguous(), b_rows=sel1_d)[0]
      j10 = ops.nn_search(F1d, F0d, both=False, a_rows=i1.contiguous(), b_rows=sel0_d)[0]
      D01ind, D10ind = sel1_d[j01], sel0_d[j10]
      pos_keys = torch.sort(pp[:, 0].to(dev) + pp[:, 1].to(dev) * hash_seed).values          # _hash(positive_pairs, hash_seed)
      isin = lambda k: pos_keys[torch.searchsorted(pos_keys, k).clamp_(max=len(pos_keys) - 1)] == k
      mask0 = ~isin(i0 + D01ind * hash_seed)
      mask1 = ~isin(D10ind + i1 * hash_seed)
    posF0, posF1 = F0[i0], F1[i1]
    D01min = torch.sqrt((posF0 - F1[D01ind]).pow(2).sum(1) + 1e-7)       # pdist(..., 'L2') at the arg-min (lib/metrics.py:22-29)
    D10min = torch.sqrt((posF1 - F0[D10ind]).pow(2).sum(1) + 1e-7)
    relu = torch.nn.functional.relu
    pos_loss = relu((posF0 - posF1).pow(2).sum(1) - self.pos_thresh)
    neg0 = relu(self.neg_thresh - D01min[mask0]).pow(2)
    neg1 = relu(self.neg_thresh - D10min[mask1]).pow(2)
    return pos_loss.mean(), (neg0.mean() + neg1.mean()) / 2
