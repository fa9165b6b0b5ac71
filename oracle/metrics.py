"""ORACLE (test infrastructure, NOT product code): evaluation metrics and FCGF's hardest-contrastive loss on CPU.

Restates
  rte_rre / success      /root/reference/scripts/test_kitti.py:188-204
  evaluate_hit_ratio     /root/reference/lib/trainer.py:406-409 (+ :66-70 apply_transform)
  hardest_contrastive    /root/reference/lib/trainer.py:412-462, util/misc.py:43-55 (_hash)
Pinned: the loss against the reference's own HardestContrastiveLossTrainer.contrastive_hardest_negative_loss run unbound in the
build container (tests/golden/hardest_loss.npz, generator tests/golden/make_golden_metrics.py); the metric formulas are literal.
"""
import numpy as np
import torch
import torch.nn.functional as F


def rte_rre(T_est: torch.Tensor, T_gth: torch.Tensor):
  """float32 tensors [4,4] -> (rte, rre in radians), test_kitti.py:188-192"""
  rte = np.linalg.norm(T_est[:3, 3] - T_gth[:3, 3])
  trace_matrix = T_est[:3, :3].t() @ T_gth[:3, :3]
  trace_matrix[[0, 1, 2], [0, 1, 2]] = torch.min(torch.ones(3), trace_matrix[[0, 1, 2], [0, 1, 2]])
  rre = np.arccos((np.trace(trace_matrix) - 1) / 2)
  return float(rte), float(rre)


def evaluate_hit_ratio(xyz0, xyz1, T_gth, thresh=0.1):
  T = T_gth.float()
  x = xyz0 @ T[:3, :3].t() + T[:3, 3]
  dist = torch.sqrt(((x - xyz1) ** 2).sum(1) + 1e-6)
  return (dist < thresh).float().mean().item()


def _hash(arr, M):
  if isinstance(arr, np.ndarray):
    N, D = arr.shape
  else:
    N, D = len(arr[0]), len(arr)
  h = np.zeros(N, dtype=np.int64)
  for d in range(D):
    h += (arr[:, d] if isinstance(arr, np.ndarray) else arr[d]) * M ** d
  return h


def pdist(A, B, dist_type="L2"):
  D2 = torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)
  return torch.sqrt(D2 + 1e-7) if dist_type == "L2" else D2


def hardest_contrastive(F0, F1, positive_pairs, sel0, sel1, pos_sel, pos_thresh=0.1, neg_thresh=1.4):
  """lib/trainer.py:412-462 with the random selections passed in (same draw order as the reference: sel0, sel1, pos_sel)"""
  N0, N1 = len(F0), len(F1)
  hash_seed = max(N0, N1)
  positive_pairs = torch.as_tensor(positive_pairs)
  sample = positive_pairs[pos_sel] if pos_sel is not None else positive_pairs
  subF0, subF1 = F0[sel0], F1[sel1]
  i0, i1 = sample[:, 0].long(), sample[:, 1].long()
  posF0, posF1 = F0[i0], F1[i1]
  D01min, D01ind = pdist(posF0, subF1).min(1)
  D10min, D10ind = pdist(posF1, subF0).min(1)
  pos_keys = _hash(np.asarray(positive_pairs, dtype=np.int64), hash_seed)
  D01ind = np.asarray(sel1)[D01ind.numpy()]
  D10ind = np.asarray(sel0)[D10ind.numpy()]
  mask0 = torch.from_numpy(np.logical_not(np.isin(_hash([i0.numpy(), D01ind], hash_seed), pos_keys)))
  mask1 = torch.from_numpy(np.logical_not(np.isin(_hash([D10ind, i1.numpy()], hash_seed), pos_keys)))
  pos_loss = F.relu((posF0 - posF1).pow(2).sum(1) - pos_thresh)
  neg0 = F.relu(neg_thresh - D01min[mask0]).pow(2)
  neg1 = F.relu(neg_thresh - D10min[mask1]).pow(2)
  return pos_loss.mean(), (neg0.mean() + neg1.mean()) / 2
