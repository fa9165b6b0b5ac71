"""ORACLE (test infrastructure, NOT product code): feature-space nearest neighbour / mutual NN on CPU.

Restates, in torch-CPU / numpy:
  pdist          /root/reference/lib/metrics.py:22-29      (direct-difference form, 'L2' adds sqrt(.+1e-7))
  find_nn_gpu    /root/reference/lib/eval.py:18-48         (chunks of nn_max_n rows, min over dim 1)
  calculate_M    /root/reference/generalization_ETH/evaluate.py:63-77   (mutual NN, ascending i)
  find_corr      /root/reference/scripts/test_kitti.py:29-43
  match_pair     /root/reference/scripts/SC2_PCR/SC2_PCR.py:276-302 (literal formula: sqrt(2 - 2 F0 F1^T + 1e-6), argmin)
Pinned against the reference's own functions imported in the build container
(tests/test_reference_live.py; golden vectors tests/golden/nn.npz, tests/test_oracle_golden.py, generator tests/golden/make_golden.py).
"""
import numpy as np
import torch


def pdist(A, B, dist_type="L2"):
  D2 = torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)
  if dist_type == "L2":
    return torch.sqrt(D2 + 1e-7)
  if dist_type == "SquareL2":
    return D2
  raise NotImplementedError("Not implemented")


def find_nn(F0, F1, nn_max_n=-1, return_distance=False, dist_type="SquareL2"):
  """lib/eval.py:18-48.  Returns int64 [N] (and [N,1] distances)."""
  N = len(F0)
  step = nn_max_n if nn_max_n > 1 else max(N, 1)
  dists, inds = [], []
  for s in range(0, N, step):
    d = pdist(F0[s:s + step], F1, dist_type)
    md, ind = d.min(dim=1)
    dists.append(md.unsqueeze(1))
    inds.append(ind)
  inds = torch.cat(inds) if inds else torch.zeros(0, dtype=torch.int64)
  dists = torch.cat(dists) if dists else torch.zeros((0, 1))
  return (inds, dists) if return_distance else inds


def mutual_nn(F0, F1, chunk=1024):
  """generalization_ETH/evaluate.py:63-77 with exact brute-force NN instead of a KD-tree:
  keep [i, nn01[i]] iff nn10[nn01[i]] == i; rows ascending in i.  Returns (pairs [K,2], nn01, nn10)."""
  nn01 = find_nn(F0, F1, nn_max_n=chunk).numpy()
  nn10 = find_nn(F1, F0, nn_max_n=chunk).numpy()
  i = np.arange(len(nn01))
  keep = nn10[nn01] == i if len(nn01) else np.zeros(0, bool)
  return np.stack([i[keep], nn01[keep]], 1).astype(np.int64), nn01, nn10


def nn_margin(F0, F1, chunk=1024):
  """second-best minus best squared distance per row: the documented near-tie test uses this."""
  out = []
  for s in range(0, len(F0), chunk):
    d = pdist(F0[s:s + chunk], F1, "SquareL2")
    v, _ = torch.topk(d, k=min(2, d.shape[1]), dim=1, largest=False)
    out.append(v[:, -1] - v[:, 0] if d.shape[1] > 1 else torch.full((len(d),), float("inf")))
  return torch.cat(out) if out else torch.zeros(0)


def find_corr(xyz0, xyz1, F0, F1, subsample_size=-1, rng=np.random):
  """scripts/test_kitti.py:29-43 (host RNG passed in so seeded runs are reproducible)."""
  subsample = len(F0) > subsample_size
  if subsample_size > 0 and subsample:
    N0, N1 = min(len(F0), subsample_size), min(len(F1), subsample_size)
    inds0 = rng.choice(len(F0), N0, replace=False)
    inds1 = rng.choice(len(F1), N1, replace=False)
    F0, F1 = F0[inds0], F1[inds1]
  nn_inds = find_nn(F0, F1, nn_max_n=500).numpy()
  if subsample_size > 0 and subsample:
    return xyz0[inds0], xyz1[inds1[nn_inds]]
  return xyz0, xyz1[nn_inds]


def match_pair(src_keypts, tgt_keypts, src_features, tgt_features):
  """scripts/SC2_PCR/SC2_PCR.py:276-302 with num_node == 'all' (inputs [1,N,3] / [1,N,C]).
  Returns (src_keypts_corr, tgt_keypts_corr, source_idx)."""
  distance = torch.sqrt(2 - 2 * (src_features[0] @ tgt_features[0].T) + 1e-6)
  source_idx = torch.argmin(distance, dim=1)
  return src_keypts, tgt_keypts[:, source_idx], source_idx
