"""Feature-space nearest-neighbour / correspondence helpers on libgclb200's K4 kernel.

Same names, arguments and return conventions as the reference helpers they replace:
  find_nn_gpu   /root/reference/lib/eval.py:18-48        -> CPU int64 [N] (+ CPU [N,1] distances)
  pdist         /root/reference/lib/metrics.py:22-29
  find_corr     /root/reference/scripts/test_kitti.py:29-43
  mutual_nn     /root/reference/generalization_ETH/evaluate.py:63-77 (calculate_M): [K,2] (i, nn01[i]) with
                nn10[nn01[i]] == i, ascending i
  match_pair    /root/reference/scripts/SC2_PCR/SC2_PCR.py:276-302 (Matcher.match_pair): putative correspondences
                [1,N,3] x 2 for the SC2-PCR registration
The N x M matrix is never materialised and there is one device->host copy per call instead of one per chunk.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import GclbError


def pdist(A, B, dist_type="L2"):
  """Full distance matrix (API parity only; the hot path uses the fused kernels and never builds it)."""
  ops.require_cuda(A, B)
  D2 = torch.sum((A.unsqueeze(1) - B.unsqueeze(0)).pow(2), 2)
  if dist_type == "L2":
    return torch.sqrt(D2 + 1e-7)
  if dist_type == "SquareL2":
    return D2
  raise NotImplementedError("Not implemented")


def nn_device(F0, F1, dist_type="SquareL2"):
  """device-resident NN: (idx int64 [N], dist float32 [N]) -- no host sync."""
  idx, d, _, _, _, _, _ = ops.nn_search(F0.detach().float(), F1.detach().float(), both=False)
  if dist_type == "L2":
    d = torch.sqrt(d + 1e-7)
  elif dist_type != "SquareL2":
    raise NotImplementedError("Not implemented")
  return idx, d


def find_nn_gpu(F0, F1, nn_max_n=-1, return_distance=False, dist_type="SquareL2"):
  """nn_max_n (the reference's memory-saving chunk size) is accepted and ignored: the fused kernel has no
  N x M temporary, so chunking is unnecessary and the result is identical to the unchunked one."""
  idx, d = nn_device(F0, F1, dist_type)
  if return_distance:
    return idx.cpu(), d.unsqueeze(1).cpu()
  return idx.cpu()


def mutual_nn_device(F0, F1, a_ptr=None, b_ptr=None):
  """Batched mutual NN, everything stays on the device.
  Returns (pairs int64 [*,2] (first pair_ptr[-1] rows valid), pair_ptr int64 [n_pairs+1], idx01, idx10)."""
  idx01, d01, idx10, d10, a_dev, b_dev, ws = ops.nn_search(F0.detach().float(), F1.detach().float(), a_ptr, b_ptr,
                                                           both=True)
  pairs, pair_ptr = ops.mutual_filter(idx01, idx10, a_dev, b_dev, ws)
  return pairs, pair_ptr, idx01, idx10


def mutual_nn(source_desc, target_desc):
  """calculate_M: numpy [K,2] of mutually nearest (source i, target j) pairs, ascending i."""
  to_t = lambda x: x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
  s, t = to_t(source_desc), to_t(target_desc)
  if not s.is_cuda:
    if not torch.cuda.is_available():
      raise GclbError("mutual_nn needs a CUDA device: gcl_b200 has no CPU fallback")
    s, t = s.cuda(), t.cuda()
  pairs, pair_ptr, _, _ = mutual_nn_device(s, t)
  k = int(pair_ptr[-1].item())
  return pairs[:k].cpu().numpy()


def find_corr(xyz0, xyz1, F0, F1, subsample_size=-1, rng=np.random):
  subsample = len(F0) > subsample_size
  if subsample_size > 0 and subsample:
    N0, N1 = min(len(F0), subsample_size), min(len(F1), subsample_size)
    inds0 = rng.choice(len(F0), N0, replace=False)
    inds1 = rng.choice(len(F1), N1, replace=False)
    F0, F1 = F0[inds0], F1[inds1]
  nn_inds = find_nn_gpu(F0, F1, nn_max_n=500)
  if subsample_size > 0 and subsample:
    return xyz0[inds0], xyz1[inds1[nn_inds]]
  return xyz0, xyz1[nn_inds]


def match_pair(src_keypts, tgt_keypts, src_features, tgt_features, num_node="all", rng=np.random):
  """Matcher.match_pair (scripts/SC2_PCR/SC2_PCR.py:276-302): for every (optionally sub-sampled) source keypoint the
  target keypoint whose descriptor minimises `sqrt(2 - 2 * F0 @ F1.T + 1e-6)`.
  Inputs carry the reference's leading batch dimension of 1: keypts [1,N,3], features [1,N,C]; returns
  (src_keypts_corr [1,N,3], tgt_keypts_corr [1,N,3]) on the device.
  The reference's distance is a monotone function of -F0.F1; on the UNIT-NORM descriptors this path produces
  (model/resunet.py:226-230; SC2-PCR is only ever fed normalised FCGF features) that is the squared-L2 ordering the
  fused K4 kernel computes, so no N x M matrix is built; ties / near-ties resolve to the smaller index."""
  ops.require_cuda(src_keypts, tgt_keypts, src_features, tgt_features)
  if src_features.dim() != 3 or src_features.shape[0] != 1 or tgt_features.shape[0] != 1:
    raise GclbError("match_pair expects a leading batch dimension of 1 like the reference (bs == 1)")
  n_src, n_tgt = src_features.shape[1], tgt_features.shape[1]
  if num_node != "all":      # :282-284 -- note: sampling WITH replacement, like the reference
    si = torch.as_tensor(rng.choice(n_src, num_node), device=src_features.device)
    ti = torch.as_tensor(rng.choice(n_tgt, num_node), device=src_features.device)
    src_keypts, tgt_keypts = src_keypts[:, si], tgt_keypts[:, ti]
    src_features, tgt_features = src_features[:, si], tgt_features[:, ti]
  idx, _ = nn_device(src_features[0].contiguous(), tgt_features[0].contiguous())
  return src_keypts, tgt_keypts[:, idx]
