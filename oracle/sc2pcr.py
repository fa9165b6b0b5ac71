"""ORACLE (test infrastructure only -- never imported by the product path): SC2-PCR registration from correspondences,
restated from /root/reference/scripts/SC2_PCR/SC2_PCR.py:304-381 (`Matcher.SC2_PCR`), :34-58 (`pick_seeds`), :60-168
(`cal_seed_trans`), :170-196 (`cal_leading_eigenvector`), :235-274 (`post_refinement`) and scripts/SC2_PCR/common.py:7-45
(`rigid_transform_3d`).  It is the step right after feature matching in scripts/test_kitti.py:180-182 (the
next row to be built on the GPU).

Parity status: pinned against the reference's own class executed on CPU tensors in the build container
(tests/golden/make_golden_sc2pcr.py -> tests/golden/sc2pcr.npz; estimated transforms agree to float32 round-off).
Batch size 1, like the reference's own asserts (:46, :247).
"""
from dataclasses import dataclass

import torch


@dataclass
class SC2Config:          # scripts/SC2_PCR/config_json/config_KITTI.json
  inlier_threshold: float = 0.6
  d_thre: float = 0.1
  num_iterations: int = 20
  ratio: float = 0.2
  nms_radius: float = 0.6
  max_points: int = 8000
  k1: int = 30
  k2: int = 20


def leading_eigenvector(M: torch.Tensor, num_iterations: int) -> torch.Tensor:
  """power iteration, SC2_PCR.py:179-190.  M [B,n,n] -> [B,n]"""
  v = torch.ones_like(M[:, :, 0:1])
  last = v
  for _ in range(num_iterations):
    v = torch.bmm(M, v)
    v = v / (torch.norm(v, dim=1, keepdim=True) + 1e-6)
    if torch.allclose(v, last):
      break
    last = v
  return v.squeeze(-1)


def rigid_transform_3d(A, B, weights=None, weight_threshold=0.0):
  """weighted Kabsch, common.py:7-45.  A, B [B,n,3], weights [B,n] -> [B,4,4] mapping A onto B"""
  bs = A.shape[0]
  if weights is None:
    weights = torch.ones_like(A[:, :, 0])
  weights = weights.clone()
  weights[weights < weight_threshold] = 0
  wsum = torch.sum(weights, dim=1, keepdim=True)[:, :, None] + 1e-6
  cA = torch.sum(A * weights[:, :, None], dim=1, keepdim=True) / wsum
  cB = torch.sum(B * weights[:, :, None], dim=1, keepdim=True) / wsum
  Am, Bm = A - cA, B - cB
  H = Am.permute(0, 2, 1) @ torch.diag_embed(weights) @ Bm
  U, S, V = torch.svd(H)                     # common.py:36 names the third factor "Vt" but torch.svd returns V
  delta = torch.det(V @ U.permute(0, 2, 1))
  eye = torch.eye(3)[None].repeat(bs, 1, 1)
  eye[:, -1, -1] = delta
  R = V @ eye @ U.permute(0, 2, 1)
  t = cB.permute(0, 2, 1) - R @ cA.permute(0, 2, 1)
  T = torch.eye(4)[None].repeat(bs, 1, 1)
  T[:, :3, :3] = R
  T[:, :3, 3:4] = t
  return T


def transform(pts, T):
  return (T[:, :3, :3] @ pts.permute(0, 2, 1) + T[:, :3, 3:4]).permute(0, 2, 1)


def pick_seeds(dists, scores, R, max_num):
  """parallel non-maximum suppression, SC2_PCR.py:34-58"""
  rel = (scores.T >= scores).bool() | (dists[0] >= R).bool()
  is_local_max = rel.min(-1)[0].float()
  order = torch.argsort(scores * is_local_max, dim=1, descending=True)
  return order[:, 0:max_num]


def cal_seed_trans(seeds, SC2_measure, src, tgt, cfg: SC2Config):
  """SC2_PCR.py:60-168: two-stage consensus sets around every seed, weighted SVD per seed, best hypothesis by inlier count"""
  bs, num_channels = SC2_measure.shape[0], SC2_measure.shape[2]
  k1, k2 = cfg.k1, cfg.k2
  if k1 > num_channels:
    k1 = k2 = 4
  knn_idx = torch.argsort(SC2_measure, dim=2, descending=True)[:, :, 0:k1]
  idx = knn_idx.contiguous().view(bs, -1)[:, :, None].expand(-1, -1, 3)
  src_knn = src.gather(1, idx).view(bs, -1, k1, 3)
  tgt_knn = tgt.gather(1, idx).view(bs, -1, k1, 3)
  sd = ((src_knn[:, :, :, None, :] - src_knn[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
  td = ((tgt_knn[:, :, :, None, :] - tgt_knn[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
  hard = (torch.abs(sd - td) < cfg.d_thre).float()
  local = torch.matmul(hard[:, :, :1, :], hard)
  fine = torch.argsort(local, dim=3, descending=True)[:, :, :, 0:k2]
  num = fine.shape[1]
  fine = fine.contiguous().view(bs, num, -1)[:, :, :, None].expand(-1, -1, -1, 3)
  src_f = src_knn.gather(2, fine).view(bs, -1, k2, 3)
  tgt_f = tgt_knn.gather(2, fine).view(bs, -1, k2, 3)
  sd = ((src_f[:, :, :, None, :] - src_f[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
  td = ((tgt_f[:, :, :, None, :] - tgt_f[:, :, None, :, :]) ** 2).sum(-1) ** 0.5
  cross = torch.abs(sd - td)
  M = torch.clamp(1 - cross ** 2 / cfg.d_thre ** 2, min=0).view(-1, k2, k2)        # :123-126 (the soft measure is what is kept)
  ar = torch.arange(k2)
  M[:, ar, ar] = 0
  w = leading_eigenvector(M, cfg.num_iterations).view(bs, -1, k2)
  w = w / (torch.sum(w, dim=-1, keepdim=True) + 1e-6)
  T = rigid_transform_3d(src_f.view(-1, k2, 3), tgt_f.view(-1, k2, 3), w.view(-1, k2)).view(bs, -1, 4, 4)
  pred = torch.einsum('bsnm,bmk->bsnk', T[:, :, :3, :3], src.permute(0, 2, 1)) + T[:, :, :3, 3:4]
  L2 = torch.norm(pred.permute(0, 1, 3, 2) - tgt[:, None, :, :], dim=-1)
  fitness = torch.sum((L2 < cfg.inlier_threshold).float(), dim=-1)
  best = fitness.argmax(dim=1)
  return T.gather(1, best[:, None, None, None].expand(-1, -1, 4, 4)).squeeze(1)


def post_refinement(T, src, tgt, it_num, cfg: SC2Config):
  """SC2_PCR.py:235-274"""
  thr = 0.10 if cfg.inlier_threshold == 0.10 else 1.2
  prev = 0
  for _ in range(it_num):
    L2 = torch.norm(transform(src, T) - tgt, dim=-1)
    inl = (L2 < thr)[0]
    n = int(inl.sum())
    if abs(n - prev) < 1:
      break
    prev = n
    T = rigid_transform_3d(src[:, inl, :], tgt[:, inl, :], weights=(1 / (1 + (L2 / thr) ** 2))[:, inl])
  return T


def sc2_pcr(src_keypts: torch.Tensor, tgt_keypts: torch.Tensor, cfg: SC2Config = SC2Config()) -> torch.Tensor:
  """SC2_PCR.py:304-381.  src/tgt [1,n,3] putative correspondences (row i of src matches row i of tgt) -> [1,4,4]"""
  src, tgt = src_keypts.float(), tgt_keypts.float()
  n = src.shape[1]
  if n > cfg.max_points:
    src, tgt, n = src[:, :cfg.max_points], tgt[:, :cfg.max_points], cfg.max_points
  sd = torch.norm(src[:, :, None, :] - src[:, None, :, :], dim=-1)
  td = torch.norm(tgt[:, :, None, :] - tgt[:, None, :, :], dim=-1)
  cross = torch.abs(sd - td)
  SC = torch.clamp(1.0 - cross ** 2 / cfg.d_thre ** 2, min=0)
  hard = (cross < cfg.d_thre).float()
  conf = leading_eigenvector(SC, cfg.num_iterations)
  seeds = pick_seeds(sd, conf, cfg.nms_radius, int(n * cfg.ratio))
  tight = (cross < cfg.d_thre / 2).float()
  seed_hard = hard.gather(1, seeds[:, :, None].expand(-1, -1, n))
  seed_tight = tight.gather(1, seeds[:, :, None].expand(-1, -1, n))
  SC2 = torch.matmul(seed_tight, tight) * seed_hard
  T = cal_seed_trans(seeds, SC2, src, tgt, cfg)
  return post_refinement(T, src, tgt, 20, cfg)
