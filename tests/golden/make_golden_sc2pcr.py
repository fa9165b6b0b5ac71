"""Regenerates tests/golden/sc2pcr.npz from the REFERENCE's own scripts/SC2_PCR/SC2_PCR.py (build container only).

    python tests/golden/make_golden_sc2pcr.py

Matcher.SC2_PCR (the registration step after feature matching, scripts/test_kitti.py:180-182) is run UNMODIFIED on CPU
tensors with the KITTI configuration (scripts/SC2_PCR/config_json/config_KITTI.json) on seeded putative correspondences:
a rigid motion, inliers with noise, and a majority of outliers."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import sc2pcr as osc  # noqa: E402
from oracle.refshim import import_reference  # noqa: E402
import oracle.me_cpu as OME  # noqa: E402


def correspondences(seed, n=1200, inlier_ratio=0.3, noise=0.03):
  rng = np.random.RandomState(seed)
  src = np.concatenate([rng.uniform(-40, 40, (n, 2)), rng.uniform(-2, 3, (n, 1))], 1)
  yaw, pitch = rng.uniform(-0.5, 0.5), rng.uniform(-0.05, 0.05)
  Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
  Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
  T = np.eye(4); T[:3, :3] = Rz @ Ry; T[:3, 3] = [rng.uniform(5, 15), rng.uniform(-2, 2), rng.uniform(-0.3, 0.3)]
  tgt = src @ T[:3, :3].T + T[:3, 3] + rng.normal(0, noise, src.shape)
  out = rng.rand(n) > inlier_ratio
  tgt[out] = np.concatenate([rng.uniform(-40, 40, (out.sum(), 2)), rng.uniform(-2, 3, (out.sum(), 1))], 1)
  return src.astype(np.float32), tgt.astype(np.float32), T


def main():
  (mod,) = import_reference(OME, ("scripts.SC2_PCR.SC2_PCR",))
  cfg = osc.SC2Config()
  m = mod.Matcher(inlier_threshold=cfg.inlier_threshold, num_node=8000, use_mutual=False, d_thre=cfg.d_thre,
                  num_iterations=cfg.num_iterations, ratio=cfg.ratio, nms_radius=cfg.nms_radius, max_points=cfg.max_points,
                  k1=cfg.k1, k2=cfg.k2)
  out = {}
  for case, (seed, n, ratio) in enumerate([(1, 1200, 0.3), (2, 800, 0.15), (3, 1500, 0.6)]):
    src, tgt, T = correspondences(seed, n, ratio)
    with torch.no_grad():
      ref = m.SC2_PCR(torch.from_numpy(src)[None], torch.from_numpy(tgt)[None])[0].numpy()
      mine = osc.sc2_pcr(torch.from_numpy(src)[None], torch.from_numpy(tgt)[None], cfg)[0].numpy()
    err_ref = np.abs(ref - T).max()
    print(f"case {case}: n={n} inliers={ratio:.2f}  |ref - gt|max={err_ref:.4f}  |oracle - ref|max={np.abs(mine - ref).max():.2e}")
    assert np.abs(mine - ref).max() < 1e-4
    out.update({f"c{case}_src": src, f"c{case}_tgt": tgt, f"c{case}_gt": T, f"c{case}_trans": ref})
  np.savez_compressed(os.path.join(HERE, "sc2pcr.npz"), **out)


if __name__ == "__main__":
  main()
