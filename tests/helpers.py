"""Shared test helpers (CPU side)."""
import numpy as np
import torch


def numpy_seeded_weights(model, seed=0):
  """Fill every parameter / BN statistic from numpy's RandomState (bit-stable across platforms and torch versions),
  in state_dict order.  Used so that golden feature vectors can be regenerated anywhere."""
  rng = np.random.RandomState(seed)
  sd = model.state_dict()
  for k, v in sd.items():
    if k.endswith("num_batches_tracked"):
      continue
    shape = tuple(v.shape)
    if k.endswith("running_var"):
      a = rng.uniform(0.5, 1.5, shape)
    elif k.endswith("running_mean"):
      a = rng.normal(0, 0.1, shape)
    elif k.endswith("bn.weight"):
      a = rng.uniform(0.5, 1.5, shape)
    elif k.endswith("bn.bias") or k.endswith(".bias"):
      a = rng.normal(0, 0.1, shape)
    else:  # conv kernels [K, Cin, Cout] or [Cin, Cout]
      fan = int(np.prod(shape[:-1]))
      a = rng.uniform(-1, 1, shape) / np.sqrt(fan)
    sd[k] = torch.from_numpy(a.astype(np.float32))
  model.load_state_dict(sd)
  return model


def small_cloud(seed=0, n=2500, extent=7.0):
  rng = np.random.RandomState(seed)
  return np.concatenate([rng.uniform(-extent, extent, (n, 2)), rng.normal(0, 0.5, (n, 1))], 1).astype(np.float32)


def sc2pcr_correspondences(seed, n=1200, inlier_ratio=0.3, noise=0.03):
  """seeded putative correspondences: a rigid motion, noisy inliers and a majority of outliers (same generator as
  tests/golden/make_golden_sc2pcr.py)"""
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
