"""Synthetic LiDAR scans (host side, numpy): the benchmark / test input generator.

No dataset or network is available, so every measured workload is a seeded ray-cast of a synthetic street
scene shaped like the reference's inputs: KITTI = 64 beams x 2083 azimuth steps
(lib/complement_data_loader.py:358-361 reads ~120k returns per .bin), nuScenes = 32 beams x 1090 steps.
The scene is a ground plane plus axis-aligned boxes (buildings, cars); range noise sigma = 2 cm.
"""
from __future__ import annotations

import numpy as np

KITTI = dict(beams=64, elev_deg=(2.0, -24.8), azimuth_steps=2083, height=1.73, max_range=120.0)
NUSCENES = dict(beams=32, elev_deg=(10.67, -30.67), azimuth_steps=1090, height=1.84, max_range=100.0)


class Scene:
  """Ground plane z=0 + `n_boxes` axis-aligned boxes scattered over [-extent, extent]^2."""

  def __init__(self, seed: int = 0, n_boxes: int = 80, extent: float = 90.0):
    rng = np.random.RandomState(seed)
    c = rng.uniform(-extent, extent, size=(n_boxes, 2))
    # keep a corridor along the x axis free so that a translated second pose is never inside a box
    c[:, 1] = np.where(np.abs(c[:, 1]) < 6.0, np.sign(c[:, 1] + 1e-9) * (6.0 + np.abs(c[:, 1])), c[:, 1])
    big = rng.rand(n_boxes) < 0.5
    sx = np.where(big, rng.uniform(6, 25, n_boxes), rng.uniform(1.6, 4.5, n_boxes))
    sy = np.where(big, rng.uniform(6, 25, n_boxes), rng.uniform(1.6, 2.2, n_boxes))
    sz = np.where(big, rng.uniform(4, 18, n_boxes), rng.uniform(1.4, 2.0, n_boxes))
    self.lo = np.stack([c[:, 0] - sx / 2, c[:, 1] - sy / 2, np.zeros(n_boxes)], 1)
    self.hi = np.stack([c[:, 0] + sx / 2, c[:, 1] + sy / 2, sz], 1)
    inside = (np.abs(c[:, 1]) - sy / 2) < 4.0
    self.lo, self.hi = self.lo[~inside], self.hi[~inside]


def cast(scene: Scene, sensor=KITTI, pose_xy_yaw=(0.0, 0.0, 0.0), seed: int = 0, noise: float = 0.02,
         dtype=np.float32) -> np.ndarray:
  """Ray-cast one scan; returns [P,3] points in the SENSOR frame (like a KITTI .bin), P ~ 0.9 * rays."""
  rng = np.random.RandomState(seed)
  nb, na = sensor["beams"], sensor["azimuth_steps"]
  el = np.deg2rad(np.linspace(sensor["elev_deg"][0], sensor["elev_deg"][1], nb))
  az = np.linspace(-np.pi, np.pi, na, endpoint=False)
  x0, y0, yaw = pose_xy_yaw
  ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
  d_s = np.stack([ce * np.cos(az)[None], ce * np.sin(az)[None], np.broadcast_to(se, (nb, na))], -1).reshape(-1, 3)
  cy, sy = np.cos(yaw), np.sin(yaw)
  Rw = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
  d = d_s @ Rw.T
  o = np.array([x0, y0, sensor["height"]])
  t = np.full(len(d), np.inf)
  dz = d[:, 2]
  down = dz < -1e-6
  t[down] = -o[2] / dz[down]
  inv = 1.0 / np.where(np.abs(d) < 1e-12, 1e-12, d)
  for lo, hi in zip(scene.lo, scene.hi):         # slab test, one box at a time to bound memory
    t0, t1 = (lo - o) * inv, (hi - o) * inv
    tn = np.minimum(t0, t1).max(1)
    tf = np.maximum(t0, t1).min(1)
    hit = (tn <= tf) & (tf > 0) & (tn > 0)
    t = np.where(hit & (tn < t), tn, t)
  ok = t < sensor["max_range"]
  r = t[ok] + rng.randn(int(ok.sum())) * noise
  return (d_s[ok] * r[:, None]).astype(dtype)


def scan_pair(scene_seed: int = 0, pair_seed: int = 0, sensor=KITTI, min_d: float = 5.0, max_d: float = 50.0):
  """LoKITTI-style distant pair (config/file_LoKITTI_50.npy holds pairs up to 50 m apart): same scene,
  second pose translated d ~ U[min_d, max_d] along the corridor with a small yaw.
  Returns (xyz0, xyz1, T_gt 4x4 mapping scan-0 coordinates into scan-1's frame)."""
  rng = np.random.RandomState(1000 + pair_seed)
  scene = Scene(scene_seed)
  d = rng.uniform(min_d, max_d)
  yaw = rng.uniform(-0.1, 0.1)
  xyz0 = cast(scene, sensor, (0.0, 0.0, 0.0), seed=2 * pair_seed)
  xyz1 = cast(scene, sensor, (d, 0.0, yaw), seed=2 * pair_seed + 1)
  c, s = np.cos(yaw), np.sin(yaw)
  R1 = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])
  T = np.eye(4)
  T[:3, :3] = R1.T
  T[:3, 3] = -R1.T @ np.array([d, 0.0, 0.0])
  return xyz0, xyz1, T


def dense_surface(n_points: int, seed: int = 0, extent: float = 100.0, dtype=np.float32) -> np.ndarray:
  """Voxel-sweep input (BASELINE config 5): points sampled on rolling terrain + facades over
  [-extent, extent]^2, dense enough that 0.1 m voxels give ~1M occupied cells for n_points ~ 10M."""
  rng = np.random.RandomState(seed)
  n_g = n_points * 2 // 3
  xy = rng.uniform(-extent, extent, size=(n_g, 2))
  z = 1.5 * np.sin(xy[:, 0] / 17.0) * np.cos(xy[:, 1] / 23.0)
  ground = np.concatenate([xy, z[:, None]], 1)
  n_w = n_points - n_g
  wall = rng.randint(0, 64, n_w)
  wx = (wall % 8 - 3.5) * (extent / 4.2)
  u = rng.uniform(-extent / 10, extent / 10, n_w)
  h = rng.uniform(0, 12, n_w)
  along_x = (wall // 8) % 2 == 0
  wy = ((wall // 8) - 3.5) * (extent / 4.2)
  facade = np.stack([np.where(along_x, wx + u, wx), np.where(along_x, wy, wy + u), h], 1)
  return np.concatenate([ground, facade], 0).astype(dtype)
