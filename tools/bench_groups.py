"""Positive-group construction timing: gcl_b200.groups.colocation_groups on the GPU vs the reference's
algorithm on the host (util/pointcloud.py:69-132: a Python loop over the centre points through 1 + J KD-trees; Open3D is
not installable here, so the KD-tree is scipy's cKDTree -- a bounded sample of centre points, extrapolated).
One sample = 1 centre scan + 6 neighbour scans of a KITTI-shape scene, each voxel-downsampled at 0.3 m like the loader,
radius 0.45 m, K = 5 (lib/colocation_data_loader.py:379-394)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcl_b200 import groups as gg, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
VOXEL, RADIUS, K, J = 0.3, 0.45, 5, 6
rng = np.random.RandomState(0)
scene = synth.Scene(11)
clouds, trans = [], []
for j in range(J + 1):
  tx, ty = (4.0 * j * (1 if j % 2 else -1), rng.uniform(-1, 1)) if j else (0.0, 0.0)
  yaw = rng.uniform(-0.2, 0.2) if j else 0.0
  T = np.eye(4)                                        # sensor j -> centre frame (the centre sensor sits at the origin)
  T[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
  T[:3, 3] = [tx, ty, 0.0]
  x = torch.from_numpy(synth.cast(scene, synth.KITTI, pose_xy_yaw=(tx, ty, yaw), seed=j)).to(dev)    # sensor frame
  cm, umap = ops.voxelize(x, VOXEL)
  clouds.append(x[umap].contiguous())
  trans.append(T)
centre, nbs, Ts = clouds[0], clouds[1:], trans[1:]

for _ in range(3):
  out = gg.colocation_groups(centre, nbs, Ts, VOXEL, RADIUS, K)
torch.cuda.synchronize()
t0 = time.perf_counter()
N = 10
for _ in range(N):
  out = gg.colocation_groups(centre, nbs, Ts, VOXEL, RADIUS, K)
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / N * 1e3

from scipy.spatial import cKDTree  # noqa: E402
C = centre.cpu().numpy().astype(np.float64)
nbT = [(x.cpu().numpy().astype(np.float64) @ T[:3, :3].T + T[:3, 3]) for x, T in zip(nbs, Ts)]
t0 = time.perf_counter()
trees = [cKDTree(C)] + [cKDTree(x) for x in nbT]
sample = min(3000, len(C))
for i in range(sample):
  p = C[i]
  for tr, pts in zip(trees, [C] + nbT):
    idx = np.asarray(tr.query_ball_point(p, RADIUS), dtype=np.int64)
    if len(idx):
      d = ((pts[idx] - p) ** 2).sum(1)
      idx = idx[np.lexsort((idx, d))][:K]
cpu_ms = (time.perf_counter() - t0) / sample * len(C) * 1e3
print(json.dumps({"centre_points": int(len(C)), "neighbour_clouds": J, "neighbour_points": int(sum(len(x) for x in nbs)),
                  "groups": int(out[0].numel()), "members": int(out[1].numel()), "gpu_ms_per_sample": round(gpu_ms, 3),
                  "cpu_kdtree_loop_ms_per_sample": round(cpu_ms, 1), "cpu_sample": f"{sample} of {len(C)} centre points, scipy cKDTree, 1 core",
                  "speedup": round(cpu_ms / gpu_ms, 1)}))
