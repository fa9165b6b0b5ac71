"""Regenerates tests/golden/groups.npz from the REFERENCE's own util/pointcloud.py (build container only).

    python tests/golden/make_golden_groups.py

The reference functions get_matching_indices_colocation / get_matching_indices (util/pointcloud.py:53-132) are imported and
run UNMODIFIED.  Open3D is not installable offline, so `open3d` is replaced by a stand-in with exactly the three things
those functions touch: geometry.PointCloud (points, transform), utility.Vector3dVector and geometry.KDTreeFlann, whose
search_radius_vector_3d is oracle.groups.radius_search (nanoflann's published rule: squared distance < r^2, nearest first).
What the fixture pins is therefore the reference's group logic, not Open3D's KD-tree.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import groups as og  # noqa: E402
from oracle.refshim import import_reference  # noqa: E402
import oracle.me_cpu as OME  # noqa: E402


class _PointCloud:
  def __init__(self):
    self.points = np.zeros((0, 3))

  def transform(self, T):
    self.points = og.transform_points(np.asarray(self.points, np.float64), np.asarray(T, np.float64))
    return self

  def __deepcopy__(self, memo):
    c = _PointCloud()
    c.points = np.array(self.points, dtype=np.float64, copy=True)
    return c


class _KDTreeFlann:
  def __init__(self, pcd):
    self.pts = np.asarray(pcd.points, np.float64)

  def search_radius_vector_3d(self, query, radius):
    idx = og.radius_search(self.pts, np.asarray(query, np.float64), radius)
    d2 = ((self.pts[idx] - np.asarray(query, np.float64)) ** 2).sum(1)
    return len(idx), [int(i) for i in idx], list(d2)


def install_open3d_stub():
  o3d = types.ModuleType("open3d")
  o3d.geometry = types.SimpleNamespace(PointCloud=_PointCloud, KDTreeFlann=_KDTreeFlann)
  o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: np.asarray(a, np.float64))
  sys.modules["open3d"] = o3d


def scene(seed, n_clouds=3, n=1500, extent=12.0, voxel=0.3):
  """centre cloud + neighbour clouds of the same surface seen from displaced sensors; every cloud voxel-downsampled like
  the loader does (lib/colocation_data_loader.py:379-390) and expressed in its own sensor frame."""
  import torch
  rng = np.random.RandomState(seed)
  surf = np.concatenate([rng.uniform(-extent, extent, (4 * n, 2)), rng.normal(0, 0.05, (4 * n, 1)) - 1.7], 1)   # ground
  walls = np.stack([rng.uniform(-extent, extent, 2 * n), np.full(2 * n, 4.0) + rng.normal(0, 0.03, 2 * n),
                    rng.uniform(-1.7, 1.5, 2 * n)], 1)
  world = np.concatenate([surf, walls]).astype(np.float64)
  clouds, trans = [], []
  for j in range(n_clouds + 1):
    yaw = rng.uniform(-0.3, 0.3) if j else 0.0
    t = np.array([rng.uniform(-6, 6), rng.uniform(-2, 2), 0.0]) if j else np.zeros(3)
    R = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = t                     # sensor j -> centre frame
    keep = rng.rand(len(world)) < 0.5
    local = ((world[keep] + rng.normal(0, 0.01, (keep.sum(), 3))) - t) @ R          # world -> sensor j frame (R^T x)
    x = torch.from_numpy(local.astype(np.float32))
    _, sel = OME.utils.sparse_quantize(x / voxel, return_index=True)
    clouds.append(x[sel].numpy())
    trans.append(T)
  return clouds[0], clouds[1:], trans[1:]


def main():
  install_open3d_stub()
  (pc,) = import_reference(OME, ("util.pointcloud",))
  install_open3d_stub()          # import_reference may have installed its empty stub first
  import importlib
  pc = importlib.reload(pc)
  out = {}
  for case, (seed, K, mult) in enumerate([(1, 5, 1.5), (2, None, 1.5), (3, 5, 1.0), (4, 2, 2.0)]):
    centre, nbs, Ts = scene(seed)
    radius = 0.3 * mult
    cp = _PointCloud(); cp.points = centre.astype(np.float64)
    nps = []
    for x in nbs:
      q = _PointCloud(); q.points = x.astype(np.float64); nps.append(q)
    import torch
    g, idx, ff, _ = pc.get_matching_indices_colocation(cp, nps, [torch.from_numpy(x) for x in nbs], Ts, radius, False, K=K)
    og_g, og_i, og_f = og.colocation_groups(centre, nbs, Ts, radius, K)
    assert np.array_equal(np.asarray(g, np.int64), og_g) and np.array_equal(np.asarray(idx, np.int64), og_i)
    assert np.array_equal(np.asarray(ff) > 0.5, og_f)
    mi = np.asarray(pc.get_matching_indices(nps[0], cp, Ts[0], radius, K=K), np.int64).reshape(-1, 2)
    assert np.array_equal(mi, og.matching_indices(nbs[0], centre, Ts[0], radius, K))
    out.update({f"c{case}_centre": centre, f"c{case}_K": np.int64(-1 if K is None else K), f"c{case}_radius": np.float64(radius),
                f"c{case}_group": np.asarray(g, np.int64), f"c{case}_index": np.asarray(idx, np.int64),
                f"c{case}_finest": np.asarray(ff) > 0.5, f"c{case}_pairs": mi})
    for j, (x, T) in enumerate(zip(nbs, Ts)):
      out[f"c{case}_nb{j}"] = x
      out[f"c{case}_T{j}"] = T
    print(f"case {case}: centre {len(centre)} pts, groups {len(g)}, index {len(idx)}, pairs {len(mi)}")
  np.savez_compressed(os.path.join(HERE, "groups.npz"), **out)


if __name__ == "__main__":
  main()
