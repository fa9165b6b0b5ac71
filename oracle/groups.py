"""ORACLE (test infrastructure only -- never imported by the product path): positive-group construction of the GCL
colocation loaders, restated from /root/reference/util/pointcloud.py:69-132 (`get_matching_indices_colocation`) and
:53-66 (`get_matching_indices`).

Parity status: the GROUP LOGIC (per-centre concatenation order, K truncation, finest-neighbour rule, skip rule, index
offsets) is pinned against the reference's own function, executed in the build container with an `open3d` stand-in whose
KDTreeFlann is this module's `radius_search` (tests/golden/make_golden_groups.py -> tests/golden/groups.npz).  The radius
search primitive itself is Open3D's (nanoflann) and Open3D is not installable here: PARITY UNPINNED for that primitive.  Its
published behaviour, restated: float64 points; a point is returned iff squared distance < radius^2 (nanoflann
RadiusResultSet::addPoint is strict); results sorted by ascending distance.  Ties in distance are implementation-defined in
nanoflann; this project defines smallest index first.
"""
import numpy as np


def radius_search(points64: np.ndarray, query64: np.ndarray, radius: float, K=None):
  """indices of `points64` within `radius` of `query64`, nearest first (ties: smallest index), at most K."""
  d2 = ((points64 - query64[None, :]) ** 2).sum(axis=1)
  idx = np.nonzero(d2 < radius * radius)[0]
  idx = idx[np.lexsort((idx, d2[idx]))]
  return idx[:K] if K is not None else idx


def transform_points(points64: np.ndarray, T: np.ndarray) -> np.ndarray:
  """open3d PointCloud.transform: x' = R x + t in float64"""
  return points64 @ T[:3, :3].T + T[:3, 3][None, :]


def colocation_groups(center_xyz, neighbourhood_xyz, list_trans, radius, K=None):
  """util/pointcloud.py:69-132.  center_xyz [Nc,3]; neighbourhood_xyz: list of [Nj,3] (sensor frame, float32 as in the
  loader: lib/colocation_data_loader.py:389); list_trans: list of 4x4 mapping cloud j into the centre frame.
  Returns (group int64 [G], index int64 [sum group], finest_flag bool [sum group])."""
  C = np.asarray(center_xyz, dtype=np.float64)
  nb32 = [np.asarray(x, dtype=np.float32) for x in neighbourhood_xyz]
  nbT = [transform_points(x.astype(np.float64), np.asarray(T, dtype=np.float64)) for x, T in zip(nb32, list_trans)]
  group, index, finest = [], [], []
  for i in range(len(C)):
    p = C[i]
    closest = np.linalg.norm(p)                                    # :97
    own = list(radius_search(C, p, radius, K))                     # :98-101
    n_own = len(own)
    lst = own
    finest_pos = 0                                                 # :106
    start = len(C)                                                 # :107
    for j in range(len(nbT)):
      idx = radius_search(nbT[j], p, radius, K)                    # :109-111
      if len(idx):
        dist = np.linalg.norm(nb32[j][idx[0]])                     # :113 (float32 norm of the un-transformed point)
        if dist < closest:                                         # :114-116
          closest = dist
          finest_pos = len(lst)
        lst = lst + [int(k) + start for k in idx]                  # :117
      start += len(nb32[j])                                        # :120
    if len(lst) == n_own:                                          # :121-122 no match in any neighbour cloud
      continue
    group.append(len(lst))
    index += [int(k) for k in lst]
    flags = [False] * len(lst)
    flags[finest_pos] = True
    finest += flags
  return np.asarray(group, np.int64), np.asarray(index, np.int64), np.asarray(finest, bool)


def matching_indices(source_xyz, target_xyz, trans, radius, K=None):
  """util/pointcloud.py:53-66: pairs (i, j) of source points (transformed) and their target neighbours, in loop order."""
  S = transform_points(np.asarray(source_xyz, np.float64), np.asarray(trans, np.float64))
  Tg = np.asarray(target_xyz, np.float64)
  out = []
  for i in range(len(S)):
    for j in radius_search(Tg, S[i], radius, K):
      out.append((i, int(j)))
  return np.asarray(out, np.int64).reshape(-1, 2)
