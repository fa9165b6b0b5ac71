"""CPU analysis for the round-2 conv redesign: how much operand reuse does a 128-row output tile offer?
For one synthetic KITTI-shape scan at 0.3 m (stride-1 3x3x3 map) and several row orders, per tile:
  stages  = kernel offsets populated anywhere in the tile (what the current kernel pays),
  gathers = valid (row, offset) entries (lines fetched from L2 today),
  unique  = distinct input rows the tile touches (lines a spatially tiled kernel would fetch once).
Pure numpy on the oracle's tables; no GPU."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle.me_cpu as OME  # noqa: E402
from gcl_b200 import synth  # noqa: E402

xyz = torch.from_numpy(synth.cast(synth.Scene(11), synth.KITTI, seed=0))
_, sel = OME.utils.sparse_quantize(xyz / 0.3, return_index=True)
C = torch.floor(xyz[sel] / 0.3).int().numpy()
C4 = np.concatenate([np.zeros((len(C), 1), np.int32), C], 1)
nbr = OME.build_neighbor_table(C4, C4, OME.kernel_offsets(3, 1))
n = len(C)
valid = nbr >= 0
k = np.arange(27)
ix, iy, iz = k % 3, (k // 3) % 3, k // 9
key = sum((valid[:, s].any(1).astype(np.int64) << b) for b, s in enumerate([ix == 0, ix == 2, iy == 0, iy == 2, iz == 0, iz == 2]))


def morton(c):
  c = (c - c.min(0)).astype(np.uint64)
  out = np.zeros(len(c), np.uint64)
  for b in range(10):
    for a in range(3):
      out |= ((c[:, a] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + a)
  return out


m = morton(C)
orders = {"scan order (first occurrence)": np.arange(n),
          "direction-key buckets, stable (current)": np.argsort(key, kind="stable"),
          "Morton": np.argsort(m, kind="stable"),
          "direction key, Morton inside a bucket": np.lexsort((m, key)),
          "Morton blocks of 2048 rows, direction key inside": None}
mo = np.argsort(m, kind="stable")
blk = np.empty(n, np.int64); blk[mo] = np.arange(n) // 2048
orders["Morton blocks of 2048 rows, direction key inside"] = np.lexsort((key, blk))

res = {}
for name, perm in orders.items():
  st = ga = un = 0
  tiles = 0
  for t in range(0, n, 128):
    rows = perm[t:t + 128]
    v = valid[rows]
    st += int(v.any(0).sum())
    ga += int(v.sum())
    un += len(np.unique(nbr[rows][v]))
    tiles += 1
  res[name] = {"stages_per_tile": round(st / tiles, 2), "gathered_lines_per_tile": round(ga / tiles, 1),
               "unique_rows_per_tile": round(un / tiles, 1), "reuse_factor": round(ga / un, 2)}
print(json.dumps({"voxels": n, "pairs_per_voxel": round(valid.sum() / n, 2), "orders": res}, indent=1))
