"""GCL training step on the CUDA operators (BASELINE config 4; lib/colocation_trainer.py:811-916).

One step = ResUNetBN2C forward in train mode on a batch of `samples` colocated scan groups (1 centre + 2 neighbour scans each:
3 clouds per sample), the fused finest-contrastive loss (positive groups + hardest negatives, lib/colocation_trainer.py:430-535),
backward through the sparse convolutions / BatchNorm kernels, the data-parallel gradient exchange (ONE NCCL all-reduce of the
flat 35 MB gradient buffer, gcl_b200/sharding.py) and SGD(momentum .8, wd 1e-4, lr .1) (lib/colocation_trainer.py:73-77).
Sharding: each rank owns its own colocated scan groups (a sample's scans must stay together: positives are intra-sample and
hardest negatives are mined inside the rank's batch); BatchNorm statistics stay rank-local like torch DDP without SyncBatchNorm.

Synthetic input (there are no datasets offline): the three scans of a sample are ray-casts of one scene from poses a few metres
apart; positive groups come from the GPU voxel-hash radius search (gcl_b200.groups = util/pointcloud.py:69-132).
"""
from __future__ import annotations

import numpy as np
import torch

from . import MinkowskiEngine as ME
from . import groups as gg
from . import ops, synth
from .loss import GroupContrastiveLoss
from .sharding import PackedGradients


def synthetic_group_batch(rank: int, samples: int, voxel: float = 0.3, sensor=None, seed: int = 0):
  """host side of a training batch: per sample 3 voxel-downsampled clouds (sensor frame) + their poses into the centre frame.
  Returns list of (clouds [3 x float32 [n,3]], transforms [2 x 4x4 neighbour -> centre])."""
  sensor = sensor or synth.NUSCENES
  out = []
  for s in range(samples):
    scene = synth.Scene(40 + s + 1000 * seed)     # the same scenes on every rank (weak scaling = equal work); the scans differ by rank below
    poses = [(0.0, 0.0, 0.0), (4.0, 0.5, 0.05), (-3.0, -0.5, -0.04)]
    clouds, Ts = [], []
    for j, (x, y, yaw) in enumerate(poses):
      pts = synth.cast(scene, sensor, (x, y, yaw), seed=7 * s + j + 100 * rank)
      # one point per voxel, like the loader's voxel_down_sample (lib/colocation_data_loader.py:379,388)
      key = np.floor(pts / voxel).astype(np.int64)
      _, first = np.unique(key, axis=0, return_index=True)
      clouds.append(np.ascontiguousarray(pts[np.sort(first)], dtype=np.float32))
      if j:
        c, sn = np.cos(yaw), np.sin(yaw)
        T = np.eye(4)
        T[:3, :3] = [[c, -sn, 0], [sn, c, 0], [0, 0, 1]]
        T[:3, 3] = [x, y, 0.0]
        Ts.append(T)
    out.append((clouds, Ts))
  return out


class GclTrainStep:
  """builds the batch (device-resident), model, loss and optimiser once; `step()` runs one optimisation step and returns the loss"""

  def __init__(self, device, rank: int = 0, samples: int = 4, voxel: float = 0.3, conv: str = "tf32", seed: int = 0,
               model_name: str = "ResUNetBN2C", host_batch=None):
    import gcl_b200
    self.dev, self.voxel, self.samples = device, voxel, samples
    ME.set_training_conv_algo(conv)
    self.host_batch = host_batch if host_batch is not None else synthetic_group_batch(rank, samples, voxel, seed=seed)
    self._upload_and_group(self.host_batch)
    torch.manual_seed(0)            # identical initial weights on every rank
    self.model = gcl_b200.load_model(model_name)(1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True).to(device)
    self.model.train()
    self.opt = torch.optim.SGD(self.model.parameters(), lr=0.1, momentum=0.8, weight_decay=1e-4)
    self.crit = GroupContrastiveLoss(pos_thresh=0.1, neg_thresh=1.4, finest_thresh=0.2, square_loss=True,
                                     rng=np.random.default_rng(0))
    self.grads = PackedGradients(self.model.parameters())

  def _upload_and_group(self, host_batch, pinned=None):
    """H2D of the clouds, K1 voxelisation of the whole batch, positive groups per sample on the GPU, pair hashes"""
    dev = self.dev
    clouds = [c for cl, _ in host_batch for c in cl]
    lens = [len(c) for c in clouds]
    if pinned is None:
      xyz = torch.from_numpy(np.concatenate(clouds)).to(dev)
    else:
      xyz = pinned.to(dev, non_blocking=True)
    ptr = torch.tensor(np.cumsum([0] + lens), dtype=torch.int64)
    cm, umap = ops.voxelize(xyz, self.voxel, ptr)
    assert cm.n == xyz.shape[0], "training clouds are voxel-downsampled: one point per voxel"
    self.cm, self.n_rows, self.n_clouds = cm, cm.n, len(clouds)
    off = np.cumsum([0] + lens)
    group_l, index_l, finest_l = [], [], []
    for s, (cl, Ts) in enumerate(host_batch):
      base = int(off[3 * s])
      centre = xyz[off[3 * s]:off[3 * s + 1]]
      nbs = [xyz[off[3 * s + j]:off[3 * s + j + 1]] for j in (1, 2)]
      grp, idx, fin = gg.colocation_groups(centre, nbs, Ts, self.voxel, 1.5 * self.voxel, 5)
      group_l.append(grp); index_l.append(idx + base); finest_l.append(fin)
    self.group = torch.cat(group_l)
    self.index = torch.cat(index_l)
    self.finest = torch.cat(finest_l)
    self.index_hash = gg.exhaustive_hash(self.group, self.index, self.n_rows)
    self.feats = torch.ones(self.n_rows, 1, device=dev)
    self.h2d_bytes = int(xyz.numel() * 4)
    self.prepared = None     # per-batch loss inputs (CSR, finest positions, sorted pair keys): built on the device, once

  def step(self):
    self.grads.zero()
    st = ME.SparseTensor(self.feats, coordinates=self.cm.coords)
    F = self.model(st).F
    if self.prepared is None:
      self.prepared = self.crit.prepare(self.group, self.index, self.index_hash, self.finest, self.dev)
    pos, fin, neg = self.crit.finest_contrastive_loss(F, None, None, None, None, max_pos_cluster=256 * self.samples,
                                                      max_hn_samples=256 * self.samples, prepared=self.prepared)
    loss = pos + fin + neg
    loss.backward()
    self.grads.allreduce()            # the path's only exchange step (NCCL over NVLink when world > 1)
    self.opt.step()
    return loss
