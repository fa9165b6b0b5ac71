"""Regenerates the committed golden fixtures from the REFERENCE's own Python (build container only).

    python tests/golden/make_golden.py

  resunet_bn2c.npz   /root/reference/model/resunet.py ResUNetBN2C (unmodified) run over the CPU oracle operators on a
                     seeded 2-cloud input with numpy-seeded weights -> output descriptors
  nn.npz             /root/reference/lib/eval.py find_nn_gpu (+ lib/metrics.py pdist) on CPU tensors, and the mutual-NN
                     pairs by the reference's algorithm (generalization_ETH/evaluate.py:63-77: two sklearn KD-trees)
  gcl_loss.npz       /root/reference/lib/colocation_trainer.py finest_contrastive_loss / location_contrastive_loss
                     (unbound, SimpleNamespace self) forward values and dL/dF
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle.me_cpu as OME  # noqa: E402
from oracle import gcl_loss as oloss  # noqa: E402
from oracle.refshim import import_reference  # noqa: E402
from helpers import numpy_seeded_weights, small_cloud  # noqa: E402


def resunet_fixture():
  (resunet,) = import_reference(OME, ("model.resunet",))
  clouds = [small_cloud(1, 2500, 7.0), small_cloud(2, 1500, 5.0)]
  cs = []
  for x in clouds:
    t = torch.from_numpy(x)
    _, sel = OME.utils.sparse_quantize(t / 0.3, return_index=True)
    cs.append(torch.floor(t[sel] / 0.3).int())
  C, F = OME.utils.sparse_collate(cs, [torch.ones(len(c), 1) for c in cs])
  model = resunet.ResUNetBN2C(1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  numpy_seeded_weights(model, seed=7).eval()
  with torch.no_grad():
    out = model(OME.SparseTensor(F, coordinates=C)).F
  np.savez_compressed(os.path.join(HERE, "resunet_bn2c.npz"), coords=C.numpy(), feats_out=out.numpy().astype(np.float32),
                      weight_seed=7, voxel=0.3)
  print("resunet_bn2c:", tuple(out.shape))


def nn_fixture():
  ev, = import_reference(OME, ("lib.eval",))
  rng = np.random.RandomState(3)
  F0 = rng.randn(700, 32).astype(np.float32); F0 /= np.linalg.norm(F0, axis=1, keepdims=True)
  F1 = rng.randn(900, 32).astype(np.float32); F1 /= np.linalg.norm(F1, axis=1, keepdims=True)
  F1[:300] = F0[:300] + 0.02 * rng.randn(300, 32).astype(np.float32)
  idx, d = ev.find_nn_gpu(torch.from_numpy(F0), torch.from_numpy(F1), nn_max_n=250, return_distance=True)
  idx_l2, d_l2 = ev.find_nn_gpu(torch.from_numpy(F0), torch.from_numpy(F1), nn_max_n=-1, return_distance=True, dist_type="L2")
  from sklearn.neighbors import KDTree
  nn01 = KDTree(F1).query(F0, 1)[1][:, 0]
  nn10 = KDTree(F0).query(F1, 1)[1][:, 0]
  pairs = np.array([[i, nn01[i]] for i in range(len(F0)) if nn10[nn01[i]] == i], dtype=np.int64)
  np.savez_compressed(os.path.join(HERE, "nn.npz"), F0=F0, F1=F1, idx=idx.numpy(), dist=d.numpy(), idx_l2=idx_l2.numpy(),
                      dist_l2=d_l2.numpy(), mutual_pairs=pairs)
  print("nn:", idx.shape, pairs.shape)


def loss_fixture():
  ct, misc = import_reference(OME, ("lib.colocation_trainer", "util.misc"))
  rng = np.random.RandomState(11)
  N, G, C = 3000, 500, 32
  F = rng.randn(N, C).astype(np.float32)
  sizes = rng.randint(2, 7, G)
  index = np.concatenate([rng.choice(N, s, replace=False) for s in sizes]).astype(np.int64)
  starts = np.concatenate([[0], np.cumsum(sizes)])
  flag = np.zeros(len(index), bool)
  flag[starts[:-1] + np.array([rng.randint(0, s) for s in sizes])] = True
  for g in range(G):
    m = index[starts[g]:starts[g + 1]]
    F[m] = F[m[0]] + 0.3 * rng.rand() * rng.randn(len(m), C).astype(np.float32)
  F /= np.linalg.norm(F, axis=1, keepdims=True)
  # the reference's own util/misc.py:29-36 (what the collate calls), not the oracle restatement
  ih = misc._exhaustive_hash(torch.split(torch.from_numpy(index), sizes.tolist()), N).astype(np.int64)
  assert np.array_equal(ih, oloss.exhaustive_hash([index[starts[g]:starts[g + 1]] for g in range(G)], N))
  out = dict(F=F, group=sizes.astype(np.int64), index=index, finest_flag=flag, index_hash=ih)
  for name, square in (("finest_sq", True), ("finest_l2", False), ("location", False)):
    self = types.SimpleNamespace(device=torch.device("cpu"), pos_thresh=0.1, neg_thresh=1.4, finest_thresh=0.2,
                                 square_loss=square, block_finest_gradient=False, use_pair_group_positive_loss=False,
                                 use_hard_negative=True)
    fn = ct.FinestContrastiveLossTrainer.location_contrastive_loss if name == "location" else \
        ct.FinestContrastiveLossTrainer.finest_contrastive_loss
    Ft = torch.from_numpy(F).clone().requires_grad_(True)
    np.random.seed(5)
    pos, fin, neg = fn(self, Ft, torch.from_numpy(sizes), torch.from_numpy(index), ih, torch.from_numpy(flag),
                       max_pos_cluster=256, max_hn_samples=512)
    (1.0 * pos + 0.5 * fin + 2.0 * neg).backward()
    out[name + "_losses"] = np.array([float(pos), float(fin), float(neg)], np.float64)
    out[name + "_grad"] = Ft.grad.numpy().astype(np.float32)
    print(name, out[name + "_losses"])
  np.savez_compressed(os.path.join(HERE, "gcl_loss.npz"), **out)


if __name__ == "__main__":
  resunet_fixture()
  nn_fixture()
  loss_fixture()
