"""Golden fixture of the circle-loss head from the REFERENCE's own Python (build container only):

    python tests/golden/make_golden_circle.py

  circle_loss.npz   /root/reference/lib/colocation_trainer.py:538-681 `location_circle_loss` (unbound, SimpleNamespace self with the
                    trainer attributes of :410-420) on a seeded 2-item batch: forward values and dL/dF for five settings
                    (square / L2 distance x blocked / open finest gradient, and the pair-positive variant)
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
from oracle.refshim import import_reference  # noqa: E402

VARIANTS = {   # name: (square_loss, block_finest_gradient, use_pair_group_positive_loss)
    "sq_block": (True, True, False), "sq_open": (True, False, False), "l2_block": (False, True, False),
    "l2_open": (False, False, False), "sq_pair": (True, True, True),
}


def inputs():
  rng = np.random.RandomState(23)
  lens = [800, 700]                              # two batch items
  N, C = sum(lens), 32
  F = rng.randn(N, C).astype(np.float32)
  pts = (rng.rand(N, 3) * 6).astype(np.float32)
  sizes, index = [], []
  for b, (s0, n) in enumerate(zip([0, lens[0]], lens)):
    for _ in range(140):                           # 280 groups > max_pos_cluster = 256: the selection branch is exercised
      k = rng.randint(2, 7)
      sizes.append(k)
      index.append(s0 + rng.choice(n, k, replace=False))
  sizes = np.array(sizes, np.int64)
  index = np.concatenate(index).astype(np.int64)
  starts = np.concatenate([[0], np.cumsum(sizes)])
  flag = np.zeros(len(index), bool)
  flag[starts[:-1] + np.array([rng.randint(0, s) for s in sizes])] = True
  for g in range(len(sizes)):
    m = index[starts[g]:starts[g + 1]]
    F[m] = F[m[0]] + 0.25 * rng.rand() * rng.randn(len(m), C).astype(np.float32)
    pts[m] = pts[m[0]] + 0.02 * rng.randn(len(m), 3).astype(np.float32)
  F /= np.linalg.norm(F, axis=1, keepdims=True)
  return dict(F=F, points=pts, group=sizes, index=index, finest_flag=flag, batch_lengths=np.array(lens, np.int64))


def main():
  (ct,) = import_reference(OME, ("lib.colocation_trainer",))
  out = inputs()
  for name, (square, block, pair) in VARIANTS.items():
    self = types.SimpleNamespace(device=torch.device("cpu"), pos_thresh=0.1, neg_thresh=1.4, finest_thresh=0.2, square_loss=square,
                                 block_finest_gradient=block, use_pair_group_positive_loss=pair, log_scale=16, safe_radius=0.75)
    Ft = torch.from_numpy(out["F"]).clone().requires_grad_(True)
    np.random.seed(5)
    pos, fin, neg = ct.FinestContrastiveLossTrainer.location_circle_loss(
        self, Ft, torch.from_numpy(out["group"]), torch.from_numpy(out["index"]), None, torch.from_numpy(out["finest_flag"]),
        max_pos_cluster=256, max_hn_samples=None, points=torch.from_numpy(out["points"]), batch_lengths=out["batch_lengths"].tolist())
    (1.0 * pos + 0.5 * fin + 2.0 * neg).backward()
    out[name + "_losses"] = np.array([float(pos), float(fin), float(neg)], np.float64)
    out[name + "_grad"] = Ft.grad.numpy().astype(np.float32)
    print(name, out[name + "_losses"])
  np.savez_compressed(os.path.join(HERE, "circle_loss.npz"), **out)


if __name__ == "__main__":
  main()
