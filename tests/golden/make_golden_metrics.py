"""Golden vectors for FCGF's hardest-contrastive loss from the REFERENCE's own lib/trainer.py (build container only):
HardestContrastiveLossTrainer.contrastive_hardest_negative_loss run unbound on CPU tensors with np.random.seed(7).

    python tests/golden/make_golden_metrics.py   ->  tests/golden/hardest_loss.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle.me_cpu as OME  # noqa: E402
from oracle.refshim import import_reference  # noqa: E402


def main():
  (tr,) = import_reference(OME, ("lib.trainer",))
  rng = np.random.RandomState(21)
  N0, N1, C, P = 3000, 2800, 32, 1500
  F0 = rng.randn(N0, C).astype(np.float32); F0 /= np.linalg.norm(F0, axis=1, keepdims=True)
  F1 = rng.randn(N1, C).astype(np.float32); F1 /= np.linalg.norm(F1, axis=1, keepdims=True)
  pairs = np.stack([rng.choice(N0, P, replace=False), rng.choice(N1, P, replace=False)], 1).astype(np.int64)
  F1[pairs[:, 1]] = F0[pairs[:, 0]] + 0.15 * rng.randn(P, C).astype(np.float32)
  F1 /= np.linalg.norm(F1, axis=1, keepdims=True)
  self = types.SimpleNamespace(pos_thresh=0.1, neg_thresh=1.4)
  fn = tr.HardestContrastiveLossTrainer.contrastive_hardest_negative_loss
  a, b = torch.from_numpy(F0).clone().requires_grad_(True), torch.from_numpy(F1).clone().requires_grad_(True)
  np.random.seed(7)
  pos, neg = fn(self, a, b, torch.from_numpy(pairs), num_pos=1024, num_hn_samples=512)
  (pos + neg).backward()
  np.random.seed(7)        # replay the reference's draws so that the fixture carries the selections
  sel0 = np.random.choice(N0, 512, replace=False); sel1 = np.random.choice(N1, 512, replace=False)
  pos_sel = np.random.choice(P, 1024, replace=False)
  np.savez_compressed(os.path.join(HERE, "hardest_loss.npz"), F0=F0, F1=F1, pairs=pairs, sel0=sel0, sel1=sel1, pos_sel=pos_sel,
                      losses=np.array([float(pos), float(neg)]), g0=a.grad.numpy(), g1=b.grad.numpy())
  print("hardest loss:", float(pos), float(neg))


if __name__ == "__main__":
  main()
