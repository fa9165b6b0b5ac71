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
