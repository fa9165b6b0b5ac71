"""MinkowskiEngine.MinkowskiFunctional subset (MEF.relu: model/resunet.py:177-223, residual_block.py:42,51)."""
import torch

from .. import ops


def relu(x, inplace=False):
  f = x.F
  if torch.is_grad_enabled() and f.requires_grad:
    return x._like(torch.relu(f))
  return x._like(ops.affine_act(f.detach(), relu=True))
