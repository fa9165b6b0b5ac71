"""Table-driven sparse residual U-Net (FCGF family) over any MinkowskiEngine-shaped operator module.

The reference's model file (/root/reference/model/resunet.py:11-232, residual_block.py:9-53, common.py:4-10)
runs unmodified on `gcl_b200.MinkowskiEngine`; this module exists because the reference tree is not present
on the GPU box, and because the fused inference engine (gcl_b200/engine.py) needs the layer table.  It
produces the *same module names and state_dict keys/shapes* as the reference classes (checkpoint contract,
SURVEY.md Appendix A10), verified in tests/test_reference_live.py (build container) and the golden tests/golden/resunet_bn2c.npz.

`make_models(ME)` returns {'ResUNetBN2C': cls, ...} built on the given operator module, so the same graph
is instantiated over the CUDA operators (product) or over the CPU oracle (tests).
"""
from __future__ import annotations

import torch
import torch.nn as nn

# name -> (CHANNELS, TR_CHANNELS); NORM 'BN' everywhere (model/resunet.py:236-266)
VARIANTS = {
    "ResUNetBN2":     ([None, 32, 64, 128, 256], [None, 32, 64, 64, 128]),
    "ResUNetBN2B":    ([None, 32, 64, 128, 256], [None, 64, 64, 64, 64]),
    "ResUNetBN2C":    ([None, 32, 64, 128, 256], [None, 64, 64, 64, 128]),
    "ResUNetBN2D":    ([None, 32, 64, 128, 256], [None, 64, 64, 128, 128]),
    "ResUNetBN2E":    ([None, 128, 128, 128, 256], [None, 64, 128, 128, 128]),
    "ResUNetFatBN":   ([None, 32, 64, 128, 256], [None, 128, 128, 128, 256]),
}


def make_models(ME):
  MEF = ME.MinkowskiFunctional
  # norm followed by ReLU: one fused launch pair on gcl_b200.MinkowskiEngine, MEF.relu(norm(x)) on any other operator module
  bn_relu = getattr(ME, "bn_relu", None) or (lambda norm, x: MEF.relu(norm(x)))

  class Block(nn.Module):
    """conv3-BN-ReLU-conv3-BN-(+x)-ReLU (residual_block.py:40-53)."""

    def __init__(self, planes, bn_momentum, D):
      super().__init__()
      self.conv1 = ME.MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dimension=D)
      self.norm1 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)
      self.conv2 = ME.MinkowskiConvolution(planes, planes, kernel_size=3, stride=1, dilation=1, bias=False,
                                           dimension=D)
      self.norm2 = ME.MinkowskiBatchNorm(planes, momentum=bn_momentum)

    def forward(self, x):
      y = bn_relu(self.norm1, self.conv1(x))
      y = self.norm2(self.conv2(y))
      y += x
      return MEF.relu(y)

  class ResUNet(ME.MinkowskiNetwork):
    CHANNELS = TR_CHANNELS = None

    def __init__(self, in_channels=3, out_channels=32, bn_momentum=0.1, normalize_feature=None,
                 conv1_kernel_size=None, D=3):
      ME.MinkowskiNetwork.__init__(self, D)
      C, T = self.CHANNELS, self.TR_CHANNELS
      self.normalize_feature = normalize_feature
      mk = dict(bias=False, dimension=D)
      bn = lambda c: ME.MinkowskiBatchNorm(c, momentum=bn_momentum)
      # encoder: level l in 1..4; registration order matches the reference so state_dict order is identical
      self.conv1 = ME.MinkowskiConvolution(in_channels, C[1], kernel_size=conv1_kernel_size, stride=1,
                                           dilation=1, **mk)
      self.norm1 = bn(C[1])
      self.block1 = Block(C[1], bn_momentum, D)
      for l in (2, 3, 4):
        setattr(self, f"conv{l}", ME.MinkowskiConvolution(C[l - 1], C[l], kernel_size=3, stride=2, dilation=1, **mk))
        setattr(self, f"norm{l}", bn(C[l]))
        setattr(self, f"block{l}", Block(C[l], bn_momentum, D))
      # decoder: level 4 -> 2, input = encoder skip (+) previous decoder output
      for l in (4, 3, 2):
        cin = C[l] if l == 4 else C[l] + T[l + 1]
        setattr(self, f"conv{l}_tr", ME.MinkowskiConvolutionTranspose(cin, T[l], kernel_size=3, stride=2,
                                                                     dilation=1, **mk))
        setattr(self, f"norm{l}_tr", bn(T[l]))
        setattr(self, f"block{l}_tr", Block(T[l], bn_momentum, D))
      self.conv1_tr = ME.MinkowskiConvolution(C[1] + T[2], T[1], kernel_size=1, stride=1, dilation=1, **mk)
      self.final = ME.MinkowskiConvolution(T[1], out_channels, kernel_size=1, stride=1, dilation=1, bias=True,
                                           dimension=D)

    def forward(self, x):
      skips = {}
      y = self.block1(self.norm1(self.conv1(x)))
      skips[1] = y
      y = MEF.relu(y)
      for l in (2, 3, 4):
        y = getattr(self, f"block{l}")(getattr(self, f"norm{l}")(getattr(self, f"conv{l}")(y)))
        skips[l] = y
        y = MEF.relu(y)
      for l in (4, 3, 2):
        y = getattr(self, f"block{l}_tr")(getattr(self, f"norm{l}_tr")(getattr(self, f"conv{l}_tr")(y)))
        y = ME.cat(MEF.relu(y), skips[l - 1])
      y = self.final(MEF.relu(self.conv1_tr(y)))
      if self.normalize_feature:
        return ME.SparseTensor(y.F / torch.norm(y.F, p=2, dim=1, keepdim=True),
                               coordinate_map_key=y.coordinate_map_key, coordinate_manager=y.coordinate_manager)
      return y

  out = {}
  for name, (ch, tr) in VARIANTS.items():
    out[name] = type(name, (ResUNet,), {"CHANNELS": ch, "TR_CHANNELS": tr})
  return out
