"""CPU, build container only (skipped where /root/reference is absent): the reference's own modules imported over
the oracle -- checkpoint contract (state_dict keys/shapes) and graph equivalence of gcl_b200.resunet."""
import numpy as np
import pytest
import torch

import oracle.me_cpu as OME
from oracle import matching as omatch
from oracle.refshim import import_reference, reference_available
from gcl_b200.resunet import make_models, VARIANTS
from helpers import numpy_seeded_weights, small_cloud

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_state_dict_contract(name):
  (ref_mod,) = import_reference(OME, ("model.resunet",))
  ref = getattr(ref_mod, name)(1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  mine = make_models(OME)[name](1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  a, b = ref.state_dict(), mine.state_dict()
  assert list(a.keys()) == list(b.keys())
  assert all(a[k].shape == b[k].shape for k in a)
  mine.load_state_dict(a)   # strict


def test_same_graph_same_output():
  (ref_mod,) = import_reference(OME, ("model.resunet",))
  ref = ref_mod.ResUNetBN2C(1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  numpy_seeded_weights(ref, 3).eval()
  mine = make_models(OME)["ResUNetBN2C"](1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  mine.load_state_dict(ref.state_dict())
  mine.eval()
  t = torch.from_numpy(small_cloud(4, 1500, 5.0))
  _, sel = OME.utils.sparse_quantize(t / 0.3, return_index=True)
  C, F = OME.utils.sparse_collate([torch.floor(t[sel] / 0.3).int()], [torch.ones(len(sel), 1)])
  with torch.no_grad():
    a = ref(OME.SparseTensor(F, coordinates=C)).F
    b = mine(OME.SparseTensor(F, coordinates=C)).F
  assert torch.equal(a, b)


def test_gcl_b200_me_module_satisfies_reference_imports():
  """the product's ME-shaped module lets the reference model file import and construct (no GPU needed for that)"""
  from gcl_b200 import MinkowskiEngine as GME
  (ref_mod,) = import_reference(GME, ("model.resunet",))
  m = ref_mod.ResUNetBN2C(1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)
  assert sum(p.numel() for p in m.parameters()) == 8753408
  assert m.conv1.kernel.shape == (125, 1, 32) and m.conv1_tr.kernel.shape == (96, 64) and m.final.bias.shape == (1, 32)
  import_reference(OME, ("model.resunet",))   # restore


def test_reference_find_nn_equals_oracle():
  (ev,) = import_reference(OME, ("lib.eval",))
  g = torch.Generator().manual_seed(1)
  A, B = torch.randn(333, 32, generator=g), torch.randn(517, 32, generator=g)
  i0, d0 = ev.find_nn_gpu(A, B, nn_max_n=100, return_distance=True)
  i1, d1 = omatch.find_nn(A, B, nn_max_n=100, return_distance=True)
  assert torch.equal(i0, i1) and torch.equal(d0, d1)
