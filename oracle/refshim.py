"""ORACLE support (test infrastructure): import the reference's own Python in the BUILD container.

/root/reference does not exist on the GPU box, so this is only used (a) by tests marked to skip when the
reference is absent and (b) by tests/golden/make_golden.py to generate committed fixtures.

Shims: a `future_fstrings` codec alias (model/resunet.py:1 etc. declare that coding),
`sys.modules['MinkowskiEngine']` pointing at whichever ME-shaped module is under test, and empty stubs for
open3d / tensorboardX / easydict which the hot path never calls.
"""
import codecs
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
  return os.path.isdir(os.path.join(REFERENCE_ROOT, "model"))


def _codec_search(name):
  if name.replace("-", "_") in ("future_fstrings", "future_fstrings_"):
    return codecs.lookup("utf-8")
  return None


def install_me(me_module):
  """Register `me_module` (oracle.me_cpu or gcl_b200.MinkowskiEngine) as the importable `MinkowskiEngine`."""
  sys.modules["MinkowskiEngine"] = me_module
  sys.modules["MinkowskiEngine.MinkowskiFunctional"] = me_module.MinkowskiFunctional
  sys.modules["MinkowskiEngine.utils"] = me_module.utils


def import_reference(me_module, modules=("model.resunet",)):
  """Returns the list of imported reference modules, freshly bound to `me_module`."""
  assert reference_available(), "reference tree not present (expected only in the build container)"
  codecs.register(_codec_search)
  install_me(me_module)
  for stub in ("open3d", "tensorboardX", "easydict"):
    if stub not in sys.modules:
      m = types.ModuleType(stub)
      if stub == "tensorboardX":
        m.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
      if stub == "easydict":
        m.EasyDict = dict
      sys.modules[stub] = m
  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  # drop previously imported reference modules so they re-bind to the requested ME module
  for name in list(sys.modules):
    if name.split(".")[0] in ("model", "lib", "util") and getattr(sys.modules[name], "__file__", "") and \
        str(sys.modules[name].__file__).startswith(REFERENCE_ROOT):
      del sys.modules[name]
  return [importlib.import_module(m) for m in modules]
