"""gcl_b200 -- B200-native (sm_100a) FCGF/GCL feature-extraction + matching hot path.

Public surface (mirrors what liuQuan98/GCL's hot path consumes):
  gcl_b200.MinkowskiEngine         MinkowskiEngine-shaped operator module (drop-in for model/resunet.py etc.)
  gcl_b200.install_as_minkowski_engine()   make `import MinkowskiEngine` resolve to it
  gcl_b200.matching                find_nn_gpu / pdist / mutual_nn (calculate_M) / find_corr
  gcl_b200.engine.ResUNetEngine    fused eval-mode ResUNet forward (BN/ReLU/residual/cat folded into the convs)
  gcl_b200.ops                     tensor-level wrappers of the C ABI (include/gclb200.h)
There is no CPU fallback anywhere in this package: without libgclb200.so + a CUDA device, calls raise.
"""
import sys

__version__ = "0.1.0"


def install_as_minkowski_engine():
  """Register gcl_b200.MinkowskiEngine under the name `MinkowskiEngine` (and its submodules)."""
  from . import MinkowskiEngine as ME
  sys.modules["MinkowskiEngine"] = ME
  sys.modules["MinkowskiEngine.MinkowskiFunctional"] = ME.MinkowskiFunctional
  sys.modules["MinkowskiEngine.utils"] = ME.utils
  return ME


def load_model(name: str):
  """name -> model class on the CUDA operators (model/__init__.py:20 `load_model`)."""
  from . import MinkowskiEngine as ME
  from .resunet import make_models
  models = make_models(ME)
  if name not in models:
    raise ValueError(f"Invalid model index. Options are {list(models)}")
  return models[name]
