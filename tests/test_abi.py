"""The C-ABI library loads on a CPU-only box and exports every symbol include/gclb200.h declares; the ctypes
table in gcl_b200/_lib.py covers exactly that set.  No compute calls here (there is no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", f) for f in sorted(os.listdir(os.path.join(ROOT, "include"))) if f.endswith(".h")]
HEADER = os.path.join(ROOT, "include", "gclb200.h")


def _declared():
  names = set()
  for h in HEADERS:
    src = open(h).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names |= set(re.findall(r"\b(gclb_[a-z0-9_]+)\s*\(", src))
  return sorted(names)


@pytest.fixture(scope="module")
def lib():
  from gcl_b200 import _lib, build
  build.build()
  return _lib.load()


def test_every_declared_symbol_is_exported(lib):
  names = _declared()
  assert len(names) >= 20
  for n in names:
    assert hasattr(lib, n), f"{n} declared in gclb200.h but not exported"


def test_ctypes_table_matches_header():
  from gcl_b200 import _lib
  assert sorted(_lib.SIGNATURES) == _declared()


def test_host_side_helpers(lib):
  assert lib.gclb_version() >= 100
  assert lib.gclb_hash_capacity(0) == 1024
  assert lib.gclb_hash_capacity(1000) == 2048            # quad slots: 2 per row
  assert lib.gclb_hash_capacity(130000) == 262144
  assert lib.gclb_hash_bytes(1024) == 1024 * 32
  assert lib.gclb_compact_workspace_bytes(5000) >= 5000 * 4
  assert lib.gclb_nn_workspace_bytes(100, 200, 1, 100, 200) >= 300 * 8


def test_argument_errors_are_reported_not_crashed(lib):
  rc = lib.gclb_hash_build(None, 1024, 1, None, 10, None, None)
  assert rc == -1 and b"null" in lib.gclb_last_error()
  rc = lib.gclb_spconv_fwd(None, 0, None, 0, 0, None, 1, 1, None, None, None, None, None, None, 0, None, 0, 0, None)
  assert rc == -1


def test_product_refuses_cpu_tensors():
  import torch
  from gcl_b200 import MinkowskiEngine as ME
  from gcl_b200._lib import GclbError
  with pytest.raises(GclbError):
    ME.SparseTensor(torch.ones(2, 1), coordinates=torch.zeros(2, 4, dtype=torch.int32))
  from gcl_b200 import ops
  with pytest.raises(GclbError):
    ops.voxelize(torch.zeros(4, 3), 0.3)


def test_product_never_imports_oracle():
  bad = []
  for dirpath, _, files in os.walk(os.path.join(ROOT, "gcl_b200")):
    for f in files:
      if f.endswith(".py"):
        s = open(os.path.join(dirpath, f)).read()
        if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M):
          bad.append(f)
  assert not bad, f"product modules import the oracle: {bad}"
