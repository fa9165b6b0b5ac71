"""Build libgclb200.so in-tree with nvcc for sm_100a (no torch headers involved; seconds per file).

    python -m gcl_b200.build [--force] [-v]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgclb200.so")
OBJ = os.path.join(HERE, "csrc", "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def source_id() -> str:
  """short hash of every CUDA source + header of the library: identifies the build an ncu capture / profile belongs to"""
  import hashlib
  h = hashlib.sha1()
  files = sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
  inc = os.path.join(os.path.dirname(HERE), "include")
  for f in files + [os.path.join(inc, h) for h in sorted(os.listdir(inc)) if h.endswith(".h")]:
    with open(f if os.path.isabs(f) else os.path.join(CSRC, f), "rb") as fh:
      h.update(fh.read())
  return h.hexdigest()[:12]


def sources():
  return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _newer(target, deps):
  if not os.path.exists(target):
    return False
  t = os.path.getmtime(target)
  return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OBJ, exist_ok=True)
  headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
  inc = os.path.join(os.path.dirname(HERE), "include")
  headers += [os.path.join(inc, h) for h in os.listdir(inc) if h.endswith(".h")]
  jobs = []
  for s in sources():
    src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
    if force or not _newer(obj, [src] + headers):
      jobs.append((src, obj))

  def cc(job):
    src, obj = job
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, r

  with cf.ThreadPoolExecutor(max_workers=8) as ex:
    for src, r in ex.map(cc, jobs):
      if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
      if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}")
  objs = [os.path.join(OBJ, s[:-3] + ".o") for s in sources()]
  if force or jobs or not _newer(OUT, objs):
    cmd = [NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
      sys.stderr.write(r.stdout + r.stderr)
      raise RuntimeError("link failed")
  return OUT


if __name__ == "__main__":
  print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
