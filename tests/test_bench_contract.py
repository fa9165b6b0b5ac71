"""bench.py contract, CPU side: the reference arm (`--impl reference`, the oracle timed on the host cores) must run without a
GPU and print ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                     capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "scan_pairs_per_sec" and d["unit"] == "pairs/s"
  assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
  assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
  assert "workload" in d["config"] and "model" not in d["config"]
  cb = d["cpu_baseline"]
  assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
  e = d["e2e"]
  assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_non_zero_rank_of_the_reference_arm_exits_quietly():
  env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                      "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
  assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_training_workload_reference_arm_and_batch_generator():
  """BASELINE config 4: the oracle arm of `bench.py --workload train` runs on the CPU and prints the contract line; the synthetic
  colocated-scan batch (host side of gcl_b200.training) yields voxel-downsampled clouds with valid neighbour poses"""
  import numpy as np
  sys.path.insert(0, ROOT)
  from gcl_b200.training import synthetic_group_batch
  batch = synthetic_group_batch(rank=0, samples=1, voxel=0.3)
  assert len(batch) == 1
  clouds, Ts = batch[0]
  assert len(clouds) == 3 and len(Ts) == 2
  for c in clouds:
    assert c.dtype == np.float32 and c.shape[1] == 3 and len(c) > 1000
    key = np.floor(c / 0.3).astype(np.int64)
    assert len(np.unique(key, axis=0)) == len(c)                 # one point per voxel, like the loader's voxel_down_sample
  for T in Ts:
    assert T.shape == (4, 4) and abs(np.linalg.det(T[:3, :3]) - 1) < 1e-9
  r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "train", "--impl", "reference", "--steps", "1",
                      "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
  assert r.returncode == 0, r.stderr[-2000:]
  d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
  assert d["impl"] == "reference" and d["metric"] == "gcl_train_scans_per_sec" and d["unit"] == "scans/s" and d["value"] > 0
  assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_weak_scaling_batches_carry_equal_work_on_every_rank():
  """bench.make_batches: the ranks of a weak-scaling run get the same base scenes under rank-specific poses, so their voxel
  counts agree to a few % even on a 6-pair batch (0.6 % on the 16-pair batches of the bench; rank-specific scenes differed by +-8 %, which max-over-ranks timing reported as a scaling loss)"""
  import numpy as np
  import bench
  counts = []
  for rank in range(2):
    xyz, ptr = bench.make_batches(1, 6, seed=rank)[0]
    q = np.floor(xyz.numpy() / bench.VOXEL).astype(np.int64)
    cl = np.repeat(np.arange(len(ptr) - 1), np.diff(ptr.numpy()))
    key = ((cl * 4096 + q[:, 0] + 2048) * 4096 + q[:, 1] + 2048) * 4096 + q[:, 2] + 2048
    counts.append(len(np.unique(key)))
    if rank:
      assert not np.array_equal(xyz.numpy()[:1000], first[:1000])       # different poses: not the same data
    else:
      first = xyz.numpy().copy()
  assert max(counts) / min(counts) < 1.05, counts


def test_rank_pinning_gives_disjoint_core_slices():
  """bench.pin_rank_to_cores: one rank is left alone, the ranks of a multi-GPU run get disjoint slices of the allowed cores"""
  import bench
  if not hasattr(os, "sched_getaffinity"):
    return
  allowed = sorted(os.sched_getaffinity(0))
  try:
    assert bench.pin_rank_to_cores(0, 1) is None and sorted(os.sched_getaffinity(0)) == allowed
    if len(allowed) >= 4:
      a = bench.pin_rank_to_cores(0, 2)
      os.sched_setaffinity(0, allowed)
      b = bench.pin_rank_to_cores(1, 2)
      assert a and b and not set(a) & set(b) and set(a) | set(b) <= set(allowed)
      assert sorted(os.sched_getaffinity(0)) == sorted(b)
  finally:
    os.sched_setaffinity(0, allowed)

