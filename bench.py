#!/usr/bin/env python
"""bench.py -- scan-pairs/sec for the GCL/FCGF inference hot path on B200 (BASELINE.json metric).

Workload = BASELINE.json configs[1]: synthetic LoKITTI-style distant scan pairs (64-beam, ~130k points per scan,
0.3 m voxels): per pair  voxelise+hash (K1) -> strided + kernel maps (K2) -> ResUNetBN2C(k5, 32-d) forward of BOTH
scans with BatchNorm/ReLU/residual/cat fused (K3) -> 5000-point subsample -> mutual nearest neighbours (K4).
A "step" = one batch of --pairs pairs through that path.  One process per GPU, pairs sharded by rank, no collective.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl reference]

Prints ONE JSON line (see the contract in the task statement): `value` = whole-job pairs/s with inputs resident in
HBM; `e2e` = the same through the public API from pinned HOST buffers incl. H2D of the points and D2H of the
correspondences; `roofline` for the dominant kernel (sparse-conv forward); `cpu_baseline` = the CPU oracle
restatement of the reference's MinkowskiEngine path on this host's cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOXEL = 0.3
SUBSAMPLE = 5000
MODEL = dict(in_channels=1, out_channels=32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True)


def seeded_model(ME, seed=0):
  """ResUNetBN2C with seeded random weights and non-trivial BN statistics (no checkpoint is available offline)."""
  from gcl_b200.resunet import make_models
  torch.manual_seed(seed)
  m = make_models(ME)["ResUNetBN2C"](**MODEL)
  g = torch.Generator().manual_seed(seed + 1)
  for mod in m.modules():
    if isinstance(mod, torch.nn.BatchNorm1d):
      mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
      mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
      mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
      mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
  return m.eval()


SENSOR = {"name": "KITTI"}      # set by --workload nuscenes


def make_batches(n_batches, pairs_per_batch, seed, n_base=6, scene_seed=0):
  """Host-side synthetic input: `n_base` ray-cast LoKITTI-style pairs; every batch entry is one of them under a fresh
  random yaw + translation (applied to both scans of the pair), so voxelisation differs in every batch.
  `seed` (the rank) drives the poses only; the base scenes depend on `scene_seed`, which is the same on every rank: weak scaling
  means the same work per GPU, and rank-specific scenes differed by +-8 % in voxel count (585 k .. 683 k per 16-pair batch),
  which the max-over-ranks timing then reported as a scaling loss (round 1: 0.91 at 8 GPUs with no collective on the path)."""
  from gcl_b200 import synth
  rng = np.random.RandomState(seed)
  sensor = synth.NUSCENES if SENSOR["name"] == "NUSCENES" else synth.KITTI
  base = [synth.scan_pair(scene_seed=scene_seed * 7 + i, pair_seed=scene_seed * 13 + i, sensor=sensor) for i in range(n_base)]
  batches = []
  for b in range(n_batches):
    clouds = []
    for p in range(pairs_per_batch):
      x0, x1, _ = base[(b * pairs_per_batch + p) % n_base]
      yaw = rng.uniform(-np.pi, np.pi)
      c, s = np.cos(yaw), np.sin(yaw)
      R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
      t = np.array([rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-0.3, 0.3)], np.float32)
      clouds += [x0 @ R.T + t, x1 @ R.T + t]
    ptr = torch.tensor(np.cumsum([0] + [len(c) for c in clouds]), dtype=torch.int64)
    xyz = torch.from_numpy(np.ascontiguousarray(np.concatenate(clouds), dtype=np.float32))
    batches.append((xyz, ptr))
  return batches


class ClockSampler:
  """SM clock / throttle reasons sampled DURING the timed region, in-process through NVML (initialised up front, one
  light query every 10 ms).  An `nvidia-smi -lms` child process was used first: its start-up (NVML init + device
  enumeration, hundreds of ms under driver locks) landed inside the ~150 ms timed region and randomly stalled kernel
  launches (runs of 2400-2500 pairs/s next to 3400 on the same box); NVML in-process does not."""
  REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

  def __init__(self, gpu_index):
    self.h = self.nv = None
    self.sm, self.bits, self.run = [], 0, False
    self.period = float(os.environ.get("GCLB_CLOCK_PERIOD", "0.05"))
    try:
      import pynvml
      pynvml.nvmlInit()
      try:   # CUDA_VISIBLE_DEVICES may renumber: go through the UUID
        uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
        self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
      except Exception:
        self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
      self.nv = pynvml
      self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
    except Exception:
      self.h = None

  def _loop(self):
    nv = self.nv
    while self.run:
      try:
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
      except Exception:
        pass
      time.sleep(self.period)

  def start(self):
    if self.h is None:
      return
    self.run = True
    self.t = threading.Thread(target=self._loop, daemon=True)
    self.t.start()

  def stop(self):
    if self.h is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable"], "samples": 0}
    self.run = False
    if getattr(self, "t", None) is not None:
      self.t.join(timeout=1)
    reasons = [name for bit, name in self.REASONS if self.bits & bit]
    return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
            "samples": len(self.sm)}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle (restatement of MinkowskiEngine's CPU algorithm) on all host cores
# ---------------------------------------------------------------------------------------------------------------
def cpu_pair_step(model, xyz_pair, rng):
  """One pair through the oracle: voxelise both scans, forward both, 5000-subsample, mutual NN.  Returns #voxels."""
  import oracle.me_cpu as OME
  from oracle import matching as omatch
  feats = []
  nvox = 0
  for x in xyz_pair:
    t = torch.from_numpy(x)
    _, sel = OME.utils.sparse_quantize(t / VOXEL, return_index=True)
    c = torch.floor(t[sel] / VOXEL).int()
    C, F = OME.utils.sparse_collate([c], [torch.ones(len(c), 1)])
    with torch.no_grad():
      feats.append(model(OME.SparseTensor(F, coordinates=C)).F)
    nvox += len(c)
  i0 = rng.choice(len(feats[0]), min(SUBSAMPLE, len(feats[0])), replace=False)
  i1 = rng.choice(len(feats[1]), min(SUBSAMPLE, len(feats[1])), replace=False)
  omatch.mutual_nn(feats[0][i0], feats[1][i1], chunk=500)
  return nvox


def run_cpu(steps, warmup, n_pairs_per_step=1):
  import oracle.me_cpu as OME
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  model = seeded_model(OME)
  batches = make_batches(1, max(2, n_pairs_per_step), seed=0, n_base=2)
  xyz, ptr = batches[0]
  ptr = ptr.tolist()
  clouds = [xyz[ptr[i]:ptr[i + 1]].numpy() for i in range(len(ptr) - 1)]
  rng = np.random.RandomState(0)
  nvox = 0
  for s in range(warmup):
    cpu_pair_step(model, clouds[0:2], rng)
  t0 = time.perf_counter()
  for s in range(steps):
    for p in range(n_pairs_per_step):
      q = (s * n_pairs_per_step + p) % (len(clouds) // 2)
      nvox += cpu_pair_step(model, clouds[2 * q:2 * q + 2], rng)
  dt = time.perf_counter() - t0
  return dict(pairs_per_s=steps * n_pairs_per_step / dt, mvox_per_s=nvox / dt / 1e6, cores=cores, seconds=dt,
              ms_per_step=dt / steps * 1e3)


# ---------------------------------------------------------------------------------------------------------------
def conv_roofline(matcher, xyz_dev, ptr, peaks):
  """Instrumented pass: CUDA events around every sparse-conv launch of one step (on the launching stream)."""
  from gcl_b200 import ops
  eng = matcher.engine
  cm1, _ = ops.voxelize(xyz_dev, VOXEL, ptr)
  maps = eng.build_maps(cm1)
  cms, km = maps
  eng.forward(cm1, torch.ones((cm1.n, 1), device=xyz_dev.device), maps)   # conv1 adds the stride-1 3x3x3 table it emits
  tabs = {k: (v if isinstance(v, tuple) else (v,)) for k, v in km.items()}
  pair_counts = {k: int((v[0] >= 0).sum().item()) for k, v in tabs.items()}
  recs = []
  orig, orig_halo = ops.spconv_fwd, ops.spconv_fwd_halo

  def record(e0, e1, in0, in1, W, out, pairs, residual):
    K = W.shape[0] if W.dim() == 3 else 1
    cin, cout, n_out = in0.shape[1] + (in1.shape[1] if in1 is not None else 0), out.shape[1], out.shape[0]
    # every tensor once at its storage width (fp16 between the layers, fp32 descriptors) + the map + weights
    alg_bytes = (in0.element_size() * in0.shape[0] * cin + out.element_size() * n_out * cout + 8 * pairs
                 + W.element_size() * K * cin * cout)
    if residual is not None:
      alg_bytes += residual.element_size() * n_out * cout
    recs.append((e0, e1, alg_bytes, 2 * pairs * cin * cout))

  def timed(in0, W, nbr, n_out, in1=None, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig(in0, W, nbr, n_out, in1=in1, **kw)
    e1.record()
    pairs = n_out if nbr is None else next(pair_counts[k] for k, v in tabs.items() if any(t is nbr for t in v))
    record(e0, e1, in0, in1, W, out, pairs, kw.get("residual"))
    return out

  def timed_halo(in0, Wimg, halo, in1=None, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = orig_halo(in0, Wimg, halo, in1=in1, **kw)
    e1.record()
    pairs = next(pair_counts[k] for k, v in tabs.items() if any(t is halo for t in v))
    record(e0, e1, in0, in1, Wimg, out, pairs, kw.get("residual"))
    return out

  feats = torch.ones((cm1.n, 1), device=xyz_dev.device)
  for _ in range(2):  # first pass warms caches, second is recorded
    recs.clear()
    ops.spconv_fwd, ops.spconv_fwd_halo = timed, timed_halo
    try:
      # keep the GPU busy for ~10 ms while the host enqueues the whole forward, so that the events bracket back-to-back
      # kernel execution and not the host's launch latency
      torch.cuda._sleep(int(2e7))
      eng.forward(cm1, feats, maps)
    finally:
      ops.spconv_fwd, ops.spconv_fwd_halo = orig, orig_halo
  torch.cuda.synchronize()
  ms = [a.elapsed_time(b) for a, b, _, _ in recs]
  tot_ms, tot_b, tot_f = sum(ms), sum(r[2] for r in recs), sum(r[3] for r in recs)
  peak = peaks.get("hbm_gbs", 6650.0)
  ach = tot_b / (tot_ms * 1e-3) / 1e9
  traffic, traffic_src = None, None
  try:   # DRAM bytes per launch of the same kernels from an ncu capture -- accepted only if it was taken from THIS build
    from gcl_b200 import build as _b
    tj = json.load(open(os.path.join(ROOT, "profiles", "conv_traffic.json")))
    if tj.get("build_id") == _b.source_id():
      traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
    else:
      traffic_src = f"profiles/conv_traffic.json is stale (build {tj.get('build_id')} != {_b.source_id()}): not reported"
  except Exception:
    pass
  # K1 (voxelise + hash) and K2 (strided maps + kernel maps + row bucketing) with SURVEY 8(d)'s byte formulas, CUDA events around
  # each group of launches on the launching stream (GPU kept busy so that the events bracket execution, not launch latency)
  def timed_ms(fn, reps=5):
    fn(); torch.cuda.synchronize()
    torch.cuda._sleep(int(2e7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
      out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out
  P = xyz_dev.shape[0]
  k1_ms, (cm_k1, _) = timed_ms(lambda: ops.voxelize(xyz_dev, VOXEL, ptr))
  k1_bytes = 12 * P + 16 * cm_k1.n + 8 * cm_k1.n
  k2_ms, (cms2, km2) = timed_ms(lambda: eng.build_maps(ops.voxelize(xyz_dev, VOXEL, ptr)[0]))
  k2_ms -= k1_ms                       # build_maps needs a fresh coordinate map each time; its cost is subtracted
  n_of = {1: cms[1].n, 2: cms[2].n, 4: cms[4].n, 8: cms[8].n}
  k2_bytes = sum(16 * n_of[s] + 16 * n_of[2 * s] for s in (1, 2, 4))                     # strided maps: 16 N_in + 16 N_out
  geo = {"k3s1": (1, 1), "k3s2": (2, 2), "k3s4": (4, 4), "k3s8": (8, 8), "down1": (1, 2), "down2": (2, 4), "down4": (4, 8),
         "up1": (2, 1), "up2": (4, 2), "up4": (8, 4)}
  for name, (s_in, s_out) in geo.items():
    if name in pair_counts and name != "k3s1":   # k3s1 is emitted by conv1's fused probe, not by a K2 launch
      k2_bytes += 16 * n_of[s_out] + 8 * pair_counts[name] + 16 * n_of[s_in]
  roof_k = {"roofline_k1": {"bound": "hbm", "kernel": "gclb_voxelize (insert_rows + count_winners + scan + scatter_winners)",
                            "achieved": round(k1_bytes / (k1_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(k1_bytes / (k1_ms * 1e-3) / 1e9 / peak, 4), "ms_per_step": round(k1_ms, 3),
                            "algorithmic_bytes_per_step": int(k1_bytes), "traffic": None},
            "roofline_k2": {"bound": "hbm", "kernel": "strided maps + kernel maps (kmap_build[_up]) + row bucketing (rowkey_scatter, permute_rows)",
                            "achieved": round(k2_bytes / (max(k2_ms, 1e-6) * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                            "frac": round(k2_bytes / (max(k2_ms, 1e-6) * 1e-3) / 1e9 / peak, 4), "ms_per_step": round(k2_ms, 3),
                            "algorithmic_bytes_per_step": int(k2_bytes), "traffic": None,
                            "note": "SM-issue / L2-latency bound integer kernels (hash probes): far from the HBM roofline by nature"}}
  return {"bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), **roof_k,
          "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": int(tot_b / max(len(recs), 1)),
          "kernel": "spconv_fwd_tc_kernel (tcgen05 sparse conv, all 22 launches of a step incl. the 2 pointwise tail layers)", "launches": len(recs),
          "avg_launch_ms": round(tot_ms / len(recs), 4), "conv_ms_per_step": round(tot_ms, 3),
          "algorithmic_bytes_per_step": tot_b, "algorithmic_gflop_per_step": round(tot_f / 1e9, 2),
          "achieved_tflops": round(tot_f / (tot_ms * 1e-3) / 1e12, 2),
          "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback"}


def cpu_train_step(samples_host, steps):
  """oracle arm of the training workload: the same step (train-mode ResUNetBN2C forward, finest-contrastive loss, backward, SGD) on
  the CPU restatement of the MinkowskiEngine operators, on a bounded sample of the batch"""
  import oracle.me_cpu as OME
  from oracle import gcl_loss as oloss, groups as ogroups
  from gcl_b200.resunet import make_models
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  torch.manual_seed(0)
  model = make_models(OME)["ResUNetBN2C"](**MODEL)
  model.train()
  opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.8, weight_decay=1e-4)
  cs, lens = [], []
  for cl, _ in samples_host:
    for x in cl:
      cs.append(torch.floor(torch.from_numpy(x) / VOXEL).int()); lens.append(len(x))
  C, F = OME.utils.sparse_collate(cs, [torch.ones(len(c), 1) for c in cs])
  off = np.cumsum([0] + lens)
  gl, il, fl = [], [], []
  for s, (cl, Ts) in enumerate(samples_host):
    g, i, f = ogroups.colocation_groups(cl[0], cl[1:], Ts, 1.5 * VOXEL, 5)
    gl.append(np.asarray(g)); il.append(np.asarray(i) + off[3 * s]); fl.append(np.asarray(f))
  group, index, finest = np.concatenate(gl), np.concatenate(il), np.concatenate(fl)
  starts = np.concatenate([[0], np.cumsum(group)])
  ih = oloss.exhaustive_hash([index[starts[g]:starts[g + 1]] for g in range(len(group))], len(C))
  rng = np.random.RandomState(0)
  t0 = time.perf_counter()
  for _ in range(steps):
    opt.zero_grad()
    Fo = model(OME.SparseTensor(F, coordinates=C)).F
    sel = oloss.draw_selections(len(group), len(C), 256 * len(samples_host), 256 * len(samples_host), rng)
    pos, fin, neg = oloss.group_contrastive_loss(Fo, group, index, ih, finest, *sel, square_loss=True, with_finest=True)
    (pos + fin + neg).backward()
    opt.step()
  dt = time.perf_counter() - t0
  return dict(scans_per_s=steps * len(cs) / dt, seconds=dt, cores=cores, scans=len(cs), ms_per_step=dt / steps * 1e3)


def main_sweep(args):
  """BASELINE config 5: voxel-size sweep 0.1-0.5 m on a dense synthetic surface (6 M points): per voxel size one step = K1 voxelise +
  strided / kernel maps + ResUNetBN2C forward of the whole cloud; reported as Mvoxels/s per size, `value` = the 0.1 m (~1 M voxel) case"""
  rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
  workload = "voxel-size sweep 0.1-0.5 m on a dense synthetic surface (6M points): voxelise + maps + ResUNetBN2C(k5,32-d) forward per size"
  if args.impl == "reference":
    if rank != 0:
      return
    import oracle.me_cpu as OME
    from gcl_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    model = seeded_model(OME)
    pts = torch.from_numpy(synth.dense_surface(300_000, seed=3))      # bounded sample: 1/20 of the points, 0.3 m voxels
    t0 = time.perf_counter()
    nv = 0
    for _ in range(max(1, min(args.steps, 3))):
      _, sel = OME.utils.sparse_quantize(pts / 0.3, return_index=True)
      c = torch.floor(pts[sel] / 0.3).int()
      C, F = OME.utils.sparse_collate([c], [torch.ones(len(c), 1)])
      with torch.no_grad():
        model(OME.SparseTensor(F, coordinates=C))
      nv += len(c)
    dt = time.perf_counter() - t0
    v = nv / dt / 1e6
    print(json.dumps({"impl": "reference", "metric": "mvoxels_per_sec", "value": round(v, 4), "unit": "Mvoxels/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / max(1, min(args.steps, 3)) * 1e3, 1),
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": workload, "sample": "300k points at 0.3 m through oracle/"},
                      "cpu_baseline": {"value": round(v, 4), "unit": "Mvoxels/s", "cores": os.cpu_count(), "kind": "port",
                                       "sample": f"300k-point surface at 0.3 m voxels through oracle/ in {dt:.1f} s"},
                      "e2e": {"value": round(v, 4), "unit": "Mvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return
  import torch.distributed as dist
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  pin_rank_to_cores(local_rank, world)
  from gcl_b200 import MinkowskiEngine as ME, _lib, synth
  from gcl_b200.engine import ResUNetEngine
  lib = _lib.load()
  eng = ResUNetEngine(seeded_model(ME), device=dev)
  pts_h = torch.from_numpy(synth.dense_surface(6_000_000, seed=3 + rank)).pin_memory()
  pts = pts_h.to(dev)
  sizes = (0.1, 0.2, 0.3, 0.5)
  res, total_ms, total_vox, launches = {}, 0.0, 0, 0
  clocks = ClockSampler(local_rank)
  if rank == 0:
    clocks.start()
  for vs in sizes:
    for _ in range(max(args.warmup, 3)):
      eng.extract(pts, vs)
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    l0 = lib.gclb_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nv = 0
    for _ in range(args.steps):
      f, cm, _ = eng.extract(pts, vs)
      nv += cm.n
    float(f[0, 0])                               # D2H of a result element: the step's output is consumed
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches += lib.gclb_kernel_launches() - l0
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(max(2, args.steps // 4)):
      f, cm, _ = eng.extract(pts_h.to(dev, non_blocking=True), vs)
      f[:16].cpu()
    e3.record()
    torch.cuda.synchronize()
    ms_e2e = e2.elapsed_time(e3) / max(2, args.steps // 4)
    if world > 1:
      t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms, ms_e2e = t.tolist()
    res[str(vs)] = {"voxels": nv // args.steps, "ms_per_step": round(ms / args.steps, 3),
                    "mvoxels_per_sec": round(world * nv / (ms * 1e-3) / 1e6, 2), "e2e_mvoxels_per_sec": round(world * cm.n / (ms_e2e * 1e-3) / 1e6, 2)}
    total_ms += ms; total_vox += nv
  clk = clocks.stop() if rank == 0 else None
  if rank == 0:
    big = res["0.1"]
    print(json.dumps({"metric": "mvoxels_per_sec", "value": big["mvoxels_per_sec"], "unit": "Mvoxels/s", "n_gpus": world, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": big["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
                      "config": {"workload": workload, "headline_case": f"voxel 0.1 m, {big['voxels']} voxels", "per_voxel_size": res,
                                 "l2": "the 0.1 m case moves ~4 GB of activations per forward (>> 126 MB L2)"},
                      "e2e": {"value": big["e2e_mvoxels_per_sec"], "unit": "Mvoxels/s", "h2d_bytes_per_step": int(pts_h.numel() * 4),
                              "d2h_bytes_per_step": 16 * 32 * 4}, "gpu_launches": int(launches), "clocks": clk,
                      "roofline": {"bound": "hbm", "achieved": None, "peak": None, "unit": "GB/s", "frac": None, "traffic": None,
                                   "note": "per-kernel roofline: the pairs workload (same kernels)"}}))
  if world > 1:
    dist.destroy_process_group()


def main_train(args):
  """BASELINE config 4: GCL training step; whole-job scans/s = clouds of all ranks / max-over-ranks step time"""
  rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
  workload = (f"GCL training step: {args.samples} colocated scan groups x 3 nuScenes-shape scans per GPU (voxel {VOXEL} m), ResUNetBN2C "
              "train-mode fwd + finest-contrastive loss (positive groups + hardest negatives) + bwd + SGD; gradient all-reduce over ranks")
  if args.impl == "reference":
    if rank != 0:
      return
    from gcl_b200.training import synthetic_group_batch
    host = synthetic_group_batch(0, 1, VOXEL)      # bounded sample: one colocated group (3 scans) per step
    r = cpu_train_step(host, max(1, min(args.steps, 3)))
    line = {"impl": "reference", "metric": "gcl_train_scans_per_sec", "value": round(r["scans_per_s"], 4), "unit": "scans/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "samples_per_step": 1},
            "cpu_baseline": {"value": round(r["scans_per_s"], 4), "unit": "scans/s", "cores": r["cores"], "kind": "port",
                             "sample": f"1 colocated group (3 scans) per step through oracle/ autograd in {r['seconds']:.1f} s"},
            "e2e": {"value": round(r["scans_per_s"], 4), "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return
  import torch.distributed as dist
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  pin_rank_to_cores(local_rank, world)
  from gcl_b200 import _lib
  from gcl_b200.training import GclTrainStep
  lib = _lib.load()
  ts = GclTrainStep(dev, rank=rank, samples=args.samples, voxel=VOXEL, conv="tf32")
  clocks = ClockSampler(local_rank)

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  seg = lambda: torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
  quiet = 0
  for i in range(max(args.warmup, 3) + 12):        # warm up until the caching allocator stops growing
    s0 = seg()
    ts.step()
    quiet = quiet + 1 if seg() == s0 else 0
    if i >= max(args.warmup, 3) and quiet >= 3:
      break

  def timed(fn, steps):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms], device=dev, dtype=torch.float64)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      ms = t.item()
    return ms

  if rank == 0:
    clocks.start()
  l0 = lib.gclb_kernel_launches()
  ms = timed(lambda: ts.step(), args.steps)
  launches = lib.gclb_kernel_launches() - l0
  clk = clocks.stop() if rank == 0 else None
  # end to end: the batch starts in pinned host memory every step (H2D + voxelisation + positive-group construction + pair hashes
  # inside the timed region) and the loss comes back to the host
  clouds = [c for cl, _ in ts.host_batch for c in cl]
  pinned = torch.from_numpy(np.concatenate(clouds)).pin_memory()

  def step_e2e():
    ts._upload_and_group(ts.host_batch, pinned=pinned)
    return float(ts.step().detach())

  for _ in range(3):
    step_e2e()
  ms_e2e = timed(step_e2e, max(3, args.steps // 2)) / max(3, args.steps // 2) * args.steps
  # ranks must hold bit-identical weights after the exchange steps
  identical = True
  if world > 1:
    w = torch.cat([p.detach().reshape(-1) for p in ts.model.parameters()])
    lo, hi = w.clone(), w.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    identical = bool(torch.equal(lo, hi))
  n_scans = torch.tensor([float(ts.n_clouds)], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(n_scans, op=dist.ReduceOp.SUM)
  total_scans = n_scans.item() * args.steps
  if rank == 0:
    peaks = {}
    try:
      peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
      pass
    n_param = sum(p.numel() for p in ts.model.parameters())
    line = {"metric": "gcl_train_scans_per_sec", "value": round(total_scans / (ms * 1e-3), 2), "unit": "scans/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 operands (tcgen05 fwd/dgrad/wgrad), f32 accumulate, f32 BN/loss/SGD",
            "data": "synthetic",
            "config": {"workload": workload, "samples_per_gpu": args.samples, "scans_per_gpu": ts.n_clouds, "voxels_rank0": ts.n_rows,
                       "groups_rank0": int(ts.group.numel()),
                       "parallelism": f"colocated-scan groups sharded x{world}; one NCCL AVG all-reduce of the flat gradient buffer "
                                      f"({n_param * 4 / 1e6:.0f} MB) per step" if world > 1 else "single GPU (no collective)",
                       "ranks_bit_identical": identical,
                       "l2": "activations + gradients of a step (~1 GB) exceed the 126 MB L2"},
            "e2e": {"value": round(total_scans / (ms_e2e * 1e-3), 2), "unit": "scans/s", "h2d_bytes_per_step": int(pinned.numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3),
                    "what": "pinned host points -> H2D -> voxelise -> positive groups + pair hashes on the GPU -> step -> loss to host"},
            "gpu_launches": int(launches), "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": None, "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": None, "traffic": None,
                         "note": "per-kernel numbers of the training step: profiles/ (ncu of tools/profile_step.py --train)"}}
    if world == 1 and not args.no_cpu_baseline:
      from gcl_b200.training import synthetic_group_batch
      r = cpu_train_step(synthetic_group_batch(0, 1, VOXEL), 2)
      line["cpu_baseline"] = {"value": round(r["scans_per_s"], 4), "unit": "scans/s", "cores": r["cores"], "kind": "port",
                              "sample": f"2 steps of 1 colocated group (3 scans) through oracle/ autograd in {r['seconds']:.1f} s"}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def pin_rank_to_cores(local_rank, world):
  """N > 1: give every rank its own slice of the host cores this process may use.  A step is ~77 launches issued from one
  Python thread (~1.6 ms of host work per 3.9 ms step); eight such processes plus their reader threads migrating over the same
  cores was the round-1 limiter of the 8-GPU run (per-rank step 4.49 -> 4.92 ms with no collective).  GCLB_BENCH_PIN=0 disables."""
  if world <= 1 or os.environ.get("GCLB_BENCH_PIN", "1") == "0" or not hasattr(os, "sched_setaffinity"):
    return None
  try:
    allowed = sorted(os.sched_getaffinity(0))
    per = len(allowed) // world
    if per < 2:
      return None
    mine = allowed[local_rank * per:(local_rank + 1) * per]
    os.sched_setaffinity(0, mine)
    return mine
  except OSError:
    return None


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=50)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--pairs", type=int, default=16, help="scan pairs per step per GPU")
  ap.add_argument("--depth", type=int, default=2, help="batches in flight (PairMatcher.match_many); 1 = plain match() calls")
  ap.add_argument("--impl", default="gcl_b200", choices=["gcl_b200", "reference"])
  ap.add_argument("--algo", type=int, default=0, help="0 auto, 1 fp32 CUDA-core conv, 2 tcgen05 conv")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--workload", default="pairs", choices=["pairs", "nuscenes", "sweep", "train"],
                  help="pairs = BASELINE config 2 (headline: feature extraction + matching on KITTI-shape pairs); nuscenes = config 3 "
                       "(32-beam scans, 8 pairs per step); sweep = config 5 (voxel-size sweep on a dense surface, up to ~1M voxels); "
                       "train = config 4 (GCL training step, NCCL gradient all-reduce when N > 1)")
  ap.add_argument("--samples", type=int, default=4, help="train workload: colocated scan groups (3 scans each) per GPU per step")
  args = ap.parse_args()
  if args.workload == "train":
    return main_train(args)
  if args.workload == "sweep":
    return main_sweep(args)
  if args.workload == "nuscenes":
    SENSOR["name"] = "NUSCENES"
    if args.pairs == 16:
      args.pairs = 8                       # BASELINE config 3: batch of 8 pairs = 16 clouds
  args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

  rank = int(os.environ.get("RANK", 0))
  world = int(os.environ.get("WORLD_SIZE", 1))
  local_rank = int(os.environ.get("LOCAL_RANK", 0))
  shape = "nuScenes-shape 32-beam ~33k pts/scan" if SENSOR["name"] == "NUSCENES" else "64-beam ~130k pts/scan"
  workload = (f"LoKITTI-style synthetic scan pairs ({shape}, voxel {VOXEL} m): voxelise + kernel maps + "
              f"ResUNetBN2C(k5,32-d) fwd x2 + {SUBSAMPLE}-pt subsample + mutual-NN")

  if args.impl == "reference":
    if rank != 0:
      return
    r = run_cpu(args.steps, args.warmup, 1)
    line = {"impl": "reference", "metric": "scan_pairs_per_sec", "value": round(r["pairs_per_s"], 4), "unit": "pairs/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 2),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mvoxels_per_sec": round(r["mvox_per_s"], 4),
            "config": {"workload": workload, "pairs_per_step": 1},
            "cpu_baseline": {"value": round(r["pairs_per_s"], 4), "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                             "sample": "1 scan pair per step through oracle/ (CPU restatement of MinkowskiEngine's "
                                       "gather-GEMM-scatter; MinkowskiEngine itself is not buildable offline)"},
            "e2e": {"value": round(r["pairs_per_s"], 4), "unit": "pairs/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return

  import torch.distributed as dist
  torch.cuda.set_device(local_rank)
  dev = torch.device("cuda", local_rank)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  pin_rank_to_cores(local_rank, world)
  from gcl_b200 import MinkowskiEngine as ME, _lib
  from gcl_b200.pipeline import PairMatcher
  lib = _lib.load()
  peaks = {}
  try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
  except Exception:
    pass

  model = seeded_model(ME)
  matcher = PairMatcher(model, voxel=VOXEL, subsample=SUBSAMPLE, device=dev, seed=rank, algo=args.algo)
  n_batches = 3
  host = make_batches(n_batches, args.pairs, seed=rank)
  pinned = [(x.pin_memory(), p) for x, p in host]
  resident = [(x.to(dev), p) for x, p in host]

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed_region(fn, steps, source=None):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nvox = 0
    if source is not None and args.depth > 1:      # pipelined public API: `depth` batches in flight
      for s, out in enumerate(matcher.match_many((source[i % n_batches] for i in range(steps)), depth=args.depth)):
        nvox += fn(s, out)
    else:
      for s in range(steps):
        nvox += fn(s, None)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
      t = torch.tensor([ms, float(nvox)], device=dev, dtype=torch.float64)
      tmax = t.clone()
      dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
      dist.all_reduce(t, op=dist.ReduceOp.SUM)
      return tmax[0].item(), t[1].item()
    return ms, float(nvox)

  def step_resident(s, out=None):
    if out is None:
      x, p = resident[s % n_batches]
      out = matcher.match(x, p)
    # the correspondences come back to the host every step here too (a step whose result nobody reads is not a step);
    # the only difference to the e2e region is that the points are already resident
    pp = out["pair_ptr"].cpu()
    out["pairs"][:int(pp[-1])].cpu()
    matcher.check(out)                             # fp16-range status word of this batch (4 bytes)
    return out["n_voxels_total"]

  d2h_bytes = [0]

  def step_e2e(s, out=None):
    if out is None:
      x, p = pinned[s % n_batches]
      out = matcher.match(x, p)                    # H2D of the points inside
    pp = out["pair_ptr"].cpu()                     # D2H: correspondences of every pair
    k = int(pp[-1])
    pairs = out["pairs"][:k].cpu()
    matcher.check(out)
    d2h_bytes[0] = pairs.numel() * 8 + pp.numel() * 8 + 4
    return out["n_voxels_total"]

  clocks = ClockSampler(local_rank)      # NVML initialised before anything is timed
  for s in range(args.warmup):
    step_resident(s)
  def n_segments():
    return torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)

  def settle(fn, source, max_rounds=10):
    """untimed: run the exact loop of the timed region until PyTorch's caching allocator stops growing (no cudaMalloc --
    a device-wide synchronisation of unpredictable length -- for two full batch/stream periods), then leave spare cached
    blocks on every pipeline stream.  Without this 4 cudaMallocs landed inside the timed region and one run in four
    measured 1300-2600 instead of 3500 pairs/s."""
    quiet = 0
    for _ in range(max_rounds):
      s0 = n_segments()
      if args.depth > 1:
        for i, out in enumerate(matcher.match_many((source[j % n_batches] for j in range(n_batches * args.depth)), depth=args.depth)):
          fn(i, out)
      else:
        for i in range(n_batches):
          fn(i, None)
      quiet = quiet + 1 if n_segments() == s0 else 0
      if quiet >= 2:
        break
    for st in (matcher._streams or []) + [torch.cuda.current_stream()]:
      with torch.cuda.stream(st):
        spare = [torch.empty(256 << 20, dtype=torch.uint8, device=dev) for _ in range(2)]
        del spare
    torch.cuda.synchronize()

  settle(step_e2e, pinned)
  settle(step_resident, resident)       # the source of the first timed region last
  if rank == 0 and not os.environ.get("GCLB_NO_CLOCKS"):
    clocks.start()
  # a cudaMalloc inside a timed region is a device-wide synchronisation of unpredictable length: such a measurement is discarded
  # and the region re-measured (at most 3 attempts; the count of the reported attempt is in `cuda_mallocs_in_timed_region`)
  for attempt in range(3):
    l0 = lib.gclb_kernel_launches()
    seg0 = n_segments()
    ms, nvox = timed_region(step_resident, args.steps, resident)
    launches = lib.gclb_kernel_launches() - l0
    new_segments = n_segments() - seg0   # cudaMalloc calls inside the timed region
    if world > 1:
      t = torch.tensor([float(new_segments)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); new_segments = int(t.item())
    if new_segments == 0:
      break
  clk = clocks.stop() if rank == 0 else None
  settle(step_e2e, pinned, max_rounds=4)
  for attempt in range(3):
    seg1 = n_segments()
    ms_e2e, _ = timed_region(step_e2e, args.steps, pinned)
    new_segments_e2e = n_segments() - seg1
    if world > 1:
      t = torch.tensor([float(new_segments_e2e)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); new_segments_e2e = int(t.item())
    if new_segments_e2e == 0:
      break

  # registered pairs/s: the same path + SC2-PCR registration of every pair (scripts/test_kitti.py:180-182), transforms read back
  matcher.do_register = True

  def step_registered(s, out=None):
    if out is None:
      x, p = resident[s % n_batches]
      out = matcher.match(x, p)
    out["trans"].cpu()
    matcher.check(out)
    return out["n_voxels_total"]

  settle(step_registered, resident, max_rounds=4)
  reg_steps = max(5, args.steps // 2)
  ms_reg, _ = timed_region(step_registered, reg_steps, resident)
  matcher.do_register = False

  # end to end FROM FILE BYTES: every cloud is a KITTI-layout .bin (float32 x, y, z, reflectance) on local storage;
  # per step: read the files into the pinned staging buffer -> one H2D -> gclb_ingest_points -> the same path -> D2H
  import tempfile
  from gcl_b200 import ingest
  tmpdir = tempfile.mkdtemp(prefix=f"gclb_bench_r{rank}_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
  file_batches = []
  for b, (x, p) in enumerate(host):
    pl = p.tolist()
    paths = []
    for c in range(len(pl) - 1):
      fn = os.path.join(tmpdir, f"b{b}_{c:03d}.bin")
      ingest.write_velodyne_bin(fn, x[pl[c]:pl[c + 1]].numpy())
      paths.append(fn)
    file_batches.append(paths)
  readers = [ingest.ScanReader(capacity_points=max(x.shape[0] for x, _ in host) + 1024, threads=min(16, os.cpu_count() or 1))
             for _ in range(max(args.depth, 1) + 2)]
  import concurrent.futures as cf
  prefetch = cf.ThreadPoolExecutor(max_workers=1)

  def disk_source(n):
    # the next batch's files are read (into their own pinned buffer) by a background thread while this thread enqueues the
    # current batch; file reads release the GIL
    nxt = prefetch.submit(readers[0].read, file_batches[0]) if n else None
    for i in range(n):
      rec, ptr = nxt.result()
      if i + 1 < n:
        nxt = prefetch.submit(readers[(i + 1) % len(readers)].read, file_batches[(i + 1) % n_batches])
      yield ingest.points_to_device(rec, ptr, dev), ptr

  def run_disk(steps):
    nv = 0
    if args.depth > 1:
      for s_, out in enumerate(matcher.match_many(disk_source(steps), depth=args.depth)):
        nv += step_e2e(s_, out)
    else:
      for s_, (xd, ptr) in enumerate(disk_source(steps)):
        nv += step_e2e(s_, matcher.match(xd, ptr))
    return nv

  for _ in range(2):
    run_disk(n_batches * max(args.depth, 1))
  barrier()
  ed0, ed1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  ed0.record()
  run_disk(args.steps)
  ed1.record()
  barrier()
  ms_disk = ed0.elapsed_time(ed1)
  if world > 1:
    t = torch.tensor([ms_disk], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms_disk = t.item()
  import shutil
  prefetch.shutdown(wait=True)
  shutil.rmtree(tmpdir, ignore_errors=True)

  total_pairs = args.pairs * args.steps * world
  value = total_pairs / (ms * 1e-3)
  e2e_value = total_pairs / (ms_e2e * 1e-3)
  if rank != 0:
    if world > 1:
      dist.destroy_process_group()
    return
  roof = conv_roofline(matcher, resident[0][0], resident[0][1], peaks)
  h2d = int(np.mean([x.numel() * 4 for x, _ in host]))
  step_ws_mb = roof["algorithmic_bytes_per_step"] / 1e6
  line = {"metric": "scan_pairs_per_sec", "value": round(value, 2), "unit": "pairs/s", "n_gpus": world,
          "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "f16 operands, f32 accumulate" if (lib.gclb_has_tcgen05() and args.algo != 1) else "f32", "data": "synthetic",
          "mvoxels_per_sec": round(nvox / (ms * 1e-3) / 1e6, 3),
          "config": {"workload": workload, "pairs_per_step_per_gpu": args.pairs, "batches_in_flight": args.depth, "parallelism": f"pair-sharded x{world}, no collective",
                     "l2": f"inputs larger than L2: {n_batches} rotating batches, ~{step_ws_mb:.0f} MB algorithmic conv traffic per step vs 126 MB L2",
                     "conv_algo": ("tcgen05 kind::f16, fp16 activations between all layers (fp32 accumulate, fp32 BN scale/shift, fp32 descriptors); conv1 (Cin=1) fused hash-probe kernel"
                                   if (lib.gclb_has_tcgen05() and args.algo != 1) else "fp32 CUDA-core implicit GEMM")},
          "e2e": {"value": round(e2e_value, 2), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                  "d2h_bytes_per_step": int(d2h_bytes[0]), "ms_per_step": round(ms_e2e / args.steps, 3)},
          "e2e_from_disk": {"value": round(total_pairs / (ms_disk * 1e-3), 2), "unit": "pairs/s", "ms_per_step": round(ms_disk / args.steps, 3),
                            "file_bytes_per_step": int(np.mean([x.shape[0] * 16 for x, _ in host])),
                            "what": "KITTI-layout .bin files (tmpfs) -> readinto pinned buffer -> H2D -> gclb_ingest_points -> same path -> D2H"},
          "registered": {"metric": "registered_scan_pairs_per_sec", "value": round(args.pairs * reg_steps * world / (ms_reg * 1e-3), 2),
                         "unit": "pairs/s", "ms_per_step": round(ms_reg / reg_steps, 3), "steps": reg_steps,
                         "what": "same step + SC2-PCR registration of every pair on the GPU (5000 putative correspondences per pair, "
                                 "config_KITTI.json), 4x4 transforms read back"},
          "gpu_launches": int(launches), "cuda_mallocs_in_timed_region": [int(new_segments), int(new_segments_e2e)], "clocks": clk,
          "roofline_k1": roof.pop("roofline_k1"), "roofline_k2": roof.pop("roofline_k2"), "roofline": roof}
  if world == 1 and args.algo == 0 and lib.gclb_has_tcgen05():
    # the same step at the other two precisions of the engine: tf32 activation storage (tcgen05 kind::tf32, what
    # the drop-in ME module runs) and exact fp32 (CUDA-core kernels).  Short timed regions: context for the fp16 headline, not
    # headline numbers themselves.
    from gcl_b200.engine import ResUNetEngine
    prec = {}
    for name, kw in (("tf32_storage_tcgen05", dict(algo=0, half=False)), ("exact_fp32_cuda_core", dict(algo=1))):
      torch.cuda.empty_cache()
      m2 = PairMatcher(ResUNetEngine(model, device=dev, **kw), voxel=VOXEL, subsample=SUBSAMPLE, device=dev, seed=rank)
      def step_p(i):
        x, p_ = resident[i % n_batches]
        out = m2.match(x, p_)
        out["pairs"][:int(out["pair_ptr"].cpu()[-1])].cpu()
      for i in range(3):
        step_p(i)
      torch.cuda.synchronize()
      n_p = 6 if kw.get("algo") != 1 else 3
      a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a0.record()
      for i in range(n_p):
        step_p(i)
      a1.record()
      torch.cuda.synchronize()
      ms_p = a0.elapsed_time(a1) / n_p
      prec[name] = {"value": round(args.pairs / (ms_p * 1e-3), 2), "unit": "pairs/s", "ms_per_step": round(ms_p, 3), "steps": n_p}
      del m2
    torch.cuda.empty_cache()
    prec["fp16_storage_tcgen05"] = {"value": round(value, 2), "unit": "pairs/s", "ms_per_step": round(ms / args.steps, 3), "note": "the headline"}
    line["precisions"] = prec
  if world == 1 and not args.no_cpu_baseline:
    r = run_cpu(steps=5, warmup=1, n_pairs_per_step=1)
    line["cpu_baseline"] = {"value": round(r["pairs_per_s"], 4), "unit": "pairs/s", "cores": r["cores"], "kind": "port",
                            "sample": f"5 scan pairs (1 per step) of the same workload through oracle/ in {r['seconds']:.1f} s; "
                                      "oracle = CPU restatement of MinkowskiEngine's gather-GEMM-scatter, not MinkowskiEngine"}
  print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
