"""The inference hot path as one call: scan pairs in -> mutual feature correspondences out
(scripts/test_kitti.py:141-154 per pair: 2x ResUNet forward, 5000-point subsample, feature NN; with the mutual
filter of generalization_ETH/evaluate.py:63-77).  Batches several pairs per launch so that the ~60 kernel
launches of a forward are amortised."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops
from .engine import ResUNetEngine


class PairMatcher:
  def __init__(self, model, voxel: float = 0.3, subsample: int = 5000, device="cuda", seed: int = 0, algo: int = 0,
               register: bool = False):
    self.engine = model if isinstance(model, ResUNetEngine) else ResUNetEngine(model, device=device, algo=algo)
    self.voxel, self.subsample = float(voxel), int(subsample)
    self.device = torch.device(device)
    self.seed, self.calls = int(seed) * 1000003, 0
    self._streams = None
    self.do_register = bool(register)    # match() also runs SC2-PCR on every pair: out["trans"], out["reg_info"]

  @torch.no_grad()
  def match(self, xyz: torch.Tensor, cloud_ptr: torch.Tensor):
    """xyz float32 [P,3] (device, or pinned host: copied asynchronously), cloud_ptr int64 [2*n_pairs+1] (host);
    clouds are ordered (pair0.scan0, pair0.scan1, pair1.scan0, ...).
    Returns dict (all tensors on the device): pairs int64 [*,2] rows (i, j) = segment-local indices into the SUBSAMPLED
    rows of scan0 / scan1 of each pair (first pair_ptr[-1] rows valid), pair_ptr int64 [n_pairs+1], a_ptr/b_ptr the CSR
    of the subsamples, sel0 / sel1 the voxel rows they select (voxel row of pair p's i-th sample = sel0[a_ptr[p] + i]),
    unique_map (voxel row -> input point), cloud_rows (first voxel row of each cloud), n_voxels_total (host int)."""
    if not xyz.is_cuda:
      xyz = xyz.to(self.device, non_blocking=True)
    n_clouds = cloud_ptr.numel() - 1
    assert n_clouds % 2 == 0, "clouds come in pairs"
    feats, cm, umap = self.engine.extract(xyz, self.voxel, cloud_ptr)
    self.calls += 1
    n_pairs = n_clouds // 2
    # per-cloud 5000-row subsample + fused-gather mutual NN, all sizes stay on the device (no host sync)
    cloud_rows, sel_ptr, sel, cap = ops.subsample(cm, n_clouds, self.subsample, groups=2, seed=self.seed + self.calls)
    idx01, d01, idx10, d10, a_dev, b_dev, ws = ops.nn_search(feats, feats, sel_ptr[0], sel_ptr[1], both=True,
                                                             a_rows=sel[0], b_rows=sel[1], max_n=cap, max_m=cap)
    pairs, pair_ptr = ops.mutual_filter(idx01, idx10, a_dev, b_dev, ws)
    out = dict(pairs=pairs, pair_ptr=pair_ptr, sel0=sel[0], sel1=sel[1], a_ptr=a_dev, b_ptr=b_dev, idx01=idx01,
               idx10=idx10, unique_map=umap, cloud_rows=cloud_rows, n_voxels_total=cm.n, feats=feats, coords=cm.coords,
               status=self.engine.range_status)
    if self.do_register:
      out["trans"], out["reg_info"] = self.register(xyz, out)
    return out

  # scripts/SC2_PCR/config_json/config_KITTI.json
  SC2_KITTI = dict(d_thre=0.1, inlier_threshold=0.6, nms_radius=0.6, ratio=0.2, num_iterations=20, k1=30, k2=20, max_points=8000)

  @torch.no_grad()
  def register(self, xyz: torch.Tensor, out: dict, **sc2_cfg):
    """SC2-PCR registration of every pair of a matched batch (scripts/test_kitti.py:180-182: `matcher.estimator(...)` right
    after the features): the putative correspondences are sample i of scan 0 <-> its feature-space nearest neighbour among the
    samples of scan 1 (Matcher.match_pair), as point coordinates of `xyz` (the tensor given to match()).
    Returns (trans float32 [n_pairs, 4, 4] scan 0 -> scan 1, info int32 [n_pairs, 4]); everything stays on the device."""
    from . import registration
    cfg = dict(self.SC2_KITTI, **sc2_cfg)
    a_ptr, b_ptr = out["a_ptr"], out["b_ptr"]
    n_pairs = a_ptr.numel() - 1
    cap = self.subsample if self.subsample > 0 else int(out["n_voxels_total"])
    bound = n_pairs * cap
    dev = out["idx01"].device
    xyz = xyz if xyz.is_cuda else xyz.to(dev, non_blocking=True)
    src = torch.empty((bound, 3), dtype=torch.float32, device=dev)
    tgt = torch.empty((bound, 3), dtype=torch.float32, device=dev)
    _lib.call("gclb_corr_points", _lib.ptr(xyz.contiguous()), _lib.ptr(out["unique_map"]), _lib.ptr(out["sel0"]), _lib.ptr(out["sel1"]),
              _lib.ptr(a_ptr), _lib.ptr(b_ptr), _lib.ptr(out["idx01"]), n_pairs, bound, _lib.ptr(src), _lib.ptr(tgt), _lib.stream())
    return registration.sc2_pcr_batch(src, tgt, a_ptr, cap, **cfg)

  @staticmethod
  def check(out):
    """read the batch's device status word (4 bytes; do it next to the read-back of the correspondences) and raise
    GclbError if an fp16-stored activation saturated or sank into the subnormal range during this batch's forward"""
    if out.get("status") is not None:
      _lib.check_status(out["status"], "PairMatcher.match (fp16 activations)")

  def match_many(self, batches, depth: int = 2):
    """Throughput API: iterate over `(xyz, cloud_ptr)` batches and yield `match()` results in order, keeping `depth`
    batches in flight on alternating CUDA streams.  A batch has one host synchronisation (the voxel / strided-map row
    counts); with two streams that bubble -- and the launch gaps of the many small map-building kernels -- are filled by
    the other batch's convolutions.  Each result is complete (its stream synchronised) when it is yielded."""
    if self._streams is None or len(self._streams) < depth:
      self._streams = [torch.cuda.Stream(device=self.device) for _ in range(depth)]
    cur = torch.cuda.current_stream()
    pending = []
    for i, (xyz, cloud_ptr) in enumerate(batches):
      st = self._streams[i % depth]
      st.wait_stream(cur)
      with torch.cuda.stream(st):
        out = self.match(xyz, cloud_ptr)
        ev = torch.cuda.Event()
        ev.record(st)
      pending.append((out, ev))
      if len(pending) >= depth:
        yield self._hand_over(pending.pop(0), cur)
    for item in pending:
      yield self._hand_over(item, cur)

  @staticmethod
  def _hand_over(item, consumer_stream):
    """the result tensors were allocated on a side stream's pool: wait for the batch, then tell the caching allocator
    that the consumer's stream uses them too, so a later batch on the side stream cannot recycle a block that
    asynchronous consumer work is still reading"""
    out, ev = item
    ev.synchronize()
    for v in out.values():
      if isinstance(v, torch.Tensor) and v.is_cuda:
        v.record_stream(consumer_stream)
    return out
