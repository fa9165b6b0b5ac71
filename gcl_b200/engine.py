"""Fused eval-mode ResUNet forward: the inference hot path (scripts/test_kitti.py:141-154).

The op-by-op MinkowskiEngine surface launches conv, BatchNorm, ReLU, add and cat separately like the reference
(~75 calls).  In eval mode all of that folds into the convolution kernels:
    BatchNorm(eval) -> per-channel scale/shift in the conv epilogue,   `out += residual`, ReLU -> epilogue,
    ME.cat(a, b)    -> two-source gather,   conv1_tr + ReLU + final + bias + L2 normalise -> one pointwise kernel,
so a forward is 3 strided-map builds, 11 kernel-map builds, 21 sparse-conv launches and 1 tail launch, with one
host synchronisation (the strided-map row counts).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from . import _lib, ops


def _fold(conv, norm, device):
  W = conv.kernel.detach().to(device=device, dtype=torch.float32).contiguous()
  bn = norm.bn
  scale = (bn.weight.detach() if bn.weight is not None else 1.0) * torch.rsqrt(bn.running_var.detach() + bn.eps)
  shift = (bn.bias.detach() if bn.bias is not None else 0.0) - bn.running_mean.detach() * scale
  return W, scale.to(device).float().contiguous(), shift.to(device).float().contiguous()


class _Tables:
  """name -> kernel map (table or (table, sorted table, perm, tile masks)); a table built on a side stream carries an
  event that the consuming stream waits on at its first use."""

  def __init__(self):
    self._t, self._ev = {}, {}

  def put(self, name, t, ev):
    self._t[name], self._ev[name] = t, ev

  def alias(self, name, other):
    self._t[name], self._ev[name] = self._t[other], self._ev.get(other)

  def __contains__(self, name):
    return name in self._t

  def __getitem__(self, name):
    ev = self._ev.get(name)
    if ev is not None:
      torch.cuda.current_stream().wait_event(ev)
      self._ev[name] = None
    return self._t[name]

  def items(self):
    return [(k, self[k]) for k in self._t]


class ResUNetEngine:
  """Build from any ResUNet2-family module (reference class or gcl_b200.resunet) holding the trained weights."""

  def __init__(self, model, device="cuda", algo: int = 0, half: bool = True):
    """algo: 0 = tcgen05 kernels where a layer qualifies, 1 = exact-fp32 CUDA-core kernels everywhere.
    half: store the activations of every layer whose input channels are a multiple of 64 as fp16 and run it with
    kind::f16 (fp32 accumulation): same 10-bit mantissa as the kind::tf32 operands but rounded to nearest instead of
    truncated (measured closer to exact fp32 than tf32: tools/precision_study.py), and half the gather bytes."""
    self.device = torch.device(device)
    self.algo = algo
    self.half = bool(half) and algo != 1
    self.normalize = bool(model.normalize_feature)
    d = self.device
    self.conv1_ks = model.conv1.kernel_size
    f = lambda c, n: _fold(c, n, d)
    self.p = {"conv1": f(model.conv1, model.norm1)}
    for name in ("block1", "block2", "block3", "block4", "block4_tr", "block3_tr", "block2_tr"):
      b = getattr(model, name)
      self.p[name + ".1"] = f(b.conv1, b.norm1)
      self.p[name + ".2"] = f(b.conv2, b.norm2)
    for l in (2, 3, 4):
      self.p[f"conv{l}"] = f(getattr(model, f"conv{l}"), getattr(model, f"norm{l}"))
      self.p[f"conv{l}_tr"] = f(getattr(model, f"conv{l}_tr"), getattr(model, f"norm{l}_tr"))
    T = type(model).TR_CHANNELS
    self.SPLIT = {"conv3_tr": T[4], "conv2_tr": T[3]}
    self.W1 = model.conv1_tr.kernel.detach().to(d).float().contiguous()
    self.W2 = model.final.kernel.detach().to(d).float().contiguous()
    self.bias = model.final.bias.detach().to(d).float().reshape(-1).contiguous() if model.final.bias is not None else None
    self.last_maps = None
    # fp16-range monitor: one (flags, max |y|) pair per fp16-storing launch of a forward + a status word that check_range()
    # reads (the reference computes in fp32; saturation or a tensor sunk into the fp16 subnormals must not be silent)
    self.monitor_range = True
    self.range_status = None
    self._mon = None
    self._mon_i = 0

    # tensor-core weight copies ([K, Cout, Cin], tf32-rounded) for every layer the tcgen05 kernel covers
    self.tc = {}
    self.tc16 = {}         # fp16 images of the layers that run with fp16 activations
    self.tail_tc = None
    self.sort_rows = True
    self.use_halo = os.environ.get("GCLB_HALO", "0") != "0"   # same-map 3x3x3 convolutions on fp16 rows run the halo-staging kernel (csrc/spconv_halo.cu)
    self._side = None      # side streams for overlapped kernel-map construction
    # conv1 with a narrow input: fuse the kernel map into the convolution (no 4*K*N-byte table for the 5^3 kernel)
    self.conv1_probe = (self.p["conv1"][0].shape[1] <= 4 and self.p["conv1"][0].shape[2] <= 128)
    if algo != 1:
      for key, (W, _, _) in self.p.items():
        K, cin, cout = W.shape
        c0 = self.SPLIT.get(key, cin)
        if ops.tc_supported(c0, cin - c0, cout, K):
          self.tc[key] = ops.weights_to_tc(W)
        if self.half and ops.tc_supported(c0, cin - c0, cout, K, half=True) and self._fits_fp16(W):
          self.tc16[key] = ops.weights_to_tc(W, half=True, c0=c0)
      # pointwise tail on the tensor cores: [y1 | s1] W1 -> ReLU -> W2 + bias -> L2 normalise (fused in the epilogue)
      c_s1 = type(model).CHANNELS[1]
      c_y1 = self.W1.shape[0] - c_s1
      self.tail_tc = None
      if (ops.tc_supported(c_y1, c_s1, self.W1.shape[1], 1) and ops.tc_supported(self.W2.shape[0], 0, self.W2.shape[1], 1)
          and self.W2.shape[1] == 32):
        if self.half and self._fits_fp16(self.W1) and self._fits_fp16(self.W2):
          self.tail_tc = (ops.weights_to_tc(self.W1, half=True, c0=c_y1), ops.weights_to_tc(self.W2, half=True))
        else:
          self.tail_tc = (ops.weights_to_tc(self.W1), ops.weights_to_tc(self.W2))

  @staticmethod
  def _fits_fp16(W) -> bool:
    """construction-time guard for the fp16 weight images (the conversion saturates): a layer whose weights leave the fp16
    range, or sit wholly below 2^-11, keeps fp32 weights and runs on the kind::tf32 path instead"""
    a = float(W.abs().max())
    return a <= 65504.0 and (a == 0.0 or a >= 2.0 ** -11)

  # first-source channel count of the layers that read a concatenation (ME.cat fused into the gather)
  SPLIT = {}

  # ---- building blocks
  def _want(self, key):
    """storage dtype a layer reads its inputs in"""
    return torch.float16 if key in self.tc16 else torch.float32

  MAX_MONITORS = 32

  def _mon_begin(self, device):
    if not (self.half and self.monitor_range):
      self._mon = None
      return
    buf = torch.zeros(2 * self.MAX_MONITORS + 1, dtype=torch.int32, device=device)   # monitors + the status word
    self._mon, self.range_status, self._mon_i = buf, buf[2 * self.MAX_MONITORS:], 0

  def _mon_next(self, out_dtype):
    """arm the thread-local monitor pointer for the next launch when it stores fp16"""
    if self._mon is None:
      return
    if out_dtype == torch.float16:
      assert self._mon_i < self.MAX_MONITORS
      _lib.call("gclb_spconv_set_range_monitor", self._mon.data_ptr() + 8 * self._mon_i)
      self._mon_i += 1
    else:
      _lib.call("gclb_spconv_set_range_monitor", None)

  def _mon_end(self):
    if self._mon is None:
      return
    _lib.call("gclb_spconv_set_range_monitor", None)
    if self._mon_i:
      _lib.call("gclb_range_check", self._mon.data_ptr(), self._mon_i, self.range_status.data_ptr(), _lib.stream())

  def check_range(self):
    """one host read: raise GclbError if the last forward saturated an fp16 activation or produced a tensor in the fp16
    subnormal range (the pipeline folds this word into the read-back it does anyway)"""
    if self.range_status is not None:
      _lib.check_status(self.range_status, "ResUNetEngine forward (fp16 activations)")

  def _conv(self, key, x, nbr, n_out, x2=None, residual=None, relu=False, out_dtype=torch.float32):
    W, sc, sh = self.p[key]
    dt = self._want(key)
    self._mon_next(out_dtype if key in self.tc else None)
    # no-ops on the tuned path (every producer already writes what its consumer reads); a model variant whose channel
    # widths mix eligible and ineligible layers converts here
    x = x if x.dtype == dt else x.to(dt)
    x2 = x2 if (x2 is None or x2.dtype == dt) else x2.to(dt)
    residual = residual if (residual is None or residual.dtype == dt) else residual.to(dt)
    if key in self.tc:
      perm = mask = None
      is_sorted = True
      if isinstance(nbr, tuple):          # (table, sorted copy or None, perm, tile masks[, halo map]) from build_maps
        halo = nbr[4] if len(nbr) > 4 else None
        if halo is not None and dt == torch.float16:
          return ops.spconv_fwd_halo(x, self.tc16[key], halo, in1=x2, scale=sc, shift=sh, residual=residual, relu=relu,
                                     out_dtype=out_dtype)
        table, srt, perm, mask = nbr[:4]
        nbr, is_sorted = (srt, True) if srt is not None else (table, False)
      Wimg = self.tc16[key] if dt == torch.float16 else self.tc[key]
      return ops.spconv_fwd(x, Wimg, nbr, n_out, in1=x2, scale=sc, shift=sh, residual=residual, relu=relu,
                            algo=2, row_perm=perm, tile_mask=mask, nbr_is_sorted=is_sorted, out_dtype=out_dtype)
    if isinstance(nbr, tuple):
      nbr = nbr[0]
    y = ops.spconv_fwd(x, W, nbr, n_out, in1=x2, scale=sc, shift=sh, residual=residual, relu=relu, algo=1)
    return y if out_dtype == torch.float32 else y.to(out_dtype)

  def _block(self, name, x, nbr, out_dtype=torch.float32):
    n = x.shape[0]
    t = self._conv(name + ".1", x, nbr, n, relu=True, out_dtype=self._want(name + ".2"))
    return self._conv(name + ".2", t, nbr, n, residual=x, relu=True, out_dtype=out_dtype)

  # residual blocks convolving over the same-map 3x3x3 table of each tensor stride
  LEVEL_BLOCKS = {1: ("block1", "block2_tr"), 2: ("block2", "block3_tr"), 4: ("block3", "block4_tr"), 8: ("block4",)}

  def _halo_level(self, s) -> bool:
    """every convolution over the stride-s same-map table runs on fp16 rows -> build the halo-staging metadata for it"""
    return (self.use_halo and self.half and
            all((b + sfx) in self.tc16 for b in self.LEVEL_BLOCKS[s] for sfx in (".1", ".2")))

  def _bucket(self, t, keys, halo):
    """row-bucketed form of a neighbour table for the tensor-core kernels: (table, sorted copy, perm, tile masks[, halo])"""
    if halo:   # the halo kernel reads the original table through the permutation once, at build time: no sorted copy
      _, perm, mask = ops.kernel_map_sort(t, keys, copy=False)
      return (t, None, perm, mask, ops.kernel_map_halo(t, perm))
    # copy=True: a physically re-ordered table.  Reading the original table through the permutation inside the conv
    # (copy=False, 108 4-byte cp.async per lane per tile) was measured 1.6x slower end to end.
    return (t,) + ops.kernel_map_sort(t, keys, copy=True)

  # order in which forward() first touches each table
  TABLE_ORDER = ("c1", "k3s1", "down1", "k3s2", "down2", "k3s4", "down4", "k3s8", "up4", "up2", "up1")

  def build_maps(self, cm1: ops.CoordMap, overlap: bool = False):
    """strided maps + all kernel maps of one forward (cacheable by the caller for repeated forwards).

    * The three strided levels are chained on the device and finished -- together with cm1 if it came from
      voxelize(sync=False) -- by ONE host read of the row counts.
    * The 10-11 kernel maps are independent chains of small, latency-bound kernels (probe, key histogram, scan, scatter,
      permute).  With `overlap` they are enqueued round-robin on side streams in the order forward() needs them, so they
      overlap each other and the first convolutions; forward() waits on a table's event right before its first use.
      Measured on B200 (round 1): SLOWER (7.0 vs 4.5 ms per 8-pair step) -- cross-stream allocator traffic and the
      persistent 200 KB-smem conv CTAs leave no room for co-residency -- so it is off by default."""
    cm2 = ops.stride_map(cm1, 2, sync=False)
    cm4 = ops.stride_map(cm2, 2, sync=False)
    cm8 = ops.stride_map(cm4, 2, sync=False)
    ops.finish_maps([cm1, cm2, cm4, cm8])
    cms = {1: cm1, 2: cm2, 4: cm4, 8: cm8}
    sort = bool(self.tc) and self.sort_rows     # row-bucketed copies for the tensor-core kernel

    def table(in_cm, out_cm, ks, transposed=False, tc=True, halo=False):
      if sort and tc:
        t, keys = ops.kernel_map(in_cm, out_cm, ks, transposed=transposed, with_keys=True)
        return self._bucket(t, keys, halo)
      return ops.kernel_map(in_cm, out_cm, ks, transposed=transposed)

    recipes = {}
    # the fused-probe conv1 (odd kernel >= 3) writes the stride-1 3x3x3 table itself, inside forward()
    fused_k3s1 = self.conv1_probe and self.conv1_ks >= 3 and self.conv1_ks % 2 == 1
    if self.conv1_ks != 1 and not self.conv1_probe:
      recipes["c1"] = lambda: table(cm1, cm1, self.conv1_ks, tc=(self.conv1_ks == 3))
    for s in (1, 2, 4, 8):
      if not (s == 1 and ((self.conv1_ks == 3 and "c1" in recipes) or fused_k3s1)):
        recipes[f"k3s{s}"] = (lambda s=s: table(cms[s], cms[s], 3, halo=self._halo_level(s)))
    for s in (1, 2, 4):
      recipes[f"down{s}"] = (lambda s=s: table(cms[s], cms[2 * s], 3))
      recipes[f"up{s}"] = (lambda s=s: table(cms[2 * s], cms[s], 3, transposed=True))

    km = _Tables()
    main = torch.cuda.current_stream()
    if not overlap:
      for name in self.TABLE_ORDER:
        if name in recipes:
          km.put(name, recipes[name](), None)
    else:
      if self._side is None:
        self._side = [torch.cuda.Stream(device=self.device) for _ in range(3)]
      for s_ in self._side:
        s_.wait_stream(main)                      # coordinate maps are ready
      i = 0
      for name in self.TABLE_ORDER:
        if name not in recipes:
          continue
        st = self._side[i % len(self._side)]
        i += 1
        with torch.cuda.stream(st):
          t = recipes[name]()
          ev = torch.cuda.Event()
          ev.record(st)
        for x in (t if isinstance(t, tuple) else (t,)):
          if x is not None:
            x.record_stream(main)                 # consumed on the main stream: keep the allocator from recycling early
        km.put(name, t, ev)
    if "k3s1" not in recipes and "c1" in recipes:
      km.alias("k3s1", "c1")
    return cms, km

  @torch.no_grad()
  def forward(self, cm1: ops.CoordMap, feats: torch.Tensor, maps=None) -> torch.Tensor:
    """feats float32 [N, Cin] on the rows of cm1 -> descriptors float32 [N, out] on the same rows."""
    cms, km = maps if maps is not None else self.build_maps(cm1)
    self.last_maps = (cms, km)
    n1, n2, n4, n8 = cms[1].n, cms[2].n, cms[4].n, cms[8].n
    x = feats.contiguous().float()
    self._mon_begin(x.device)
    if self.conv1_probe:
      W, sc, sh = self.p["conv1"]
      od = self._want("block1.1")
      self._mon_next(od)
      if "k3s1" in km:
        c1 = ops.spconv_fwd_probe(x, W, cms[1], self.conv1_ks, scale=sc, shift=sh, out_dtype=od)
      else:   # conv1's inner probes are the stride-1 3x3x3 kernel map: emitted by the same kernel, bucketed here
        c1, (t, keys) = ops.spconv_fwd_probe(x, W, cms[1], self.conv1_ks, scale=sc, shift=sh, emit_k3=True, out_dtype=od)
        sort = bool(self.tc) and self.sort_rows
        km.put("k3s1", self._bucket(t, keys, self._halo_level(1)) if sort else t, None)
    else:
      c1 = self._conv("conv1", x, km["c1"] if self.conv1_ks != 1 else None, n1, out_dtype=self._want("block1.1"))
    # every layer writes the storage dtype its consumer reads (fp16 between the 64..256-channel layers, fp32 at the
    # 32-channel stride-1 level and for the descriptors), converting in its epilogue
    w = self._want
    s1 = self._block("block1", c1, km["k3s1"], out_dtype=w("conv2"))
    s2 = self._block("block2", self._conv("conv2", s1, km["down1"], n2, out_dtype=w("block2.1")), km["k3s2"], w("conv3"))
    s4 = self._block("block3", self._conv("conv3", s2, km["down2"], n4, out_dtype=w("block3.1")), km["k3s4"], w("conv4"))
    s8 = self._block("block4", self._conv("conv4", s4, km["down4"], n8, out_dtype=w("block4.1")), km["k3s8"], w("conv4_tr"))
    y4 = self._block("block4_tr", self._conv("conv4_tr", s8, km["up4"], n4, out_dtype=w("block4_tr.1")), km["k3s4"],
                     w("conv3_tr"))
    y2 = self._block("block3_tr", self._conv("conv3_tr", y4, km["up2"], n2, x2=s4, out_dtype=w("block3_tr.1")), km["k3s2"],
                     w("conv2_tr"))
    tail_dt = self.tail_tc[0].dtype if getattr(self, "tail_tc", None) is not None else torch.float32
    y1 = self._block("block2_tr", self._conv("conv2_tr", y2, km["up1"], n1, x2=s2, out_dtype=w("block2_tr.1")), km["k3s1"],
                     tail_dt)
    s1 = s1 if s1.dtype == tail_dt else s1.to(tail_dt)        # no-op on the tuned path
    if getattr(self, "tail_tc", None) is not None:
      self._mon_next(y1.dtype)
      h = ops.spconv_fwd(y1, self.tail_tc[0], None, n1, in1=s1, relu=True, algo=2)
      self._mon_end()
      return ops.spconv_fwd(h, self.tail_tc[1], None, n1, shift=self.bias, normalize=self.normalize, algo=2,
                            out_dtype=torch.float32)
    self._mon_end()
    return ops.pointwise_tail(y1, s1, self.W1, self.W2, self.bias, normalize=self.normalize)

  @torch.no_grad()
  def extract(self, xyz: torch.Tensor, voxel: float, cloud_ptr: Optional[torch.Tensor] = None):
    """K1 -> K2 -> K3 for a batch of clouds (xyz float32 [P,3] on the device, cloud_ptr int64 [n+1]).
    Returns (descriptors [V, out], CoordMap, unique_map): row v describes point xyz[unique_map[v]]."""
    cm1, umap = ops.voxelize(xyz, voxel, cloud_ptr, sync=False)
    maps = self.build_maps(cm1)                     # the only host synchronisation of the forward
    feats = torch.ones((cm1.n, 1), dtype=torch.float32, device=xyz.device)
    return self.forward(cm1, feats, maps), cm1, umap[:cm1.n]
