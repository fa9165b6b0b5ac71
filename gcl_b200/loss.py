"""GCL group-wise contrastive loss on libgclb200's fused K5 kernels.

Mirrors /root/reference/lib/colocation_trainer.py:
  finest_contrastive_loss    :430-535   (pos = variance-to-mean, finest-to-mean, hardest negative)
  location_contrastive_loss  :734-809   (no finest term, non-squared positive term)
Same arguments and return value `(pos_loss, finest_loss, neg_loss)`; the host-side random selections are drawn
with the same `np.random` calls in the same order as the reference (:456-459, :506-507), so a seeded run picks
the same groups / rows.  The three returned scalars are outputs of a custom autograd Function whose backward runs the
fused gradient kernels with the upstream gradients autograd hands it (device scalars, no host read), so the trainer's
`pos_loss /= iter_size ... (pos_w * pos + finest_w * finest + neg_w * neg).backward()` (:874-879) works unchanged for
any weights / scaling; the constructor's `*_weight` arguments are kept for API compatibility and are not needed.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import call, ptr, stream


def _exhaustive_hash(index_split, M):
  """util/misc.py:29-36 (host-side, part of the loader's collate; kept for API parity)."""
  out = []
  for idx in index_split:
    idx = np.asarray(idx.cpu() if isinstance(idx, torch.Tensor) else idx, dtype=np.int64)
    for i in range(len(idx) - 1):
      out.append(np.minimum(idx[i] + idx[i + 1:] * M, idx[i] * M + idx[i + 1:]))
  return np.concatenate(out, axis=0) if out else np.zeros(0, np.int64)


class _GroupLossFn(torch.autograd.Function):
  """forward: gclb_group_loss (forward only); backward: gclb_group_loss_bwd with the ACTUAL upstream gradients of the
  three losses, handed to the kernels as a device float[3] (no host read).  Any scaling of the losses before
  `.backward()` -- the reference's `loss / iter_size` (lib/colocation_trainer.py:874-878), AMP loss scaling, using only
  one of the terms -- therefore back-propagates correctly."""

  @staticmethod
  def forward(ctx, F_out, args):
    (group_ptr, index, finest_pos, pos_sel, sel1, sel2, keys, thr, square) = args
    F_c = F_out.detach().contiguous().float()
    N, Cd = F_c.shape
    dev = F_c.device
    lib = _lib.load()
    losses = torch.empty(4, dtype=torch.float32, device=dev)
    ws = torch.empty(int(lib.gclb_loss_workspace_bytes(pos_sel.numel(), sel1.numel())), dtype=torch.uint8, device=dev)
    call("gclb_group_loss", ptr(F_c), N, Cd, ptr(group_ptr), ptr(index), ptr(finest_pos), ptr(pos_sel),
         pos_sel.numel(), ptr(sel1), ptr(sel2), sel1.numel(), ptr(keys), keys.numel(), thr[0], thr[1], thr[2],
         int(square), None, ptr(losses), None, ptr(ws), stream())
    ctx.save_for_backward(F_c)
    ctx.args = args
    # three separately owned 0-d tensors (not views of one buffer): the reference divides them in place
    # (`pos_loss /= iter_size`, :874-876), which autograd forbids on views returned by a multi-output Function
    pos, fin, neg = losses[0].clone(), losses[1].clone(), losses[2].clone()
    return pos, fin, neg

  @staticmethod
  def backward(ctx, g_pos, g_fin, g_neg):
    (F_c,) = ctx.saved_tensors
    (group_ptr, index, finest_pos, pos_sel, sel1, sel2, keys, thr, square) = ctx.args
    dev = F_c.device
    z = torch.zeros((), dtype=torch.float32, device=dev)
    up = torch.stack([(g if g is not None else z).to(torch.float32).reshape(()) for g in (g_pos, g_fin, g_neg)]).contiguous()
    N, Cd = F_c.shape
    lib = _lib.load()
    grad = torch.zeros_like(F_c)
    scratch = torch.empty(4, dtype=torch.float32, device=dev)
    ws = torch.empty(int(lib.gclb_loss_workspace_bytes(pos_sel.numel(), sel1.numel())), dtype=torch.uint8, device=dev)
    call("gclb_group_loss_bwd", ptr(F_c), N, Cd, ptr(group_ptr), ptr(index), ptr(finest_pos), ptr(pos_sel),
         pos_sel.numel(), ptr(sel1), ptr(sel2), sel1.numel(), ptr(keys), keys.numel(), thr[0], thr[1], thr[2],
         int(square), ptr(up), ptr(scratch), ptr(grad), ptr(ws), stream())
    return grad, None


class GroupContrastiveLoss:
  """loss = GroupContrastiveLoss(pos_thresh=.1, neg_thresh=1.4, finest_thresh=.2, square_loss=True,
                                 pos_weight=1, finest_weight=1, neg_weight=1)
  pos, finest, neg = loss.finest_contrastive_loss(F_out, group, index, index_hash, finest_flag,
                                                  max_pos_cluster=1024, max_hn_samples=1024)"""

  def __init__(self, pos_thresh=0.1, neg_thresh=1.4, finest_thresh=0.2, square_loss=True, pos_weight=1.0,
               finest_weight=1.0, neg_weight=1.0, rng=np.random):
    self.pos_thresh, self.neg_thresh, self.finest_thresh = pos_thresh, neg_thresh, finest_thresh
    self.square_loss = square_loss
    self.weights = (float(pos_weight), float(finest_weight), float(neg_weight))
    self.rng = rng

  def prepare(self, group, index, index_hash, finest_flag, device):
    """Batch-dependent (step-independent) inputs of the loss on the device: CSR of the groups, position of the finest member of
    every group, sorted positive-pair keys.  The colocation loaders produce these once per batch (the reference builds
    `index_hash` in its collate function); the trainer passes the result as `prepared=` and the per-step work is the kernels."""
    dev = torch.device(device)
    on_dev = isinstance(group, torch.Tensor) and group.is_cuda
    g = torch.as_tensor(group).to(device=dev, dtype=torch.int64)
    G = g.numel()
    if G == 0:
      raise _lib.GclbError("group sizes must be >= 1 and there must be at least one group")
    group_ptr = torch.zeros(G + 1, dtype=torch.int64, device=dev)
    group_ptr[1:] = torch.cumsum(g, 0)
    index_d = torch.as_tensor(index).to(device=dev, dtype=torch.int64).contiguous()
    ff = torch.as_tensor(finest_flag).to(device=dev, dtype=torch.bool)
    pos_in_group = torch.arange(ff.numel(), device=dev) - torch.repeat_interleave(group_ptr[:-1], g)
    gid = torch.repeat_interleave(torch.arange(G, device=dev), g)
    big = torch.full((G,), 1 << 30, dtype=torch.int64, device=dev)
    big.scatter_reduce_(0, gid[ff], pos_in_group[ff], reduce="amin")          # FIRST True inside each group (:484)
    keys = torch.as_tensor(index_hash if isinstance(index_hash, torch.Tensor) else np.asarray(index_hash), dtype=torch.int64)
    keys = torch.sort(keys.to(dev)).values.contiguous()
    if not on_dev:   # host inputs: validate like the reference would fail (IndexError on an empty finest selection, :484)
      if bool((g < 1).any()):
        raise _lib.GclbError("group sizes must be >= 1 and there must be at least one group")
    return dict(group_ptr=group_ptr, index=index_d, finest_big=big, keys=keys, G=G, validated=False)

  def _run(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster, max_hn_samples, square,
           with_finest, selections=None, prepared=None):
    if not F_out.is_cuda:
      raise _lib.GclbError("GroupContrastiveLoss runs on CUDA tensors only (no CPU fallback)")
    dev = F_out.device
    N = len(F_out)
    prep = prepared if prepared is not None else self.prepare(group, index, index_hash, finest_flag, dev)
    G = prep["G"]
    if selections is None:
      # the reference's calls in the reference's order (:456-459, :506-507).  With a legacy RandomState each `choice(N, k,
      # replace=False)` permutes all N rows (2 ms at N = 120k); a numpy Generator (rng=np.random.default_rng(seed)) draws the
      # same distribution in O(k)
      if G > max_pos_cluster:
        pos_sel = self.rng.choice(G, max_pos_cluster, replace=False)
      else:
        pos_sel = np.arange(G)
      sel1 = self.rng.choice(N, min(N, max_hn_samples), replace=False)
      sel2 = self.rng.choice(N, min(N, max_hn_samples), replace=False)
    else:
      pos_sel, sel1, sel2 = selections
    sels = np.concatenate([np.asarray(pos_sel, np.int64), np.asarray(sel1, np.int64), np.asarray(sel2, np.int64)])
    sels_d = torch.from_numpy(sels).pin_memory().to(dev, non_blocking=True)      # one small H2D for the three selections
    n0, n1 = len(pos_sel), len(sel1)
    pos_sel_d, sel1_d, sel2_d = sels_d[:n0], sels_d[n0:n0 + n1], sels_d[n0 + n1:]
    finest_pos = None
    if with_finest:
      big = prep["finest_big"]
      if not prep["validated"]:
        # a selected group without a finest member: the reference raises IndexError (`feature_set[finest_flag_set][0]`, :484).
        # One host read per prepared batch (every group is checked, so later selections need no further check).
        if bool((big >= (1 << 30)).any()):
          if bool((big[pos_sel_d] >= (1 << 30)).any()):
            raise _lib.GclbError("finest_contrastive_loss: a selected group has no member with finest_flag set")
        else:
          prep["validated"] = True
      finest_pos = big.clamp(max=(1 << 30) - 1).to(torch.int32)
    args = (prep["group_ptr"], prep["index"], finest_pos, pos_sel_d, sel1_d, sel2_d, prep["keys"],
            (self.pos_thresh, self.finest_thresh, self.neg_thresh), square)
    return _GroupLossFn.apply(F_out, args)

  def finest_contrastive_loss(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster=256,
                              max_hn_samples=2048, points=None, batch_lengths=None, selections=None, prepared=None):
    return self._run(F_out, group, index, index_hash, finest_flag, max_pos_cluster, max_hn_samples,
                     self.square_loss, True, selections, prepared)

  def location_contrastive_loss(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster=256,
                                max_hn_samples=None, points=None, batch_lengths=None, selections=None, prepared=None):
    pos, _, neg = self._run(F_out, group, index, index_hash, finest_flag, max_pos_cluster, max_hn_samples, False,
                            False, selections, prepared)
    return pos, torch.zeros((), device=F_out.device), neg

  def location_circle_loss(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster=256, max_hn_samples=None,
                           points=None, batch_lengths=None, block_finest_gradient=True, use_pair_group_positive_loss=False,
                           log_scale=16, safe_radius=0.75):
    """The circle-loss head (lib/colocation_trainer.py:538-681; trainer attributes :415-420 as keyword arguments), same call
    signature and return triple.  Everything runs on F_out's device without a per-group loop: the members of the selected groups
    are one ragged batch (segment mean / segment log-sum-exp by scatter), the negative term is the dense <= 256 x 256 block the
    reference builds.  Host work: the reference's RNG calls in the reference's order (group selection, then -- only with
    use_pair_group_positive_loss -- one `choice(len, 2)` per selected group)."""
    if not F_out.is_cuda:
      raise _lib.GclbError("GroupContrastiveLoss runs on CUDA tensors only (no CPU fallback)")
    if points is None or batch_lengths is None:
      raise _lib.GclbError("location_circle_loss needs `points` and `batch_lengths` (anchor coordinates / in-batch mask)")
    dev = F_out.device
    g_host = np.asarray(group.cpu() if isinstance(group, torch.Tensor) else group, np.int64)
    G = len(g_host)
    pos_sel = np.sort(self.rng.choice(G, max_pos_cluster, replace=False)) if G > max_pos_cluster else np.arange(G)
    S = len(pos_sel)
    starts = np.concatenate([[0], np.cumsum(g_host)])
    sizes = g_host[pos_sel]
    pair = None
    if use_pair_group_positive_loss:
      pair = np.stack([self.rng.choice(int(n), 2, replace=False) for n in sizes]).astype(np.int64)      # [S, 2] positions
    # ragged member list of the selected groups: flat positions into `index`, segment id per member
    seg_h = np.repeat(np.arange(S), sizes)
    first_h = starts[pos_sel]
    flat_h = np.repeat(first_h, sizes) + (np.arange(int(sizes.sum())) - np.repeat(np.cumsum(sizes) - sizes, sizes))
    seg = torch.from_numpy(seg_h).to(dev)
    flat = torch.from_numpy(flat_h).to(dev)
    index_d = torch.as_tensor(index).to(device=dev, dtype=torch.int64)
    flag_d = torch.as_tensor(finest_flag).to(device=dev, dtype=torch.bool)
    members = index_d[flat]
    fl = flag_d[flat]
    fs = F_out[members]                                                     # [M, C]
    size_d = torch.from_numpy(sizes).to(dev)
    mean = torch.zeros((S, F_out.shape[1]), dtype=F_out.dtype, device=dev).index_add_(0, seg, fs) / size_d[:, None].to(F_out.dtype)
    square = self.square_loss

    def gap(a, b, thr):
      d2 = (a - b).pow(2).sum(-1)
      return (d2 if square else torch.sqrt(d2 + 1e-7)) - thr

    def seg_soft_lse(x, sg, n_seg):            # softplus(logsumexp_over_segment(log_scale * x * relu(x).detach())) / log_scale
      z = log_scale * x * torch.clamp(x, min=0).detach()
      m = torch.full((n_seg,), -float("inf"), dtype=z.dtype, device=dev).scatter_reduce_(0, sg, z.detach(), reduce="amax")
      e = torch.zeros((n_seg,), dtype=z.dtype, device=dev).index_add_(0, sg, torch.exp(z - m[sg]))
      return torch.nn.functional.softplus(torch.log(e) + m) / log_scale

    first_pos = torch.from_numpy(first_h).to(dev)
    if pair is not None:
      pa = index_d[first_pos + torch.from_numpy(pair[:, 0]).to(dev)]
      pb = index_d[first_pos + torch.from_numpy(pair[:, 1]).to(dev)]
      pos_loss = torch.nn.functional.softplus(gap(F_out[pa], F_out[pb], self.pos_thresh)).sum() / S
    else:
      pos_loss = seg_soft_lse(gap(mean[seg], fs, self.pos_thresh / 2), seg, S).sum() / S
    # first finest member of every selected group
    pos_in = torch.arange(len(flat_h), device=dev) - torch.repeat_interleave(torch.from_numpy(np.cumsum(sizes) - sizes).to(dev), size_d)
    big = torch.full((S,), 1 << 30, dtype=torch.int64, device=dev).scatter_reduce_(0, seg[fl], pos_in[fl], reduce="amin")
    if bool((big >= (1 << 30)).any()):
      raise _lib.GclbError("location_circle_loss: a selected group has no member with finest_flag set")
    anchor = F_out[index_d[first_pos + big]]                                # [S, C]
    if block_finest_gradient:
      keep = ~fl
      fin_rows = seg_soft_lse(gap(fs[keep], anchor.detach()[seg[keep]], self.finest_thresh), seg[keep], S)
    else:
      fin_rows = seg_soft_lse(gap(fs, anchor[seg], self.finest_thresh), seg, S)
    finest_loss = fin_rows.sum() / S
    # negatives: anchors = group means at the coordinates of each group's first member, same batch item only
    pts = torch.as_tensor(points).to(device=dev, dtype=torch.float32)
    pivot = index_d[first_pos]
    coords = pts[pivot]
    ends = torch.cumsum(torch.as_tensor(np.asarray(batch_lengths, np.float64)), 0).to(dev)
    bins = (pivot[:, None].to(torch.float64) > ends[None, :]).sum(1)
    counts = torch.zeros(len(batch_lengths), dtype=torch.int64, device=dev).index_add_(0, bins, torch.ones_like(bins))
    block = torch.searchsorted(torch.cumsum(counts, 0), torch.arange(S, device=dev), right=True)   # diagonal blocks BY COUNT (:647-652)
    same_item = block[:, None] == block[None, :]

    def sqdist(x, normalised):
      d = -2 * x @ x.T
      d = d + 2 if normalised else d + (x ** 2).sum(-1)[:, None] + (x ** 2).sum(-1)[None, :]
      return torch.clamp(d, min=1e-12)

    cd = torch.sqrt(sqdist(coords, False))
    fd = torch.sqrt(sqdist(mean, True))
    neg_mask = (cd > safe_radius) & same_item
    has_neg = neg_mask.sum(-1) > 0
    w = torch.clamp(self.neg_thresh - (fd + 1e5 * (~neg_mask).to(fd.dtype)), min=0).detach()
    rows = torch.nn.functional.softplus(torch.logsumexp(log_scale * (self.neg_thresh - fd) * w, dim=-1)) / log_scale
    return pos_loss, finest_loss, rows[has_neg].mean()

