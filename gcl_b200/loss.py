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

  def _run(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster, max_hn_samples, square,
           with_finest, selections=None):
    if not F_out.is_cuda:
      raise _lib.GclbError("GroupContrastiveLoss runs on CUDA tensors only (no CPU fallback)")
    dev = F_out.device
    N = len(F_out)
    group_h = torch.as_tensor(group).cpu().to(torch.int64)
    G = len(group_h)
    if selections is None:
      if G > max_pos_cluster:
        pos_sel = self.rng.choice(G, max_pos_cluster, replace=False)
      else:
        pos_sel = np.arange(G)
      sel1 = self.rng.choice(N, min(N, max_hn_samples), replace=False)
      sel2 = self.rng.choice(N, min(N, max_hn_samples), replace=False)
    else:
      pos_sel, sel1, sel2 = selections
    if G == 0 or bool((group_h < 1).any()):
      raise _lib.GclbError("group sizes must be >= 1 and there must be at least one group")
    group_ptr = torch.zeros(G + 1, dtype=torch.int64)
    group_ptr[1:] = torch.cumsum(group_h, 0)
    index_d = torch.as_tensor(index).to(device=dev, dtype=torch.int64).contiguous()
    finest_pos = None
    if with_finest:
      ff = torch.as_tensor(finest_flag).cpu().to(torch.bool)
      # position of the FIRST True inside each group (reference: feature_set[finest_flag_set][0])
      pos_in_group = torch.arange(len(ff)) - torch.repeat_interleave(group_ptr[:-1], group_h)
      big = torch.full((G,), 1 << 30, dtype=torch.int64)
      gid = torch.repeat_interleave(torch.arange(G), group_h)
      big.scatter_reduce_(0, gid[ff], pos_in_group[ff], reduce="amin")
      sel_groups = torch.as_tensor(np.asarray(pos_sel), dtype=torch.int64)
      if bool((big[sel_groups] >= (1 << 30)).any()):
        # the reference raises IndexError here (`feature_set[finest_flag_set][0]` on an empty selection, :484)
        raise _lib.GclbError("finest_contrastive_loss: a selected group has no member with finest_flag set")
      finest_pos = big.clamp_(max=(1 << 30) - 1).to(torch.int32).to(dev)
    keys = torch.sort(torch.as_tensor(np.asarray(index_hash), dtype=torch.int64).to(dev)).values.contiguous()
    to_d = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.int64).to(dev).contiguous()
    args = (group_ptr.to(dev), index_d, finest_pos, to_d(pos_sel), to_d(sel1), to_d(sel2), keys,
            (self.pos_thresh, self.finest_thresh, self.neg_thresh), square)
    return _GroupLossFn.apply(F_out, args)

  def finest_contrastive_loss(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster=256,
                              max_hn_samples=2048, points=None, batch_lengths=None, selections=None):
    return self._run(F_out, group, index, index_hash, finest_flag, max_pos_cluster, max_hn_samples,
                     self.square_loss, True, selections)

  def location_contrastive_loss(self, F_out, group, index, index_hash, finest_flag, max_pos_cluster=256,
                                max_hn_samples=None, points=None, batch_lengths=None, selections=None):
    pos, _, neg = self._run(F_out, group, index, index_hash, finest_flag, max_pos_cluster, max_hn_samples, False,
                            False, selections)
    return pos, torch.zeros((), device=F_out.device), neg
