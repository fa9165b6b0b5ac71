"""Multi-GPU plumbing (one process per GPU, torch.distributed).

Inference: scan pairs are independent -> pair p goes to rank p mod world, weights replicated, NO data-path collective
(SURVEY.md section 8e).  Training: shard by colocated scan group, then one all-reduce of the gradients
(NCCL over NVLink on B200; gloo in the CPU tests) -- the only exchange step on the path.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def shard_pairs(n_pairs: int, rank: int, world: int) -> List[int]:
  """round-robin assignment of pair indices to ranks"""
  if not (0 <= rank < world):
    raise ValueError("rank out of range")
  return list(range(rank, n_pairs, world))


def gather_counts(local_units: float, local_ms: float, group=None):
  """whole-job aggregate for throughput reporting: (sum of units over ranks, max of elapsed ms over ranks)"""
  if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
    return float(local_units), float(local_ms)
  dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
  s = torch.tensor([local_units], dtype=torch.float64, device=dev)
  m = torch.tensor([local_ms], dtype=torch.float64, device=dev)
  dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
  dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
  return s.item(), m.item()


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None, bucket_bytes: int = 32 << 20):
  """Average gradients over ranks with flat buckets (8.75 M fp32 grads of ResUNetBN2C = 35 MB -> 2 buckets): sized for
  launch latency and overlap, not link count (NVSwitch gives every peer full bandwidth)."""
  if not (dist.is_available() and dist.is_initialized()):
    return
  world = dist.get_world_size(group)
  if world == 1:
    return
  grads = [p.grad for p in params if p.grad is not None]
  bucket, size = [], 0

  def flush():
    nonlocal bucket, size
    if not bucket:
      return
    flat = torch.cat([g.reshape(-1) for g in bucket])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for g in bucket:
      g.copy_(flat[off:off + g.numel()].view_as(g))
      off += g.numel()
    bucket, size = [], 0

  for g in grads:
    bucket.append(g)
    size += g.numel() * g.element_size()
    if size >= bucket_bytes:
      flush()
  flush()


class FlatGradients:
  """Gradient storage for the data-parallel training step: ONE flat fp32 buffer, every `p.grad` a view into it, so the
  exchange step is a single in-place NCCL all-reduce (AVG) with no pack/unpack copies (35 MB for ResUNetBN2C, one launch
  over NVSwitch).  autograd accumulates into an existing `.grad` in place, so the views survive `backward()`; call
  `zero()` instead of `optimizer.zero_grad(set_to_none=True)`."""

  def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
    self.params = [p for p in params if p.requires_grad]
    self.group = group
    total = sum(p.numel() for p in self.params)
    ref = self.params[0]
    self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
    off = 0
    for p in self.params:
      p.grad = self.flat[off:off + p.numel()].view_as(p)
      off += p.numel()

  def zero(self):
    self.flat.zero_()

  def allreduce(self):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
      return
    if dist.get_backend(self.group) == "nccl":
      dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)
    else:  # gloo has no AVG
      dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
      self.flat.div_(dist.get_world_size(self.group))


class PackedGradients:
  """Gradient exchange of the data-parallel training step without per-parameter kernels.  `zero()` drops every `.grad`, so
  autograd hands its gradient tensors over as they are (with a pre-existing `.grad` -- FlatGradients' views -- it launches one
  add per parameter: 83 launches per step for ResUNetBN2C, ~1 ms of host time in a host-bound step).  `allreduce()` packs the
  gradients into one flat buffer (one cat), runs ONE all-reduce (AVG) and copies the result back with one multi-tensor copy;
  with a single rank it does nothing at all."""

  def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
    self.params = [p for p in params if p.requires_grad]
    self.group = group

  def zero(self):
    for p in self.params:
      p.grad = None

  def allreduce(self):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
      return
    for p in self.params:          # every rank must exchange the same buffer: an unused parameter contributes zeros
      if p.grad is None:
        p.grad = torch.zeros_like(p)
    grads = [p.grad for p in self.params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    if dist.get_backend(self.group) == "nccl":
      dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group)
    else:  # gloo has no AVG
      dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
      flat.div_(dist.get_world_size(self.group))
    torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])

