"""CPU, world_size 2 over gloo: the N>1 host logic (pair sharding, throughput aggregation, gradient all-reduce)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gcl_b200.sharding import allreduce_gradients, gather_counts, shard_pairs


def test_shard_pairs_partition():
  for world in (1, 2, 4, 8):
    allp = sorted(sum((shard_pairs(37, r, world) for r in range(world)), []))
    assert allp == list(range(37))
  assert shard_pairs(5, 1, 2) == [1, 3]


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  units, ms = gather_counts(local_units=10.0 * (rank + 1), local_ms=5.0 + rank)
  torch.manual_seed(0)
  lin = torch.nn.Linear(300, 200)          # same init on both ranks
  x = torch.full((4, 300), float(rank + 1))
  lin(x).sum().backward()
  local = [p.grad.clone() for p in lin.parameters()]
  allreduce_gradients(lin.parameters(), bucket_bytes=1 << 16)   # forces several buckets
  # same exchange through the flat-view buffer the training step uses: views must survive backward(), result identical
  from gcl_b200.sharding import FlatGradients
  lin2 = torch.nn.Linear(300, 200)
  lin2.load_state_dict(lin.state_dict())
  fg = FlatGradients(lin2.parameters())
  for _ in range(2):  # second pass checks zero() + in-place accumulation
    fg.zero()
    lin2(x).sum().backward()
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() and p.grad.data_ptr() < fg.flat.data_ptr() + fg.flat.numel() * 4
               for p in lin2.parameters())
    fg.allreduce()
  for a, b in zip(lin.parameters(), lin2.parameters()):
    assert torch.allclose(a.grad, b.grad)
  # and through the packed exchange the training step uses now (grads dropped, one cat + one all-reduce + one multi-copy)
  from gcl_b200.sharding import PackedGradients
  lin3 = torch.nn.Linear(300, 200)
  lin3.load_state_dict(lin.state_dict())
  pg = PackedGradients(lin3.parameters())
  for _ in range(2):
    pg.zero()
    assert all(p.grad is None for p in lin3.parameters())
    lin3(x).sum().backward()
    pg.allreduce()
  for a, b in zip(lin.parameters(), lin3.parameters()):
    assert torch.allclose(a.grad, b.grad)
  q.put((rank, units, ms, [p.grad.clone() for p in lin.parameters()], local))
  dist.barrier()
  dist.destroy_process_group()


def test_two_rank_gloo():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  import socket
  with socket.socket() as sk:          # a port that is free right now (a fixed formula can collide with a lingering rendezvous)
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  (_, u0, m0, g0, l0), (_, u1, m1, g1, l1) = res
  assert u0 == u1 == 30.0 and m0 == m1 == 6.0                       # sum of units, max of times
  for a, b, x, y in zip(g0, g1, l0, l1):
    assert torch.allclose(a, b) and torch.allclose(a, (x + y) / 2)  # averaged, identical on both ranks
