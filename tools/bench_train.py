"""GCL training step on the CUDA operators (BASELINE config 4), timing only -- parity of every piece is covered by
tests/test_gpu_parity.py (conv fwd/dgrad/wgrad + train-mode BN vs oracle autograd; fused loss fwd/bwd vs the reference's
trainer).  One step = ResUNetBN2C forward in train mode on a batch of 4 samples x 3 colocated scans (12 clouds), the
fused finest-contrastive loss (positive groups of 3 + hardest negatives), backward, SGD(momentum .8, wd 1e-4, lr .1)
(lib/colocation_trainer.py:811-916, config.py:87-96).  Groups are synthetic: the three scans of a sample are jittered
copies of one 32-beam scan, a group = the rows of the three clouds that share a voxel coordinate.

  python tools/bench_train.py [--steps 10] [--samples 4] [--tf32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/bench_train.py --tf32        # N ranks, each its own colocated-scan groups, NCCL gradient all-reduce

Multi-GPU: samples (colocated scan groups) are sharded by rank; the only collective is the bucketed gradient all-reduce
(gcl_b200/sharding.py:FlatGradients -- one flat buffer, one NCCL AVG all-reduce) between backward and the SGD step.  BatchNorm statistics stay per rank, as with
torch DDP without SyncBatchNorm (the reference trains on one GPU: no DistributedDataParallel anywhere in /lib).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--steps", type=int, default=10)
  ap.add_argument("--samples", type=int, default=4)
  ap.add_argument("--voxel", type=float, default=0.3)
  ap.add_argument("--tf32", action="store_true", help="forward, dgrad and wgrad on the tcgen05 kind::tf32 kernels")
  args = ap.parse_args()
  import gcl_b200
  from gcl_b200 import MinkowskiEngine as ME, ops, synth
  from gcl_b200.loss import GroupContrastiveLoss, _exhaustive_hash
  import torch.distributed as dist
  from gcl_b200.sharding import FlatGradients, gather_counts
  rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
  local = int(os.environ.get("LOCAL_RANK", 0))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  if args.tf32:
    ME.set_training_conv_algo("tf32")
  torch.manual_seed(0)
  rng = np.random.RandomState(rank)
  clouds = []
  for s in range(args.samples):
    base = synth.cast(synth.Scene(40 + s + 100 * rank), synth.NUSCENES, seed=s + 100 * rank)
    for j in range(3):
      clouds.append((base + rng.normal(0, 0.01, base.shape)).astype(np.float32))
  xyz = torch.from_numpy(np.concatenate(clouds)).to(dev)
  ptr = torch.tensor(np.cumsum([0] + [len(c) for c in clouds]))
  cm, _ = ops.voxelize(xyz, args.voxel, ptr)
  N = cm.n
  C = cm.coords
  # groups: voxel coordinate present in all three clouds of a sample
  idx_rows, sizes, flags = [], [], []
  for s in range(args.samples):
    c0 = C[C[:, 0] == 3 * s]
    rows0 = torch.nonzero(C[:, 0] == 3 * s)[:, 0]
    q1, q2 = c0.clone(), c0.clone()
    q1[:, 0], q2[:, 0] = 3 * s + 1, 3 * s + 2
    r1, r2 = ops.hash_query(cm, q1), ops.hash_query(cm, q2)
    ok = (r1 >= 0) & (r2 >= 0)
    g = torch.stack([rows0[ok], r1[ok].long(), r2[ok].long()], 1)
    idx_rows.append(g.reshape(-1))
    sizes.append(torch.full((g.shape[0],), 3, dtype=torch.int64))
    f = torch.zeros_like(g, dtype=torch.bool); f[:, 0] = True
    flags.append(f.reshape(-1))
  index = torch.cat(idx_rows).cpu(); group = torch.cat(sizes); finest = torch.cat(flags).cpu()
  split = index.numpy().reshape(-1, 3)
  index_hash = _exhaustive_hash(list(split), N)
  print(f"clouds={len(clouds)} voxels={N} groups={len(group)}", file=sys.stderr)

  torch.manual_seed(0)  # identical initial weights on every rank
  model = gcl_b200.load_model("ResUNetBN2C")(1, 32, bn_momentum=0.05, conv1_kernel_size=5, normalize_feature=True).to(dev)
  model.train()
  opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.8, weight_decay=1e-4)
  crit = GroupContrastiveLoss(pos_thresh=0.1, neg_thresh=1.4, finest_thresh=0.2, square_loss=True,
                              rng=np.random.RandomState(0))
  feats = torch.ones(N, 1, device=dev)

  grads = FlatGradients(model.parameters())

  def step():
    grads.zero()
    st = ME.SparseTensor(feats, coordinates=C)
    F = model(st).F
    pos, fin, neg = crit.finest_contrastive_loss(F, group, index, index_hash, finest, max_pos_cluster=256 * args.samples,
                                                 max_hn_samples=256 * args.samples)
    loss = pos + fin + neg
    loss.backward()
    grads.allreduce()
    opt.step()
    return loss

  l0 = step()
  # warm up until PyTorch's caching allocator stops growing: a cudaMalloc inside the timed steps is a device-wide
  # synchronisation of unpredictable length (measured: 16 vs 27 ms per step on the same box)
  seg = lambda: torch.cuda.memory_stats(dev).get("segment.all.allocated", 0)
  quiet = 0
  for _ in range(12):
    s0 = seg()
    step()
    quiet = quiet + 1 if seg() == s0 else 0
    if quiet >= 3:
      break
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  prof = os.environ.get("GCLB_PROFILE_LAST_STEP") == "1"      # tools/profile_step.py --train: ncu range = the last step
  for i in range(args.steps):
    if prof and i == args.steps - 1:
      torch.cuda.synchronize()
      torch.cuda.profiler.start()
    l = step()
  if prof:
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / args.steps
  if world > 1:
    tot_clouds, ms = gather_counts(float(len(clouds)), ms)
    # every rank starts from the same seeded weights and applies the same averaged gradient: weights must stay identical
    w = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    lo, hi = w.clone(), w.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks diverged after the gradient all-reduce"
  else:
    tot_clouds = len(clouds)
  if rank == 0:
    print(json.dumps({"metric": "gcl_train_step_ms", "value": round(ms, 2), "unit": "ms/step", "n_gpus": world,
                    "scans_per_s": round(tot_clouds / ms * 1e3, 1), "clouds_per_rank": len(clouds), "voxels_rank0": N,
                    "groups_rank0": int(len(group)), "loss_first": round(float(l0), 4), "loss_last": round(float(l), 4),
                    "conv_kernels": ("tcgen05 kind::tf32 fwd + dgrad + wgrad" if args.tf32 else
                                     "exact-fp32 CUDA-core fwd/dgrad/wgrad")}))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
