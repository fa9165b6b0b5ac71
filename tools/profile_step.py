"""One hot-path step inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`.

  ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis \\
      --section ComputeWorkloadAnalysis --section Occupancy --section LaunchStats \\
      --metrics dram__bytes_read.sum,dram__bytes_write.sum -o gpurun_out/step python tools/profile_step.py [--train]

default: one 16-pair inference step (voxelise + maps + ResUNetBN2C fwd x 32 clouds + subsample + mutual NN) through
PairMatcher.match, the same call bench.py times.  --train: one GCL training step (tools/bench_train.py's step, tf32 mode)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--train", action="store_true")
  ap.add_argument("--pairs", type=int, default=16)
  ap.add_argument("--register", action="store_true", help="also run SC2-PCR registration of every pair inside the profiled step")
  args = ap.parse_args()
  dev = torch.device("cuda:0")
  if args.train:     # one GCL training step (bench.py --workload train's step) inside the profiled range
    from gcl_b200.training import GclTrainStep
    ts = GclTrainStep(dev, samples=4)
    for _ in range(6):
      ts.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ts.step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    return
  import bench
  from gcl_b200 import MinkowskiEngine as ME
  from gcl_b200.pipeline import PairMatcher
  matcher = PairMatcher(bench.seeded_model(ME), voxel=bench.VOXEL, subsample=bench.SUBSAMPLE, device=dev, seed=0,
                        register=args.register)
  batches = [(x.to(dev), p) for x, p in bench.make_batches(2, args.pairs, seed=0)]
  for i in range(3):
    matcher.match(*batches[i % 2])
  torch.cuda.synchronize()
  torch.cuda.profiler.start()
  matcher.match(*batches[1])
  torch.cuda.synchronize()
  torch.cuda.profiler.stop()


if __name__ == "__main__":
  main()
