"""GPU-side kernel breakdown of one GCL training step (torch.profiler / CUPTI): which kernels carry the step.
usage: python tools/train_kernels.py [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from gcl_b200.training import GclTrainStep
dev = torch.device("cuda:0")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ts = GclTrainStep(dev, samples=4)
for _ in range(6): ts.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(steps): ts.step()
  torch.cuda.synchronize()
rows = [(e.key, e.count, getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0.0)) for e in prof.key_averages()]
rows = [r for r in rows if r[2] > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"GPU time per step {tot / steps / 1e3:.2f} ms over {sum(r[1] for r in rows) / steps:.0f} launches")
for k, c, t in rows[:28]:
  print(f"{t / steps:9.1f} us  {100 * t / tot:5.1f}%  x{c / steps:5.1f}  {k[:110]}")
