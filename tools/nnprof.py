"""time K4 (mutual NN of 16 pairs x 5000 x 5000 x 32) on both kernels"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcl_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
P, N = 16, 5000
A = torch.nn.functional.normalize(torch.randn(P * N, 32, device=dev), dim=1)
B = torch.nn.functional.normalize(torch.randn(P * N, 32, device=dev), dim=1)
ptr = torch.arange(P + 1, device=dev, dtype=torch.int64) * N
def t(fn, reps=10):
  fn(); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps): fn()
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / reps * 1e3
for algo in (2, 1):
  us = t(lambda: ops.nn_search(A, B, ptr, ptr, both=True, algo=algo, max_n=N, max_m=N))
  print(f"algo {algo}: {us:.1f} us per 16-pair mutual NN ({2 * P * N * N * 32 * 2 / us / 1e6:.1f} useful TFLOP/s)")
r2 = ops.nn_search(A, B, ptr, ptr, both=True, algo=2, max_n=N, max_m=N)
r1 = ops.nn_search(A, B, ptr, ptr, both=True, algo=1, max_n=N, max_m=N)
print("index agreement tc vs fp32:", (r2[0] == r1[0]).float().mean().item(), (r2[2] == r1[2]).float().mean().item(),
      "max |d| diff", (r2[1] - r1[1]).abs().max().item())
