import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gcl_b200 import _lib
lib = _lib.load()
out = torch.zeros(8, dtype=torch.int64, device="cuda")
for n in (32, 64, 128, 256):
  for issuers in (1, 2, 4):
    if issuers * n > 512: continue
    for per in (0, 1, 16, 17, 18):   # +16: warp-uniform issue loop under elect.sync; 2/18: commit + wait per 4 MMAs
      for _ in range(2):
        out.zero_()
        _lib.call("gclb_debug_umma_rate", n, 4000, per, 1, issuers, out.data_ptr(), None)
        torch.cuda.synchronize()
      t = out.tolist()
      print(f"N={n:3d} issuers={issuers} commit_mode={per}: per-issuer cycles per MMA {[round(t[2*i]/4000,1) for i in range(issuers)]} => aggregate {max(t[0::2]) / (4000 * issuers):.1f} cycles per MMA")
