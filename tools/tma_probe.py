import ctypes, sys, torch
sys.path.insert(0, ".")
from gcl_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
n, c = 1000, 64
X = (torch.arange(n * c, dtype=torch.float32, device=dev).reshape(n, c))
for box_rows in (1, 4):
  for rows in ([5, 17, 900, 3], [5, -1, 2000, 3]):
    out = torch.full((256,), -1.0, device=dev)
    r = (ctypes.c_int32 * 4)(*rows)
    rc = lib.gclb_debug_tma_gather4(X.data_ptr(), n, c, box_rows, 32, ctypes.cast(r, ctypes.c_void_p), out.data_ptr(), None)
    try:
      torch.cuda.synchronize()
      o = out.cpu().reshape(8, 32)
      print("box_rows", box_rows, "rows", rows, "rc", rc)
      for i in range(4):
        print("  smem row", i, "chunk starts:", [int(o[i, 4 * j]) for j in range(8)])
      print("  rows 4-7 first:", [float(o[i, 0]) for i in range(4, 8)])
    except Exception as e:
      print("box_rows", box_rows, "rows", rows, "rc", rc, "ERROR", str(e)[:200]); sys.exit(0)
