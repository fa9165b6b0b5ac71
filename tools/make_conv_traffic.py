"""profiles/conv_traffic.json from an ncu report of the CURRENT build (bench.py refuses the file when the build id differs):

  gpurun -- 'ncu --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum \\
             -k regex:spconv_fwd -o gpurun_out/conv_traffic python tools/profile_step.py'
  python tools/make_conv_traffic.py gpurun_out/conv_traffic.ncu-rep
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gcl_b200 import build  # noqa: E402

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, units, rows = r[0], r[1], r[2:]
col = {h: i for i, h in enumerate(hdr)}
UN = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}
tot, n, per = 0.0, 0, {}
for row in rows:
  name = row[col["Kernel Name"]]
  if "spconv_fwd_tc_kernel" not in name and "spconv_fwd_halo_kernel" not in name:
    continue
  b = 0.0
  for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    b += float(row[col[k]].replace(",", "")) * UN[units[col[k]].lower()]
  tot += b; n += 1
  key = name.split("(")[0].replace("void gclb::", "")
  per.setdefault(key, [0, 0.0]); per[key][0] += 1; per[key][1] += b
out = {"build_id": build.source_id(), "dram_bytes_per_launch": int(tot / max(n, 1)), "dram_bytes_per_step": int(tot), "launches": n,
       "per_kernel": {k: {"launches": v[0], "dram_bytes": int(v[1])} for k, v in per.items()},
       "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of {os.path.basename(rep)} (tools/profile_step.py, one 16-pair step, "
                 f"serialised + cold L2), build {build.source_id()}"}
json.dump(out, open(os.path.join(ROOT, "profiles", "conv_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
