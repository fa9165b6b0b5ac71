"""Summarise an ncu report (all library kernels of one bench step) into a markdown table, one row per kernel.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xxx.md"""
import collections
import csv
import subprocess
import sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr, rows = r[0], r[2:]
col = {h: i for i, h in enumerate(hdr)}


def g(row, k, default=0.0):
  try:
    return float(row[col[k]].replace(",", ""))
  except Exception:
    return default


UNITS = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
unit_of = lambda k: r[1][col[k]].lower().split("/")[0] if k in col else "byte"
SC_RD = UNITS.get(unit_of("dram__bytes_read.sum"), 1.0)
SC_BW = UNITS.get(unit_of("dram__bytes.sum.per_second"), 1.0)
agg = collections.OrderedDict()
for row in rows:
  name = row[col["Kernel Name"]]
  name = name.replace("void ", "").replace("gclb::", "").split("(")[0]
  a = agg.setdefault(name, dict(n=0, t=0.0, rd=0.0, wr=0.0, dram=0.0, l2=0.0, l1=0.0, tens=0.0, sm=0.0, regs=0, occ=0.0))
  t = g(row, "gpu__time_duration.sum")            # us
  a["n"] += 1; a["t"] += t
  if "dram__bytes_read.sum" in col:
    a["rd"] += g(row, "dram__bytes_read.sum") * SC_RD; a["wr"] += g(row, "dram__bytes_write.sum") * SC_RD   # bytes
  else:   # lighter section sets only carry the achieved DRAM rate
    a["rd"] += g(row, "dram__bytes.sum.per_second") * SC_BW * t * 1e-6
  for key, m in (("dram", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("l2", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("l1", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("tens", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                 ("sm", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                 ("occ", "sm__warps_active.avg.pct_of_peak_sustained_active")):
    a[key] += g(row, m) * t                      # time-weighted
  a["regs"] = int(g(row, "launch__registers_per_thread"))
tot = sum(a["t"] for a in agg.values())
print(f"# {title}\n")
print("One bench step (16 KITTI-shape scan pairs, ~650k voxels) under `ncu` (serialised, cold caches: compare shares, not "
      "absolutes). Percentages are time-weighted means over the launches of a kernel; DRAM = read + write bytes.\n")
print("| kernel | launches | time us | share | DRAM MB | DRAM GB/s | DRAM % | L2 % | tensor pipe % | SM % | warps active % | regs |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
  t = a["t"]
  gbs = (a["rd"] + a["wr"]) / (t * 1e-6) / 1e9 if t else 0
  w = lambda k: a[k] / t if t else 0
  print(f"| `{name}` | {a['n']} | {t:.0f} | {100 * t / tot:.1f}% | {(a['rd'] + a['wr']) / 1e6:.1f} | "
        f"{gbs:.0f} | {w('dram'):.1f} | {w('l2'):.1f} | {w('tens'):.1f} | {w('sm'):.1f} | {w('occ'):.1f} | {a['regs']} |")
print(f"\nTotal {tot / 1e3:.2f} ms over {sum(a['n'] for a in agg.values())} launches.")
