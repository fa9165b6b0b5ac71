set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -k "test_tc_" -x -q 2>&1 | tail -6
ncu --profile-from-start off --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:spconv_fwd -o /tmp/conv_traffic python tools/profile_step.py > /tmp/ct.log 2>&1; tail -2 /tmp/ct.log
python tools/make_conv_traffic.py /tmp/conv_traffic.ncu-rep > gpurun_out/conv_traffic.json; head -5 gpurun_out/conv_traffic.json
ncu --profile-from-start off --clock-control none --section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section Occupancy --section LaunchStats --metrics dram__bytes_read.sum,dram__bytes_write.sum -o /tmp/step python tools/profile_step.py --register > /tmp/st.log 2>&1; tail -2 /tmp/st.log
python tools/ncu_summary.py /tmp/step.ncu-rep "Round 2 final build: every libgclb200 kernel of one 16-pair step incl. SC2-PCR registration (ncu SpeedOfLight + memory/compute workload sections)" > gpurun_out/r02_fin2_step_all_kernels.md; head -12 gpurun_out/r02_fin2_step_all_kernels.md | cut -c1-160
ncu --profile-from-start off --clock-control none --set full --import-source on -k regex:spconv_fwd_tc_kernel -s 19 -c 1 -o /tmp/full python tools/profile_step.py > /tmp/fl.log 2>&1; tail -2 /tmp/fl.log
ncu -i /tmp/full.ncu-rep --page raw --csv > gpurun_out/r02_fin2_conv_full_raw.csv; wc -c gpurun_out/r02_fin2_conv_full_raw.csv
ncu -i /tmp/full.ncu-rep --page details --csv > gpurun_out/r02_fin2_conv_full_details.csv 2>/dev/null; wc -c gpurun_out/r02_fin2_conv_full_details.csv
python bench.py > gpurun_out/r02_fin2_pairs.json 2> gpurun_out/r02_fin2_pairs.err; tail -c 600 gpurun_out/r02_fin2_pairs.json
python bench.py --workload nuscenes --no-cpu-baseline > gpurun_out/r02_fin2_nuscenes.json 2>/dev/null
python bench.py --workload sweep --no-cpu-baseline > gpurun_out/r02_fin2_sweep.json 2>/dev/null
python bench.py --workload train > gpurun_out/r02_fin2_train.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_fin2_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_fin2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /tmp/b.log 2>&1
for f in pairs nuscenes sweep train reference; do python - <<PY
import json
d=json.loads(open("gpurun_out/r02_fin2_$f.json").read().strip().splitlines()[-1])
print("$f", d.get("value"), d.get("unit"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("roofline") or {}).get("traffic"), d.get("e2e_from_disk"), d.get("registered"))
PY
done
