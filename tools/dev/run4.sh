# development helper: run-wise chain walk verification, launch lists, full counters exported as CSV (the .ncu-rep stays on the box)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -15
for w in hifi clr; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r4_${w}.json 2> gpurun_out/r4_${w}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r4_${w}.json").read().strip().splitlines()[-1])
print("RES $w", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"], d["work"])
PY
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r4_fullparity_${w}.json 2> gpurun_out/r4_fullparity_${w}.err; tail -c 300 gpurun_out/r4_fullparity_${w}.json
done
python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_plain_hifi.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_hifi.csv python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_ncu_hifi.log 2>&1
python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_plain_clr.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_launches_clr.csv python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r4_ncu_clr.log 2>&1
ncu --set full --clock-control none -k regex:'chainRunDpKernel|chainFillKernel|chainRunsKernel|rebuildHitsKernel|Onesweep|expandKernel|chainWalkKernel|wfaKernel' -c 14 -o /tmp/r4_full_hifi python bench.py --workload hifi --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r4_ncufull_hifi.log 2>&1
ncu -i /tmp/r4_full_hifi.ncu-rep --page raw --csv > gpurun_out/r4_full_hifi_raw.csv 2> gpurun_out/r4_full_export.err
ls -la gpurun_out /tmp/r4_full_hifi.ncu-rep
du -sh gpurun_out
