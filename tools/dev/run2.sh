# development helper (one gpurun call): run-compressed DP + sort skips verification
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -15
for w in hifi clr; do
 for v in "2" "1"; do
  FG_DP_MODE=$v timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_${w}_dp$v.json 2> gpurun_out/r2_${w}_dp$v.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_${w}_dp$v.json").read().strip().splitlines()[-1])
print("RES $w dpmode=$v", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"], d["work"])
PY
 done
done
for w in hifi clr; do
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r2_fullparity_${w}_rundp.json 2> gpurun_out/r2_fullparity_${w}_rundp.err; tail -c 700 gpurun_out/r2_fullparity_${w}_rundp.json
done
