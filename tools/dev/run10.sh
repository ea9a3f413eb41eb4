# development helper: threshold-bounded edit distance (per-query max divergence)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge_cases.py tests/test_gpu_host_mirror.py -m gpu -x -q 2>&1 | tail -15
for w in hifi; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r10_${w}.json 2> gpurun_out/r10_${w}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r10_${w}.json").read().strip().splitlines()[-1])
print("RES $w", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"], d["work"], d["api_wall_ms"], d["e2e"])
PY
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r10_fullparity_${w}.json 2> gpurun_out/r10_fullparity_${w}.err; tail -c 300 gpurun_out/r10_fullparity_${w}.json
done
