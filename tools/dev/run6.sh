# development helper: segmented radix sort of packed hits + presorted flag verification
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -15
for w in hifi clr; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6_${w}.json 2> gpurun_out/r6_${w}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r6_${w}.json").read().strip().splitlines()[-1])
print("RES $w", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"], d["work"], d["api_wall_ms"])
PY
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r6_fullparity_${w}.json 2> gpurun_out/r6_fullparity_${w}.err; tail -c 300 gpurun_out/r6_fullparity_${w}.json
done
FG_HIT_RADIX=2 timeout 600 python bench.py --workload hifi --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r6_hifi_cub.json 2> gpurun_out/r6_hifi_cub.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r6_hifi_cub.json").read().strip().splitlines()[-1])
print("RES hifi cub", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"])
PY
