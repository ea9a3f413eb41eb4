# development helper: fg_overlaps_refilter + extended variant tests, bench lines of both workloads
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6
for w in clr hifi; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r13_bench_${w}.json 2> gpurun_out/r13_bench_${w}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r13_bench_${w}.json").read().strip().splitlines()[-1])
print("RES $w", round(d["ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), d["phases_ms"], d["api_wall_ms"], d["work"]["overlaps"], d.get("cpu_baseline",{}).get("value"))
PY
done
