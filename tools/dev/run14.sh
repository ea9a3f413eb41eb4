# development helper: band-pruned wavefronts
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py --workload hifi --steps 3 --warmup 3 > gpurun_out/r14_bench_hifi.json 2> gpurun_out/r14_bench_hifi.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r14_bench_hifi.json").read().strip().splitlines()[-1])
print("RES hifi", round(d["ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), d["phases_ms"], d["api_wall_ms"], d["work"]["overlaps"], d.get("cpu_baseline",{}).get("value"))
PY
timeout 900 python tools/full_scale_parity.py hifi > gpurun_out/r14_fullparity_hifi.json 2> gpurun_out/r14_fullparity_hifi.err; tail -c 200 gpurun_out/r14_fullparity_hifi.json
