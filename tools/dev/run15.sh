# final check of the committed tree: build entry + smoke, whole GPU suite, default bench line
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py > gpurun_out/r15_bench_default.json 2> gpurun_out/r15_bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r15_bench_default.json").read().strip().splitlines()[-1])
print("RES default", round(d["ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), d["gpu_launches"], d["clocks"], d["roofline"]["kernel"], d["cpu_baseline"]["value"])
PY
