# development helper: whole GPU suite + benches + full-scale parity of the current tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for w in hifi clr; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r9_${w}.json 2> gpurun_out/r9_${w}.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r9_${w}.json").read().strip().splitlines()[-1])
print("RES $w", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"], d["work"], d["api_wall_ms"])
PY
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r9_fullparity_${w}.json 2> gpurun_out/r9_fullparity_${w}.err; tail -c 300 gpurun_out/r9_fullparity_${w}.json
done
