# final check after the bandLo/bandHi fix: whole GPU suite, default bench line, full-scale HiFi parity
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r18_bench_hifi.json 2> gpurun_out/r18_bench_hifi.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r18_bench_hifi.json").read().strip().splitlines()[-1])
print("RES hifi", round(d["ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), d["phases_ms"].get("edit"), d["work"]["overlaps"])
PY
timeout 200 python tools/full_scale_parity.py hifi > gpurun_out/r18_fullparity_hifi.json 2> gpurun_out/r18_fullparity_hifi.err; tail -c 200 gpurun_out/r18_fullparity_hifi.json
