# development helper (one gpurun call): radix fast path + pruned DP verification
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -5
FG_DP_PRUNE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -5
for w in hifi clr; do
 for v in "1 0" "0 0" "1 1"; do
  set -- $v
  FG_HIT_RADIX=$1 FG_DP_PRUNE=$2 timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_${w}_radix$1_prune$2.json 2> gpurun_out/r1_${w}_radix$1_prune$2.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r1_${w}_radix$1_prune$2.json").read().strip().splitlines()[-1])
print("RES $w radix=$1 prune=$2", round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1), d["phases_ms"], d["work"])
PY
 done
done
for w in hifi clr; do
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r1_fullparity_${w}_radix.json 2> gpurun_out/r1_fullparity_${w}_radix.err; tail -c 600 gpurun_out/r1_fullparity_${w}_radix.json
  FG_DP_PRUNE=1 timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r1_fullparity_${w}_radix_prune.json 2> gpurun_out/r1_fullparity_${w}_radix_prune.err; tail -c 600 gpurun_out/r1_fullparity_${w}_radix_prune.json
done
