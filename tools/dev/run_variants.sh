# development helper: A/B runs of kernel variants selected by environment variables (one gpurun call)
set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -15
for w in hifi clr; do
 for v in "0 512 1" "1 512 1" "1 384 1" "1 384 3"; do
  set -- $v
  FG_DP_PRUNE=$1 FG_SORT_SMALL=$2 FG_SORT_VAR=$3 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/v_${w}_$1_$2_$3.json 2> gpurun_out/v_${w}_$1_$2_$3.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/v_${w}_$1_$2_$3.json").read().strip().splitlines()[-1])
ph=d.get("phases_ms",{})
print("RES $w prune=$1 pairsSmall=$2 pairsVar=$3", round(d["ms_per_step"],1), {k:ph.get(k) for k in ("hit_sort_small","chain_extsort_small","chain_ordsort_small","hit_sort_top","chain_extsort_top","chain_ordsort_top","chain_dp","edit")}, d["work"])
PY
 done
done
