# round-2 experiment 1: device epilogue + deep prefetch in the segmented sort, A/B on configs[0] and configs[1]
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/exp1_env.txt; nproc >> gpurun_out/exp1_env.txt
timeout 600 python -m pytest tests/test_gpu_edge_cases.py tests/test_gpu_golden.py "tests/test_gpu_parity.py::test_kernel_variants_agree_with_oracle" tests/test_gpu_parity.py::test_clr_small_full_parity tests/test_gpu_parity.py::test_hifi_small_full_parity -x -q -m gpu > gpurun_out/exp1_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp1_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
for wl in clr hifi; do
  timeout 300 $B --workload $wl > gpurun_out/exp1_${wl}_default.json 2> gpurun_out/exp1_${wl}_default.err; echo "rc=$?" >> gpurun_out/exp1_${wl}_default.err
  FG_SRS_DEEP=0 timeout 300 $B --workload $wl --no-e2e > gpurun_out/exp1_${wl}_shallow.json 2> gpurun_out/exp1_${wl}_shallow.err; echo "rc=$?" >> gpurun_out/exp1_${wl}_shallow.err
  FG_DEVICE_EPILOGUE=0 timeout 300 $B --workload $wl --no-e2e > gpurun_out/exp1_${wl}_hostepi.json 2> gpurun_out/exp1_${wl}_hostepi.err; echo "rc=$?" >> gpurun_out/exp1_${wl}_hostepi.err
done
FG_LANES=3 timeout 300 $B --workload clr --no-e2e > gpurun_out/exp1_clr_lanes3.json 2> gpurun_out/exp1_clr_lanes3.err
FG_SUBS_PER_LANE=2 timeout 300 $B --workload clr --no-e2e > gpurun_out/exp1_clr_subs2.json 2> gpurun_out/exp1_clr_subs2.err
tail -3 gpurun_out/exp1_pytest.log
for f in gpurun_out/exp1_*.json; do python - "$f" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ph=d["phases_ms_one_lane"]
    print(sys.argv[1], "ms", round(d["ms_per_step"],2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],2), "one-lane", round(d["roofline_pass"]["ms_per_step"],2), "sort", ph.get("hit_sort_radix"), "epi", ph.get("epilogue_dev"), "hostres", ph.get("host_results"), "hostdiv", ph.get("host_divergence"), "match", d["parity"]["match"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
done
