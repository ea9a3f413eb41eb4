# final 1-GPU run of the round: whole GPU suite, bench lines (default workload with CPU baseline, reference arm, CLR),
# full-scale parity of both configs, launch lists of the final state
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py > gpurun_out/r11_bench_hifi.json 2> gpurun_out/r11_bench_hifi.err; tail -c 400 gpurun_out/r11_bench_hifi.err
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r11_bench_reference_hifi.json 2> gpurun_out/r11_bench_reference_hifi.err
python bench.py --workload clr > gpurun_out/r11_bench_clr.json 2> gpurun_out/r11_bench_clr.err
for w in hifi clr; do
  timeout 900 python tools/full_scale_parity.py $w > gpurun_out/r11_fullparity_${w}.json 2> gpurun_out/r11_fullparity_${w}.err; tail -c 200 gpurun_out/r11_fullparity_${w}.json
done
python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r11_plain_hifi.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r11_launches_hifi.csv python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r11_ncu_hifi.log 2>&1
python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r11_plain_clr.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r11_launches_clr.csv python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r11_ncu_clr.log 2>&1
python - <<PY
import json
for f in ["hifi","clr"]:
    d=json.loads(open("gpurun_out/r11_bench_%s.json"%f).read().strip().splitlines()[-1])
    print("RES", f, round(d["ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), d["phases_ms"], d.get("cpu_baseline"))
PY
