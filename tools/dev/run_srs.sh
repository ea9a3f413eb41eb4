timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "variants or clr_small or hifi" 2>&1 | tail -5
summ() { python - "$1" <<PY
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1) if d["e2e"] else None, d["parity"]["match"], {k:round(d["phases_ms"][k],2) for k in ("expand","hit_sort_radix","group","chain_prep")})
except Exception as e: print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-400:])
PY
}
for c in 0 1 2 4; do FG_LANES=1 FG_SRS_CLUSTER=$c timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/r3_srs_clr_c$c.json 2> gpurun_out/r3_srs_clr_c$c.err; summ r3_srs_clr_c$c; done
FG_LANES=1 FG_SEG_SORT=0 timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/r3_srs_clr_old.json 2> gpurun_out/r3_srs_clr_old.err; summ r3_srs_clr_old
for c in 0 2 4 8; do FG_LANES=1 FG_SRS_CLUSTER=$c timeout 200 python bench.py --workload hifi --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/r3_srs_hifi_c$c.json 2> gpurun_out/r3_srs_hifi_c$c.err; summ r3_srs_hifi_c$c; done
FG_LANES=1 FG_SEG_SORT=0 timeout 200 python bench.py --workload hifi --no-cpu-baseline --no-e2e --steps 3 > gpurun_out/r3_srs_hifi_old.json 2> gpurun_out/r3_srs_hifi_old.err; summ r3_srs_hifi_old
