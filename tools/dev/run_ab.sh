summ() { python - "$1" <<PY
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1) if d["e2e"] else None, d["parity"]["match"], d["api_wall_ms"], d["step_wall_ms_rank0"])
except Exception as e: print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-400:])
PY
}
for w in clr hifi; do
for cfg in "2 1" "2 2" "2 3" "3 1" "3 2"; do set -- $cfg
FG_LANES=$1 FG_SUBS_PER_LANE=$2 python bench.py --workload $w --no-cpu-baseline --steps 4 > gpurun_out/r6_${w}_l$1s$2.json 2> gpurun_out/r6_${w}_l$1s$2.err; summ r6_${w}_l$1s$2
done; done
