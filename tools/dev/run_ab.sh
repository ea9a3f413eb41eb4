summ() { python - "$1" <<PY
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    a=d["phases_ms"]; b=d.get("phases_ms_e2e",{}); c=d["phases_ms_one_lane"]
    print(f, round(d["ms_per_step"],1), round(d["e2e"]["ms_per_step"],1) if d["e2e"] else None, d["parity"]["match"], "radix1lane", c.get("hit_sort_radix"), "lookup", c.get("lookup"))
    print("   prep:", {k:(a.get(k),b.get(k)) for k in a if k.startswith("prep") or k=="host_chunk_prep"})
except Exception as e: print(f, "ERR", e, open("gpurun_out/%s.err"%f).read()[-400:])
PY
}
FG_L2_PIN=7 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r5_clr_pin7.json 2> gpurun_out/r5_clr_pin7.err; summ r5_clr_pin7
FG_L2_PIN=3 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r5_clr_pin3.json 2> gpurun_out/r5_clr_pin3.err; summ r5_clr_pin3
FG_L2_PIN=3 FG_SEG_OCC=3 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r5_clr_occ3.json 2> gpurun_out/r5_clr_occ3.err; summ r5_clr_occ3
FG_SEG_OCC=4 python bench.py --workload hifi --no-cpu-baseline --steps 3 > gpurun_out/r5_hifi_occ4.json 2> gpurun_out/r5_hifi_occ4.err; summ r5_hifi_occ4
FG_SEG_OCC=3 python bench.py --workload hifi --no-cpu-baseline --steps 3 > gpurun_out/r5_hifi_occ3.json 2> gpurun_out/r5_hifi_occ3.err; summ r5_hifi_occ3
