# 2-GPU run: NCCL-merged counts / index, sharded queries (parity test + bench lines)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
for w in hifi clr; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --steps 3 --warmup 3 > gpurun_out/r12_bench_${w}_2gpu.json 2> gpurun_out/r12_bench_${w}_2gpu.err
  tail -c 300 gpurun_out/r12_bench_${w}_2gpu.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r12_bench_${w}_2gpu.json").read().strip().splitlines()[-1])
print("RES2 $w", round(d["ms_per_step"],1), round(d["value"]), round(d["e2e"]["value"]), d["phases_ms"])
PY
done
