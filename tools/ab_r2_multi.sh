set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/exp4_pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/exp4_pytest_multi.log
for wl in clr hifi; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl > gpurun_out/exp4_${wl}_2gpu.json 2> gpurun_out/exp4_${wl}_2gpu.err; echo "rc=$?" >> gpurun_out/exp4_${wl}_2gpu.err
done
tail -3 gpurun_out/exp4_pytest_multi.log
