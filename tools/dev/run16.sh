set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_host_mirror.py -m gpu -x -q -k second_container 2>&1 | tail -40
FG_WFA_BAND=0 timeout 300 python -m pytest tests/test_gpu_host_mirror.py -m gpu -x -q -k second_container 2>&1 | tail -12
