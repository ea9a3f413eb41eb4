# development helper: launch list + full counters of the new kernels (HiFi full scale; CLR launch list)
set -x
mkdir -p gpurun_out
python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_plain_hifi.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches_hifi.csv python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_ncu_hifi.log 2>&1
python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_plain_clr.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches_clr.csv python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r3_ncu_clr.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'chainRunDpKernel|chainFillKernel|chainRunsKernel|rebuildHitsKernel|Onesweep|expandKernel|chainWalkKernel|wfaKernel|sortLevelKernel' -c 40 -o gpurun_out/r3_full_hifi python bench.py --workload hifi --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r3_ncufull_hifi.log 2>&1
ls -la gpurun_out
