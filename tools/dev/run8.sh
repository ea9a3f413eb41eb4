# profiles of the current state: launch lists (both workloads, full scale) and full counters of our kernels exported as CSV
set -x
mkdir -p gpurun_out
python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r8_plain_hifi.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r8_launches_hifi.csv python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r8_ncu_hifi.log 2>&1
python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r8_plain_clr.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r8_launches_clr.csv python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r8_ncu_clr.log 2>&1
# --steps 1 --warmup 0: the first pass and the e2e pass -> every kernel appears twice; capture the second half
ncu --set full --clock-control none -k regex:'segRadixSortKernel|expandKernel|chainRunDpKernel|chainFillKernel|chainRunsKernel|chainWalkKernel|wfaKernel|sortHugeKernel|groupFlagKernel|queryLookupKernel|minimizerKernel|pairFilterKernel|pairPrepKernel' -c 30 -o /tmp/r8_full_hifi python bench.py --workload hifi --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r8_ncufull_hifi.log 2>&1
ncu -i /tmp/r8_full_hifi.ncu-rep --page raw --csv > gpurun_out/r8_full_hifi_raw.csv 2> gpurun_out/r8_full_export.err
ls -la gpurun_out /tmp/r8_full_hifi.ncu-rep
du -sh gpurun_out
