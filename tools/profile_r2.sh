# Final artefacts of round 2 (committed under profiles/): gpurun -- 'bash tools/profile_r2.sh'.  Every ncu capture follows a plain
# run of the same command that exited 0.  The captures use one lane (FG_LANES=1): under ncu the kernels are serialised anyway,
# and the plain run then has the same launch sequence.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 400 python bench.py > gpurun_out/r2_bench_clr_1gpu.json 2> gpurun_out/r2_bench_clr_1gpu.err; echo "rc=$?" >> gpurun_out/r2_bench_clr_1gpu.err
timeout 400 python bench.py --workload hifi > gpurun_out/r2_bench_hifi_1gpu.json 2> gpurun_out/r2_bench_hifi_1gpu.err; echo "rc=$?" >> gpurun_out/r2_bench_hifi_1gpu.err
export FG_LANES=1
CLR="python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
HIFI="python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CLR > gpurun_out/plain_clr.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_clr_full.csv $CLR > gpurun_out/ncu_clr_l.log 2>&1
$HIFI > gpurun_out/plain_hifi.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_hifi_full.csv $HIFI > gpurun_out/ncu_hifi_l.log 2>&1
K='regex:segTileSortKernel|chainRunDp|chainWalk|chainFill|expandKernel|queryLookup|denseCount|selectKernel|pairPrep|divergenceKernel|gatherKept'
$CLR > gpurun_out/plain_clr2.log 2>&1 && timeout 420 ncu --set full --clock-control none --import-source on -k "$K" -c 12 -f -o gpurun_out/prof_clr $CLR > gpurun_out/ncu_clr.log 2>&1
ncu -i gpurun_out/prof_clr.ncu-rep --page raw --csv > gpurun_out/r2_full_clr_raw.csv 2> gpurun_out/ncu_export.log
$HIFI > gpurun_out/plain_hifi2.log 2>&1 && timeout 240 ncu --set full --clock-control none --import-source on -k 'regex:wfaKernel|segTileSortKernel' -c 2 -f -o gpurun_out/prof_hifi $HIFI > gpurun_out/ncu_hifi.log 2>&1
ncu -i gpurun_out/prof_hifi.ncu-rep --page raw --csv > gpurun_out/r2_full_hifi_raw.csv 2>> gpurun_out/ncu_export.log
rm -f gpurun_out/prof_clr.ncu-rep gpurun_out/prof_hifi.ncu-rep
tail -3 gpurun_out/r2_pytest_gpu.log
