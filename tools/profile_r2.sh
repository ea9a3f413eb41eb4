set -x
export FG_LANES=1
K='regex:denseCount|selectKernel|queryLookup|segRadixSort|chainRunDp|expandKernel|chainWalk|insertIndex|chainFill|pairPrep|emitWrite|denseHist|pairFilter|sortTail|sortLevel|sortSmall|classify|rleWrite|tandemCount|finalizeSelection'
CLR="python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
HIFI="python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CLR > gpurun_out/plain_clr.log 2>&1 && ncu --set full --clock-control none --import-source on -k "$K" -c 60 -f -o gpurun_out/prof_clr $CLR > gpurun_out/ncu_clr.log 2>&1
ncu -i gpurun_out/prof_clr.ncu-rep --page raw --csv > gpurun_out/r2_full_clr_raw.csv 2> gpurun_out/ncu_export.log
$CLR > gpurun_out/plain_clr2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_clr_full.csv $CLR > gpurun_out/ncu_clr_l.log 2>&1
$HIFI > gpurun_out/plain_hifi.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:wfaKernel|segRadixSort|minimizerReg|hpcReads|chainRunDp|expandKernel' -c 16 -f -o gpurun_out/prof_hifi $HIFI > gpurun_out/ncu_hifi.log 2>&1
ncu -i gpurun_out/prof_hifi.ncu-rep --page raw --csv > gpurun_out/r2_full_hifi_raw.csv 2>> gpurun_out/ncu_export.log
$HIFI > gpurun_out/plain_hifi2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_hifi_full.csv $HIFI > gpurun_out/ncu_hifi_l.log 2>&1
rm -f gpurun_out/prof_clr.ncu-rep gpurun_out/prof_hifi.ncu-rep
ls -la gpurun_out | tail -20
