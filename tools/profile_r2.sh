# ncu captures committed under profiles/ (round 2).  Every capture follows a plain run of the same command that exited 0.
# One lane (FG_LANES=1): under ncu the kernels are serialised anyway, and the plain run then has the same launch sequence.
set -x
export FG_LANES=1
CLR="python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
HIFI="python bench.py --workload hifi --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
K='regex:denseCount|selectKernel|queryLookup|segRadixSortKernel|chainRunDp|expandKernel|chainWalk|insertIndex|chainFill|pairPrep|emitWrite|pairFilter|sortTail|sortSmall|classify|indexBits|tandemCount|finalizeSelection|denseHist'
$CLR > gpurun_out/plain_clr.log 2>&1 && ncu --set full --clock-control none --import-source on -k "$K" -c 40 -f -o gpurun_out/prof_clr $CLR > gpurun_out/ncu_clr.log 2>&1
ncu -i gpurun_out/prof_clr.ncu-rep --page raw --csv > gpurun_out/r2_full_clr_raw.csv 2> gpurun_out/ncu_export.log
$CLR > gpurun_out/plain_clr2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_clr_full.csv $CLR > gpurun_out/ncu_clr_l.log 2>&1
$HIFI > gpurun_out/plain_hifi.log 2>&1 && ncu --set full --clock-control none --import-source on -k 'regex:wfaKernel|segRadixSortKernel|minimizerReg|hpcReads|chainRunDp|expandKernel' -c 12 -f -o gpurun_out/prof_hifi $HIFI > gpurun_out/ncu_hifi.log 2>&1
ncu -i gpurun_out/prof_hifi.ncu-rep --page raw --csv > gpurun_out/r2_full_hifi_raw.csv 2>> gpurun_out/ncu_export.log
$HIFI > gpurun_out/plain_hifi2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_hifi_full.csv $HIFI > gpurun_out/ncu_hifi_l.log 2>&1
# the thread-block-cluster variant of the segmented sort (not the default): its DRAM traffic next to the default kernel's
export FG_SEG_SORT=1 FG_SRS_CLUSTER=1
$CLR > gpurun_out/plain_clr_cluster.log 2>&1 && ncu --set full --clock-control none -k 'regex:segRadixSortCluster' -c 2 -f -o gpurun_out/prof_clr_cluster $CLR > gpurun_out/ncu_clr_cluster.log 2>&1
ncu -i gpurun_out/prof_clr_cluster.ncu-rep --page raw --csv > gpurun_out/r2_full_clr_cluster_raw.csv 2>> gpurun_out/ncu_export.log
rm -f gpurun_out/prof_clr.ncu-rep gpurun_out/prof_hifi.ncu-rep gpurun_out/prof_clr_cluster.ncu-rep
ls -la gpurun_out | tail -20
