# round-2 A/B runs on the GPU box (gpurun -- 'bash tools/ab_r2.sh'): the tile version of the segmented sort against the
# register-scatter version on configs[0] and configs[1], parity tests of the variants first; one ncu capture of the new kernel
# after the plain run of the same command has exited 0.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_gpu_parity.py::test_kernel_variants_agree_with_oracle" tests/test_gpu_edge_cases.py -x -q -m gpu > gpurun_out/exp2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp2_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for wl in clr hifi; do
  for v in "0 3" "2 3" "2 4"; do
    set -- $v
    FG_SEG_SORT=$1 FG_SEG_OCC=$2 timeout 300 $B --workload $wl > gpurun_out/exp2_${wl}_sort$1_occ$2.json 2> gpurun_out/exp2_${wl}_sort$1_occ$2.err; echo "rc=$?" >> gpurun_out/exp2_${wl}_sort$1_occ$2.err
  done
done
export FG_LANES=1 FG_SEG_SORT=2
C="python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C > gpurun_out/exp2_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:segTileSortKernel -c 2 -f -o gpurun_out/exp2_tile $C > gpurun_out/exp2_ncu.log 2>&1
ncu -i gpurun_out/exp2_tile.ncu-rep --page raw --csv > gpurun_out/exp2_tile_raw.csv 2> gpurun_out/exp2_ncu_export.log
rm -f gpurun_out/exp2_tile.ncu-rep
tail -3 gpurun_out/exp2_pytest.log
