# round-2 A/B runs on the GPU box (gpurun -- 'bash tools/ab_r2.sh'): parity tests of the kernel variants first, then configs[0] and
# configs[1] with one switch flipped per run (every bench run checks the digest of its overlaps against the reference's), then
# one ncu capture of the segmented sort after the plain run of the same command has exited 0.
set -x
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest "tests/test_gpu_parity.py::test_kernel_variants_agree_with_oracle" tests/test_gpu_edge_cases.py tests/test_gpu_parity.py::test_clr_small_full_parity tests/test_gpu_parity.py::test_hifi_small_full_parity tests/test_gpu_real_sequence.py -x -q -m gpu > gpurun_out/exp3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/exp3_pytest.log
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for wl in clr hifi; do
  timeout 300 $B --workload $wl > gpurun_out/exp3_${wl}_default.json 2> gpurun_out/exp3_${wl}_default.err; echo "rc=$?" >> gpurun_out/exp3_${wl}_default.err
  FG_FILL_PAIRED=0 timeout 300 $B --workload $wl > gpurun_out/exp3_${wl}_fill0.json 2> gpurun_out/exp3_${wl}_fill0.err; echo "rc=$?" >> gpurun_out/exp3_${wl}_fill0.err
done
export FG_LANES=1
C="python bench.py --workload clr --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$C > gpurun_out/exp3_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:'segTileSortKernel|chainFillKernel' -c 4 -f -o gpurun_out/exp3_tile $C > gpurun_out/exp3_ncu.log 2>&1
ncu -i gpurun_out/exp3_tile.ncu-rep --page raw --csv > gpurun_out/exp3_tile_raw.csv 2> gpurun_out/exp3_ncu_export.log
rm -f gpurun_out/exp3_tile.ncu-rep
tail -3 gpurun_out/exp3_pytest.log
