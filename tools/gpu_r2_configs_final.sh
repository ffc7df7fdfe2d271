# refresh of the BASELINE configs[0], [2], [3], [4] lines at the final round-2 commit (single GPU = per-GPU slice)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/bench_r2f_$tag.json 2> gpurun_out/bench_r2f_$tag.err; tail -1 gpurun_out/bench_r2f_$tag.json | cut -c1-200; }
run c1 --config c1 --no-cpu-baseline
run c3_slice --config c3 --no-cpu-baseline
run c4_slice --config c4 --steps 3 --warmup 3 --no-cpu-baseline
for n in 1 4 10 20 50; do run c5_slice_n$n --config c5 --ddim-steps $n --no-cpu-baseline; done
# ncu summaries of the normalisation kernels at the two shapes VERDICT r1 asked for (320 @ 64x64, 1280 @ 8x8)
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:norm_apply -s 2 -c 1 -o gpurun_out/ncu_r2_norm_320_64 python tools/ncu_norm.py 8 64 64 320 > gpurun_out/r2f_ncu5.log 2>&1
timeout 300 $NCU -k regex:norm_apply -s 2 -c 1 -o gpurun_out/ncu_r2_norm_1280_8 python tools/ncu_norm.py 8 8 8 1280 > gpurun_out/r2f_ncu6.log 2>&1
for r in norm_320_64 norm_1280_8; do python tools/ncu_summary.py gpurun_out/ncu_r2_$r.ncu-rep > gpurun_out/ncu_r2_$r.txt 2>&1; done
grep -E "gpu__time_duration|dram__bytes_read.sum |dram__bytes_write.sum |lts__t_sector_hit" gpurun_out/ncu_r2_norm_320_64.txt gpurun_out/ncu_r2_norm_1280_8.txt
