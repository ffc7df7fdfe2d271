# BASELINE.json configs[0], [2], [3], [4] as bench lines (single GPU: the per-GPU slice of the multi-GPU configurations)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/bench_r2_$tag.json 2> gpurun_out/bench_r2_$tag.err; tail -1 gpurun_out/bench_r2_$tag.json | cut -c1-260; }
run c1 --config c1
run c3_slice --config c3 --no-cpu-baseline
run c4_slice --config c4 --steps 3 --warmup 3 --no-cpu-baseline
for n in 1 4 10 20 50; do run c5_slice_n$n --config c5 --ddim-steps $n --no-cpu-baseline; done
run c2 --config c2 --torch-eager
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err; tail -1 gpurun_out/bench_r2_reference.json | cut -c1-260
