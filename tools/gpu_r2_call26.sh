cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q --timeout 60 ) > gpurun_out/r2c27_gemm_test.log 2>&1
tail -3 gpurun_out/r2c27_gemm_test.log
( echo "== lean kernel, trimmed per-tile prologue"; timeout 300 python tools/bench_chain.py --ring 2 ) > gpurun_out/r2c27_chain.txt 2>&1
cat gpurun_out/r2c27_chain.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c27_bench.json 2> gpurun_out/r2c27_bench.err; tail -1 gpurun_out/r2c27_bench.json | cut -c1-200
