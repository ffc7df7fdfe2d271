cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_blocks_gpu.py -m gpu -q -x --timeout 120 ) > gpurun_out/r2c33_test.log 2>&1
tail -3 gpurun_out/r2c33_test.log
timeout 300 python tools/bench_chain.py --ring 2 2>&1 | tee gpurun_out/r2c33_chain.txt
timeout 300 python tools/bench_gemm.py 2>&1 | tee gpurun_out/r2c33_bench_gemm.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c33_bench.json 2> gpurun_out/r2c33_bench.err; tail -1 gpurun_out/r2c33_bench.json | cut -c1-200
