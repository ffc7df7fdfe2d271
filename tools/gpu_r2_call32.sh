cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_blocks_gpu.py -m gpu -q --timeout 120 ) > gpurun_out/r2c32_test.log 2>&1
tail -2 gpurun_out/r2c32_test.log
timeout 200 python tools/bench_chain.py --ring 2 --cases qkv64,qkv32,qkv16,lin64 2>&1 | tee gpurun_out/r2c32_chain.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c32_bench.json 2> gpurun_out/r2c32_bench.err; tail -1 gpurun_out/r2c32_bench.json | cut -c1-200
