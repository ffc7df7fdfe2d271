cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_conv_gemm_gpu.py tests/test_blocks_gpu.py -m gpu -q -x --timeout 60 ) > gpurun_out/r2c40_test.log 2>&1
tail -2 gpurun_out/r2c40_test.log
timeout 300 python tools/bench_chain.py --ring 2 2>&1 | tee gpurun_out/r2c40_chain.txt
timeout 300 python tools/bench_gemm.py --shapes unet_c3_320_64,unet_c3_640_32,unet_c3_1280_16,vae_c3_256_256 2>&1 | tee gpurun_out/r2c40_bench_gemm.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c40_bench.json 2> gpurun_out/r2c40_bench.err; tail -1 gpurun_out/r2c40_bench.json | cut -c1-200
