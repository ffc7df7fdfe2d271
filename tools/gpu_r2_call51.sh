cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_blocks_gpu.py tests/test_parity_rounded_gpu.py -m gpu -q --timeout 300 ) > gpurun_out/r2c51_test.log 2>&1
tail -3 gpurun_out/r2c51_test.log
timeout 200 python tools/bench_chain.py --ring 2 --cases gelu64,geglu64,geglu32,geglu16 2>&1 | tee gpurun_out/r2c51_chain.txt
