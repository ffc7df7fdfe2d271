cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q --timeout 60 ) > gpurun_out/r2c30_gemm_test.log 2>&1
tail -2 gpurun_out/r2c30_gemm_test.log
( for bn in 0 64 128 160 256; do echo "== N tile $bn"; timeout 200 python tools/bench_chain.py --ring 2 --bn $bn --cases lin64,lin64+res,qkv64,ffout64+res,lin32+res,qkv32,lin16+res; done ) > gpurun_out/r2c30_chain_bn.txt 2>&1
cat gpurun_out/r2c30_chain_bn.txt
