cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q --timeout 60 ) > gpurun_out/r2c25_gemm_test.log 2>&1
tail -3 gpurun_out/r2c25_gemm_test.log
( echo "== lean epilogue kernel (default)"; timeout 300 python tools/bench_chain.py --ring 2
  echo "== general kernel"; UR_GEMM_LEAN=0 timeout 300 python tools/bench_chain.py --ring 2 ) > gpurun_out/r2c25_chain.txt 2>&1
cat gpurun_out/r2c25_chain.txt
timeout 200 python tools/trace_gemm.py lin_320_320_4096 2>&1 | grep -v "      -       -       -       -   d=0"
