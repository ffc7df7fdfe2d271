cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -k "resident" --timeout 60 ) > gpurun_out/r2c24_wres_test.log 2>&1
tail -4 gpurun_out/r2c24_wres_test.log
( echo "== resident weights (default)"; timeout 300 python tools/bench_chain.py --ring 2
  echo "== streamed weights"; UR_GEMM_WRES=0 timeout 300 python tools/bench_chain.py --ring 2 ) > gpurun_out/r2c24_chain.txt 2>&1
cat gpurun_out/r2c24_chain.txt
