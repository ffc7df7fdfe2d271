cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_blocks_gpu.py tests/test_parity_rounded_gpu.py -m gpu -q --timeout 120 ) > gpurun_out/r2c44_test.log 2>&1
tail -2 gpurun_out/r2c44_test.log
timeout 200 python tools/bench_chain_norm.py 2>&1 | grep -E "^ln" | tee gpurun_out/r2c44_chain_ln.txt
