cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_attention_gpu.py -q -x --timeout 60 ) > gpurun_out/r2c14_pytest.log 2>&1
tail -5 gpurun_out/r2c14_pytest.log
if grep -q "failed\|error" gpurun_out/r2c14_pytest.log; then echo "ATTENTION TESTS FAILED"; else
for poly in 0 1 2 3; do timeout 100 python tools/bench_attn.py --impl 4 --poly $poly --cases self64,self32,cross64,ctrl64,self16 >> gpurun_out/r2c14_bench_attn.log 2>&1; done
timeout 100 python tools/bench_attn.py --impl 1 --cases self64,self32,cross64,ctrl64,self16 >> gpurun_out/r2c14_bench_attn.log 2>&1
cat gpurun_out/r2c14_bench_attn.log
fi
