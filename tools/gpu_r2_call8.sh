cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests/test_attention_gpu.py -q -x --timeout 60 ) > gpurun_out/r2c8_pytest.log 2>&1
tail -4 gpurun_out/r2c8_pytest.log
for impl in 2 3; do for poly in 0 1 2 3; do timeout 120 python tools/bench_attn.py --impl $impl --poly $poly --cases self64,self32,ctrl64 >> gpurun_out/r2c8_bench_attn.log 2>&1; done; done
timeout 120 python tools/bench_attn.py --impl 1 --cases self64,self32,ctrl64 >> gpurun_out/r2c8_bench_attn.log 2>&1
cat gpurun_out/r2c8_bench_attn.log
