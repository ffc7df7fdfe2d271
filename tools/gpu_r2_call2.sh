# round-2 call 2: attention2 kernel parity + timing, full suite (no -x), bench
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 600 python -m pytest tests/test_attention_gpu.py -q --timeout 120 ) > gpurun_out/r2c2_pytest_attn.log 2>&1
tail -12 gpurun_out/r2c2_pytest_attn.log
for impl in 1 2; do timeout 300 python tools/bench_attn.py --impl $impl --cases self64,self32,cross64,ctrl64,self16,ctrl128 >> gpurun_out/r2c2_bench_attn.log 2>&1; done
cat gpurun_out/r2c2_bench_attn.log
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_attention_gpu.py ) > gpurun_out/r2c2_pytest.log 2>&1
tail -25 gpurun_out/r2c2_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c2_parity.log
timeout 600 python bench.py --torch-eager > gpurun_out/r2c2_bench.json 2> gpurun_out/r2c2_bench.err; tail -1 gpurun_out/r2c2_bench.json | cut -c1-400
