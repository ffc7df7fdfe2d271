cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_parity_rounded_gpu.py tests/test_blocks_gpu.py -q -x --timeout 120 ) > gpurun_out/r2c9_pytest.log 2>&1
tail -6 gpurun_out/r2c9_pytest.log
timeout 120 python tools/bench_geglu.py > gpurun_out/r2c9_bench_geglu.log 2>&1; cat gpurun_out/r2c9_bench_geglu.log
timeout 100 python tools/trace_attn3.py --poly 3 > gpurun_out/r2c9_trace_attn3_poly3.log 2>&1; cat gpurun_out/r2c9_trace_attn3_poly3.log
timeout 100 python tools/trace_attn3.py --poly 0 > gpurun_out/r2c9_trace_attn3_poly0.log 2>&1; cat gpurun_out/r2c9_trace_attn3_poly0.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c9_bench.json 2> gpurun_out/r2c9_bench.err; tail -1 gpurun_out/r2c9_bench.json | cut -c1-200
