cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -x --timeout 120 ) > gpurun_out/r2c17_gemm_test.log 2>&1
tail -4 gpurun_out/r2c17_gemm_test.log
timeout 200 python tools/bench_geglu.py > gpurun_out/r2c17_bench_geglu.txt 2>&1; cat gpurun_out/r2c17_bench_geglu.txt
UR_GEMM_TMA_STORE=2 timeout 200 python tools/bench_geglu.py > gpurun_out/r2c17_bench_geglu_mode2.txt 2>&1; cat gpurun_out/r2c17_bench_geglu_mode2.txt
timeout 200 python tools/trace_gemm.py lin_320_320_4096 2>&1 | head -8
timeout 200 python tools/bench_gemm.py > gpurun_out/r2c17_bench_gemm.txt 2>&1; cat gpurun_out/r2c17_bench_gemm.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c17_bench.json 2> gpurun_out/r2c17_bench.err; tail -1 gpurun_out/r2c17_bench.json | cut -c1-200
