cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 200 python tools/trace_gemm.py lin_320_320_4096 lin_320_2560_4096 > gpurun_out/r2c16_trace_lin.txt 2>&1; head -30 gpurun_out/r2c16_trace_lin.txt
timeout 200 python tools/bench_gemm.py --shapes vae_c3_128_512,vae_c3_256_256,lin_320_320_4096 > gpurun_out/r2c16_bench_gemm.txt 2>&1; cat gpurun_out/r2c16_bench_gemm.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c16_bench.json 2> gpurun_out/r2c16_bench.err; tail -1 gpurun_out/r2c16_bench.json | cut -c1-200
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c16_pytest.log 2>&1
tail -6 gpurun_out/r2c16_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c16_parity.log
