cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_parity_rounded_gpu.py tests/test_blocks_gpu.py tests/test_e2e_gpu.py -q -x --timeout 120 ) > gpurun_out/r2c10_pytest.log 2>&1
tail -6 gpurun_out/r2c10_pytest.log
timeout 120 python tools/bench_geglu.py > gpurun_out/r2c10_bench_geglu.log 2>&1; cat gpurun_out/r2c10_bench_geglu.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c10_bench.json 2> gpurun_out/r2c10_bench.err; tail -1 gpurun_out/r2c10_bench.json | cut -c1-200
UR_GEMM_TMA_STORE=1 timeout 200 python tools/profile_graph.py --reps 2 > gpurun_out/r2c10_profile_store1.txt 2>/dev/null; sed -n 2,2p gpurun_out/r2c10_profile_store1.txt
timeout 200 python tools/profile_graph.py --reps 2 > gpurun_out/r2c10_profile_store2.txt 2>/dev/null; sed -n 2,2p gpurun_out/r2c10_profile_store2.txt
