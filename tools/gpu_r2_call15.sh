cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# quick: resident-weights test first (a hang here must not take the box: own timeout)
( timeout 300 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q -k "resident or store_modes" --timeout 120 ) > gpurun_out/r2c15_wres_test.log 2>&1
tail -4 gpurun_out/r2c15_wres_test.log
timeout 300 python tools/bench_wres.py > gpurun_out/r2c15_bench_wres.txt 2>&1; cat gpurun_out/r2c15_bench_wres.txt
timeout 200 python tools/bench_gemm.py --shapes vae_c3_128_512,vae_c3_256_256 > gpurun_out/r2c15_bench_gemm.txt 2>&1; cat gpurun_out/r2c15_bench_gemm.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c15_bench.json 2> gpurun_out/r2c15_bench.err; tail -1 gpurun_out/r2c15_bench.json | cut -c1-200
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c15_pytest.log 2>&1
tail -6 gpurun_out/r2c15_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c15_parity.log
