cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_chain_norm.py 2>&1 | grep "^gn" | tee gpurun_out/r2c36_chain_norm.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c36_bench.json 2> gpurun_out/r2c36_bench.err; tail -1 gpurun_out/r2c36_bench.json | cut -c1-200
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c36_pytest.log 2>&1
tail -6 gpurun_out/r2c36_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c36_parity.log
