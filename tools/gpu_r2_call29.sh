cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( echo "== raw rcp in SiLU / GELU"; timeout 300 python tools/bench_chain.py --ring 2 --cases gelu64,geglu64,geglu32,lin64 ) > gpurun_out/r2c29_chain.txt 2>&1
cat gpurun_out/r2c29_chain.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c29_bench.json 2> gpurun_out/r2c29_bench.err; tail -1 gpurun_out/r2c29_bench.json | cut -c1-200
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c29_pytest.log 2>&1
tail -6 gpurun_out/r2c29_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c29_parity.log
