cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c5_pytest.log 2>&1
tail -8 gpurun_out/r2c5_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c5_parity.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:attention2_kernel -s 3 -c 1 -o gpurun_out/ncu_r2_attn2_self64 python tools/bench_attn.py --impl 2 --cases self64 --iters 1 > gpurun_out/r2c5_ncu1.log 2>&1
timeout 300 $NCU -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/ncu_r2_attn1_self64 python tools/bench_attn.py --impl 1 --cases self64 --iters 1 > gpurun_out/r2c5_ncu2.log 2>&1
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; tail -1 gpurun_out/r2c5_bench.json | cut -c1-300
ls -la gpurun_out/*.ncu-rep
