cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_parity_rounded_gpu.py tests/test_path_gpu.py tests/test_e2e_gpu.py -q --timeout 600 ) > gpurun_out/r2c3_pytest.log 2>&1
tail -12 gpurun_out/r2c3_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c3_parity.log
timeout 600 python tools/profile_graph.py > gpurun_out/r2c3_profile_graph.txt 2> gpurun_out/r2c3_profile_graph.err; head -40 gpurun_out/r2c3_profile_graph.txt; tail -3 gpurun_out/r2c3_profile_graph.err
timeout 300 python tools/profile_graph.py --steps 1 > gpurun_out/r2c3_profile_graph_1step.txt 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c3_bench.json 2> gpurun_out/r2c3_bench.err; tail -1 gpurun_out/r2c3_bench.json | cut -c1-300
