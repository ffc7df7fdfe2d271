cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_blocks_gpu.py tests/test_parity_rounded_gpu.py tests/test_e2e_gpu.py -m gpu -q --timeout 300 ) > gpurun_out/r2c48_test.log 2>&1
tail -3 gpurun_out/r2c48_test.log
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c48_bench.json 2> gpurun_out/r2c48_bench.err; tail -1 gpurun_out/r2c48_bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['launches_per_forward'], d['launches_by_entry_point'].get('ur_chan_stats'))"
