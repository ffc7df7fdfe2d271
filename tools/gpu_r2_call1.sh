# round-2 call 1: parity suite (new rounded-oracle + benchmarked-path tests), bench line, sanitizer passes
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 ) > gpurun_out/r2c1_pytest.log 2>&1
tail -15 gpurun_out/r2c1_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c1_parity.log
timeout 600 python bench.py --torch-eager > gpurun_out/r2c1_bench.json 2> gpurun_out/r2c1_bench.err; tail -1 gpurun_out/r2c1_bench.json | cut -c1-1500
for tool in memcheck synccheck racecheck; do
  ( time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_forward.py --batch 1 --steps 2 ) > gpurun_out/r2c1_sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/r2c1_sanitizer_$tool.log
done
