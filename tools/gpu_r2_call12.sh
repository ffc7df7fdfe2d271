cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c12_pytest.log 2>&1
tail -6 gpurun_out/r2c12_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c12_parity.log
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python tools/sanitize_forward.py --batch 1 --steps 2 --eager ) > gpurun_out/r2c12_sanitizer_racecheck_eager.log 2>&1
grep -E "SUMMARY|sanitize_forward ok|Error:" gpurun_out/r2c12_sanitizer_racecheck_eager.log | sed 's/+0x.*//' | sort | uniq -c | head -8
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c12_bench.json 2> gpurun_out/r2c12_bench.err; tail -1 gpurun_out/r2c12_bench.json | cut -c1-200
