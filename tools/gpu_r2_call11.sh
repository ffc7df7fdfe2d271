cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
for i in 1 2 3; do timeout 300 python -m pytest tests/test_parity_rounded_gpu.py -q -x --timeout 120 -k "controller_unet_encode_decode" 2>&1 | tail -2; done
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2c11_pytest.log 2>&1
tail -6 gpurun_out/r2c11_pytest.log
cp gpurun_out/parity.log gpurun_out/r2c11_parity.log
for tool in memcheck racecheck; do
  ( time timeout 600 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_forward.py --batch 1 --steps 2 --eager ) > gpurun_out/r2c11_sanitizer_${tool}_eager.log 2>&1
  grep -E "SUMMARY|sanitize_forward ok|Error:" gpurun_out/r2c11_sanitizer_${tool}_eager.log | sort | uniq -c | head -8
done
( time timeout 400 compute-sanitizer --tool synccheck --print-limit 30 python tools/sanitize_forward.py --batch 1 --steps 2 ) > gpurun_out/r2c11_sanitizer_synccheck_graph.log 2>&1
grep -E "SUMMARY|sanitize_forward ok|Error|Barrier error" gpurun_out/r2c11_sanitizer_synccheck_graph.log | sort | uniq -c | head -8
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c11_bench.json 2> gpurun_out/r2c11_bench.err; tail -1 gpurun_out/r2c11_bench.json | cut -c1-200
