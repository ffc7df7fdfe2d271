# compute-sanitizer pass over the final round-2 library (lean GEMM kernel with store warps, new norm_apply prologue)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  ( time timeout 500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_forward.py --batch 1 --steps 2 ) > gpurun_out/r2f_sanitizer_${tool}_graph.log 2>&1
  grep -E "SUMMARY|sanitize_forward ok" gpurun_out/r2f_sanitizer_${tool}_graph.log | head -3
done
( time timeout 600 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_forward.py --batch 1 --steps 1 --eager ) > gpurun_out/r2f_sanitizer_racecheck_eager.log 2>&1
grep -E "SUMMARY|sanitize_forward ok" gpurun_out/r2f_sanitizer_racecheck_eager.log | head -3
grep -E "Error:|hazard" gpurun_out/r2f_sanitizer_racecheck_eager.log | sed 's/+0x.*//' | sort | uniq -c | sort -rn | head -12
grep -A6 "hazard" gpurun_out/r2f_sanitizer_racecheck_eager.log | grep -E "ur_|\.cu" | sed 's/^ *//' | sort | uniq -c | sort -rn | head -12
