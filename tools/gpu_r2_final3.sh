# validation of HEAD after the cross-attention kernel: sanitizer (memcheck, synccheck), full suite, smoke, bench, roofline
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  ( time timeout 400 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_forward.py --batch 1 --steps 2 ) > gpurun_out/r2h_sanitizer_${tool}_graph.log 2>&1
  grep -E "SUMMARY|sanitize_forward ok" gpurun_out/r2h_sanitizer_${tool}_graph.log | head -3
done
bash tools/gpu_r2_final2.sh
