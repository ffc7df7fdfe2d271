cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ab in 0 1 2 4 3 7; do
  echo "== ablate $ab (1 = no stores, 2 = no TMEM loads, 4 = no arithmetic), L2 flushed"
  UR_GEMM_ABLATE=$ab timeout 120 python tools/bench_geglu.py 2>&1 | grep -E "lin64|qkv64|geglu64|gelu64"
  echo "== ablate $ab, warm L2"
  UR_GEMM_ABLATE=$ab timeout 120 python tools/bench_geglu.py --noflush 2>&1 | grep -E "lin64|qkv64|geglu64|gelu64"
done > gpurun_out/r2c21_ablate.txt 2>&1
cat gpurun_out/r2c21_ablate.txt
timeout 200 python tools/trace_gemm.py lin_320_320_4096 > gpurun_out/r2c21_trace.txt 2>&1; grep -v "      -       -       -       -   d=0" gpurun_out/r2c21_trace.txt
