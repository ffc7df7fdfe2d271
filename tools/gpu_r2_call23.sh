cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( echo "== all cases, direct stores (default)"; timeout 300 python tools/bench_chain.py
  echo "== all cases, warp-TMA stores"; UR_GEMM_TMA_STORE=2 timeout 300 python tools/bench_chain.py --ring 2
  for ab in 1 2 4 7; do
    echo "== ablate $ab (1 = no stores, 2 = no TMEM loads, 4 = no arithmetic)"
    UR_GEMM_ABLATE=$ab timeout 200 python tools/bench_chain.py --ring 2 --cases lin64,lin64+res,qkv64,gelu64,geglu64
  done ) > gpurun_out/r2c23_chain.txt 2>&1
cat gpurun_out/r2c23_chain.txt
