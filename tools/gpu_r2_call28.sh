cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for ab in 0 8 16 32 48 56; do
    echo "== ablate $ab (8 = no residual loads, 16 = no proxy fence, 32 = no TMA store issue)"
    UR_GEMM_ABLATE=$ab timeout 200 python tools/bench_chain.py --ring 2 --cases lin64,lin64+res,lin64+res+stats,qkv64,lin32+res
  done ) > gpurun_out/r2c28_ablate.txt 2>&1
cat gpurun_out/r2c28_ablate.txt
