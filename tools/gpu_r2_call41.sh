cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for w in 8 4 3 6; do echo "== 64-register norm_apply, UR_NORM_WAVES=$w"; UR_NORM_WAVES=$w timeout 200 python tools/bench_chain_norm.py 2>&1 | grep -E "^gn"; done | tee gpurun_out/r2c41_norm_waves.txt
