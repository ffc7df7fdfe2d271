cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_chain_norm.py 2>&1 | tee gpurun_out/r2c35_chain_norm.txt
for w in 4 16; do echo "== UR_NORM_WAVES=$w"; UR_NORM_WAVES=$w timeout 300 python tools/bench_chain_norm.py 2>&1 | grep -E "^gn|^stats" ; done | tee gpurun_out/r2c35_chain_norm_waves.txt
