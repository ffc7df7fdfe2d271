cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for pm in -1 1; do echo "== pair mode $pm"; timeout 200 python tools/bench_chain.py --ring 2 --pair $pm --cases geglu64,geglu32,geglu16,wide64,qkv64,qkv32,ffout64+res,lin64; done | tee gpurun_out/r2c45_chain_pair.txt
