cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( for g in 256 128; do echo "== gated N tile $g"; timeout 200 python tools/bench_chain.py --ring 2 --gbn $g --cases geglu64,geglu32,geglu16; done
  for bn in 256 160 128; do echo "== N tile $bn"; timeout 200 python tools/bench_chain.py --ring 2 --bn $bn --cases wide64,qkv16,naf768; done ) > gpurun_out/r2c31_chain_bn.txt 2>&1
cat gpurun_out/r2c31_chain_bn.txt
