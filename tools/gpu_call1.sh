set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
tail -3 gpurun_out/c1_pytest.log
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -2 gpurun_out/c1_bench.json
timeout 300 python tools/bench_gemm.py > gpurun_out/c1_bench_gemm.log 2>&1
timeout 300 python tools/bench_attn.py --cases self64,self32,cross64,ctrl64,self16 > gpurun_out/c1_bench_attn.log 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1_gemm_unet320 python tools/bench_gemm.py --shapes unet_c3_320_64 --iters 1 > gpurun_out/c1_ncu1.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1_gemm_vae256 python tools/bench_gemm.py --shapes vae_c3_256_256 --iters 1 > gpurun_out/c1_ncu2.log 2>&1
timeout 300 $NCU -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/ncu_r1_attn_self64 python tools/bench_attn.py --cases self64 --iters 2 > gpurun_out/c1_ncu3.log 2>&1
timeout 300 $NCU -k "regex:norm_apply|chan_stats|layernorm" -s 4 -c 7 -o gpurun_out/ncu_r1_norm python tools/ncu_norm.py > gpurun_out/c1_ncu4.log 2>&1
ls -la gpurun_out
