# final round-1 evidence: ncu full captures of the top kernels, launch lists, bench lines (ours + reference arm)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1f_gemm_unet320 python tools/bench_gemm.py --shapes unet_c3_320_64 --iters 1 > gpurun_out/f_ncu1.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1f_gemm_vae256 python tools/bench_gemm.py --shapes vae_c3_256_256 --iters 1 > gpurun_out/f_ncu2.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1f_gemm_lin320 python tools/bench_gemm.py --shapes lin_320_320_4096 --iters 1 > gpurun_out/f_ncu5.log 2>&1
timeout 300 $NCU -k regex:attention_kernel -s 3 -c 1 -o gpurun_out/ncu_r1f_attn_self64 python tools/bench_attn.py --cases self64 --iters 2 > gpurun_out/f_ncu3.log 2>&1
timeout 300 $NCU -k "regex:norm_apply|chan_stats|layernorm" -s 4 -c 7 -o gpurun_out/ncu_r1f_norm python tools/ncu_norm.py > gpurun_out/f_ncu4.log 2>&1
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
UNIRESTORE_OVERLAP_CONTROLLER=0 timeout 500 $L --log-file gpurun_out/launches_r1f_1step.csv python tools/ncu_run.py --steps 1 > gpurun_out/f_l1.log 2>&1
UNIRESTORE_OVERLAP_CONTROLLER=0 timeout 600 $L --log-file gpurun_out/launches_r1f_3step.csv python tools/ncu_run.py --steps 3 > gpurun_out/f_l3.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1f_n1.json 2> gpurun_out/bench_r1f_n1.err; tail -1 gpurun_out/bench_r1f_n1.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/bench_r1f_reference.json 2> gpurun_out/bench_r1f_reference.err; tail -1 gpurun_out/bench_r1f_reference.json | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
ls -la gpurun_out | tail -20
