# final round-2 evidence: full -m gpu suite, bench lines (ours incl. CPU baseline + same-GPU torch eager, reference arm),
# roofline report, ncu full captures of the top kernels, launch lists of the bench command, smoke
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2f_pytest.log 2>&1
tail -4 gpurun_out/r2f_pytest.log
cp gpurun_out/parity.log gpurun_out/r2f_parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 900 python bench.py --torch-eager > gpurun_out/bench_r2f_c2.json 2> gpurun_out/bench_r2f_c2.err; tail -1 gpurun_out/bench_r2f_c2.json | cut -c1-240
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2f_reference.json 2> gpurun_out/bench_r2f_reference.err; tail -1 gpurun_out/bench_r2f_reference.json | cut -c1-240
timeout 600 python tools/roofline_report.py > gpurun_out/roofline_r2f.md 2> gpurun_out/r2f_roofline.err; head -12 gpurun_out/roofline_r2f.md | cut -c1-200
timeout 300 python tools/bench_chain.py --ring 2 > gpurun_out/r2f_chain.txt 2>&1
timeout 300 python tools/bench_chain_norm.py > gpurun_out/r2f_chain_norm.txt 2>&1
timeout 300 python tools/bench_gemm.py > gpurun_out/r2f_bench_gemm.txt 2>&1
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r2_gemm_unet320 python tools/bench_gemm.py --shapes unet_c3_320_64 --iters 1 > gpurun_out/r2f_ncu1.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r2_gemm_lin320 python tools/ncu_gemm_one.py 32768 320 320 > gpurun_out/r2f_ncu2.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r2_gemm_geglu320 python tools/ncu_gemm_one.py 32768 320 2560 geglu > gpurun_out/r2f_ncu3.log 2>&1
timeout 300 $NCU -k "regex:norm_apply|layernorm" -s 4 -c 6 -o gpurun_out/ncu_r2_norm python tools/ncu_norm.py > gpurun_out/r2f_ncu4.log 2>&1
for r in gemm_unet320 gemm_lin320 gemm_geglu320 norm; do python tools/ncu_summary.py gpurun_out/ncu_r2_$r.ncu-rep > gpurun_out/ncu_r2_$r.txt 2>&1; done
L="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
timeout 900 $L -c 12000 --log-file gpurun_out/launches_r2_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_l1.log 2>&1
ls -la gpurun_out | grep -E "r2f|ncu_r2|launches_r2|bench_r2f|roofline_r2f" | awk '{print $5, $9}'
