# end-of-round refresh: parity suite, ncu captures of the GEMM kernel (epilogue changed last), bench lines, smoke
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --timeout 400 ) > gpurun_out/final_pytest.log 2>&1; tail -4 gpurun_out/final_pytest.log
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1f_gemm_unet320 python tools/bench_gemm.py --shapes unet_c3_320_64 --iters 1 > gpurun_out/f_ncu1.log 2>&1
timeout 300 $NCU -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/ncu_r1f_gemm_lin320 python tools/bench_gemm.py --shapes lin_320_320_4096 --iters 1 > gpurun_out/f_ncu5.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1f_n1.json 2> gpurun_out/bench_r1f_n1.err; tail -1 gpurun_out/bench_r1f_n1.json | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 2 > gpurun_out/bench_r1f_reference.json 2> gpurun_out/bench_r1f_reference.err; tail -1 gpurun_out/bench_r1f_reference.json | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; tail -2 gpurun_out/f_smoke.log
