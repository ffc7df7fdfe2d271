cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/bench_gemm.py --shapes vae_c3_128_512,vae_c3_256_256,unet_c3_320_64,lin_320_320_4096,lin_320_2560_4096,lin_1280_1280_256 > gpurun_out/r2c13_bench_gemm.log 2>&1; cat gpurun_out/r2c13_bench_gemm.log
bash tools/gpu_r2_configs.sh
timeout 400 python tools/roofline_report.py > gpurun_out/roofline_r2_c13.md 2> gpurun_out/r2c13_roofline.err; sed -n 5,6p gpurun_out/roofline_r2_c13.md
