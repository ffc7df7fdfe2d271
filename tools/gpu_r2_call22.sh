cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/roofline_report.py > gpurun_out/roofline_r2_c22.md 2> gpurun_out/r2c22_roofline.err; tail -3 gpurun_out/r2c22_roofline.err; tail -50 gpurun_out/roofline_r2_c22.md
