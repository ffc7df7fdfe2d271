cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_persistent -s 3 -c 1 -o gpurun_out/r2c19_lin64 -f python tools/ncu_gemm_one.py 32768 320 320 > gpurun_out/r2c19_ncu.log 2>&1
tail -3 gpurun_out/r2c19_ncu.log
ls -la gpurun_out/*.ncu-rep
