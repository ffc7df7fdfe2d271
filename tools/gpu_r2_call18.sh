cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -m gpu -q --timeout 120 ) > gpurun_out/r2c18_gemm_test.log 2>&1
tail -4 gpurun_out/r2c18_gemm_test.log
timeout 200 python tools/trace_gemm.py lin_320_320_4096 lin_320_2560_4096 > gpurun_out/r2c18_trace.txt 2>&1; grep -v "      -       -       -       -   d=0" gpurun_out/r2c18_trace.txt
