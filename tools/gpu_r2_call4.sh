cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_conv_gemm_gpu.py -q --timeout 600 ) > gpurun_out/r2c4_pytest.log 2>&1
tail -12 gpurun_out/r2c4_pytest.log
for impl in 1 2; do timeout 300 python tools/bench_attn.py --impl $impl --cases self64,self32,cross64,ctrl64,self16,ctrl128 >> gpurun_out/r2c4_bench_attn.log 2>&1; done
cat gpurun_out/r2c4_bench_attn.log
UR_PDL=0 UNIRESTORE_OVERLAP_CONTROLLER=0 UNIRESTORE_OVERLAP_SCTUNER=0 timeout 600 python tools/profile_graph.py > gpurun_out/r2c4_profile_serial.txt 2>/dev/null; head -24 gpurun_out/r2c4_profile_serial.txt
UR_PDL=0 UNIRESTORE_OVERLAP_CONTROLLER=0 UNIRESTORE_OVERLAP_SCTUNER=0 timeout 600 python tools/profile_graph.py --steps 1 > gpurun_out/r2c4_profile_serial_1step.txt 2>/dev/null
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2c4_bench.json 2> gpurun_out/r2c4_bench.err; tail -1 gpurun_out/r2c4_bench.json | cut -c1-300
