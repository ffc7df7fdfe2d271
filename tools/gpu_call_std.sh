# standard GPU round: parity tests, micro-benchmarks, traces, full bench.  usage: bash tools/gpu_call_std.sh TAG
TAG=${1:-cX}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 ) > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/trace_gemm.py unet_c3_320_64:160:1 lin_320_320_4096:160:0 > gpurun_out/${TAG}_trace.log 2>&1
timeout 300 python tools/bench_gemm.py > gpurun_out/${TAG}_bench_gemm.log 2>&1
timeout 300 python tools/bench_norm.py > gpurun_out/${TAG}_bench_norm.log 2>&1; cat gpurun_out/${TAG}_bench_norm.log
timeout 100 python tools/trace_attn.py > gpurun_out/${TAG}_trace_attn.log 2>&1
timeout 300 python tools/bench_attn.py --cases self64,self32,cross64,ctrl64,self16 > gpurun_out/${TAG}_bench_attn.log 2>&1
cat gpurun_out/${TAG}_bench_gemm.log gpurun_out/${TAG}_bench_attn.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.json
