cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 300 python -m pytest tests/test_attention_gpu.py -q -x --timeout 60 ) > gpurun_out/r2c7_pytest.log 2>&1
tail -8 gpurun_out/r2c7_pytest.log
if grep -q "failed" gpurun_out/r2c7_pytest.log; then echo "ATTENTION TESTS FAILED: skipping attention3 timing"; else
for impl in 1 3; do timeout 120 python tools/bench_attn.py --impl $impl --cases self64,self32,cross64,ctrl64,self16,ctrl128 >> gpurun_out/r2c7_bench_attn.log 2>&1; done
cat gpurun_out/r2c7_bench_attn.log
timeout 200 python -c "
from unirestore_b200 import _cabi; _cabi.lib().ur_debug_set_attention_impl(3)
import runpy, sys; sys.argv=['bench.py','--no-cpu-baseline']; runpy.run_path('bench.py', run_name='__main__')" > gpurun_out/r2c7_bench_impl3.json 2> gpurun_out/r2c7_bench_impl3.err; tail -1 gpurun_out/r2c7_bench_impl3.json | cut -c1-200
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 200 $NCU -k regex:attention3_kernel -s 3 -c 1 -o gpurun_out/ncu_r2_attn3_self64 python tools/bench_attn.py --impl 3 --cases self64 --iters 1 > gpurun_out/r2c7_ncu1.log 2>&1
fi
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/r2c7_bench_impl1.json 2> gpurun_out/r2c7_bench_impl1.err; tail -1 gpurun_out/r2c7_bench_impl1.json | cut -c1-200
timeout 400 python tools/roofline_report.py > gpurun_out/roofline_r2_c7.md 2> gpurun_out/r2c7_roofline.err; head -40 gpurun_out/roofline_r2_c7.md; tail -3 gpurun_out/r2c7_roofline.err
