# last validation of HEAD: full -m gpu suite, smoke, bench (with CPU baseline + same-GPU eager), roofline report
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
rm -f gpurun_out/parity.log
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r2g_pytest.log 2>&1
tail -4 gpurun_out/r2g_pytest.log
cp gpurun_out/parity.log gpurun_out/r2g_parity.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; tail -1 gpurun_out/r2g_smoke.log
timeout 900 python bench.py --torch-eager > gpurun_out/bench_r2g_c2.json 2> gpurun_out/bench_r2g_c2.err; tail -1 gpurun_out/bench_r2g_c2.json | cut -c1-240
timeout 600 python tools/roofline_report.py > gpurun_out/roofline_r2g.md 2> gpurun_out/r2g_roofline.err; sed -n 5p gpurun_out/roofline_r2g.md
