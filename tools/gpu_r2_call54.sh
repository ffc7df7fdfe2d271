cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_attention_gpu.py -m gpu -q -x --timeout 120 ) > gpurun_out/r2c54_test.log 2>&1
tail -4 gpurun_out/r2c54_test.log
for kv1 in 1 0; do timeout 120 python tools/bench_attn.py --graph --impl 1 --kv1 $kv1 --cases cross64,cross32,cross16 2>&1 | grep impl | sed "s/^/kv1=$kv1 /"; done | tee gpurun_out/r2c54_attn_kv1.txt
