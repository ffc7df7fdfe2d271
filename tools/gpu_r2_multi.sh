# 2-GPU call: on-GPU multi-rank parity test + 2-rank bench lines (c2 weak scaling, c3 'seg' slice)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_dist_gpu.py -m gpu -q --timeout 500 ) > gpurun_out/r2_multi_pytest.log 2>&1
tail -5 gpurun_out/r2_multi_pytest.log
grep "NCCL ranks" gpurun_out/parity.log | tail -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_c2_n2.json 2> gpurun_out/bench_r2_c2_n2.err; tail -1 gpurun_out/bench_r2_c2_n2.json | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c3 --steps 5 --warmup 3 > gpurun_out/bench_r2_c3_n2.json 2> gpurun_out/bench_r2_c3_n2.err; tail -1 gpurun_out/bench_r2_c3_n2.json | cut -c1-300
