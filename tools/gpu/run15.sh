set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_nccl.py tests/test_sharded_gpu.py tests/test_cluster_gpu.py -m gpu -x -q > gpurun_out/r2_t15.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t15.log; tail -3 gpurun_out/r2_t15.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err; echo "bench n2 rc=$?" >> gpurun_out/r2_bench_n2b.err
cat gpurun_out/r2_bench_n2b.json | cut -c1-1800; tail -4 gpurun_out/r2_bench_n2b.err | cut -c1-300
