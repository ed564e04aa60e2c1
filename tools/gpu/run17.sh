set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?" >> gpurun_out/r2_bench_n8.err
cat gpurun_out/r2_bench_n8.json | cut -c1-1800; tail -3 gpurun_out/r2_bench_n8.err | cut -c1-300
