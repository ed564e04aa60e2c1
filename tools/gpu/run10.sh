set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t10.log
tail -5 gpurun_out/r2_t10.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?" >> gpurun_out/r2_bench_n2.err
cat gpurun_out/r2_bench_n2.json | cut -c1-1500; tail -5 gpurun_out/r2_bench_n2.err
