set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
RB2_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-verify > gpurun_out/r2_n2_trace.json 2> gpurun_out/r2_n2_trace.err; echo "rc=$?"
grep "rb2 trace" gpurun_out/r2_n2_trace.err | grep "rank 0"
cut -c1-300 gpurun_out/r2_n2_trace.json
