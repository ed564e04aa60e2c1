set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -8
# sharded tests: virtual ranks on one GPU (LocalComm direct delivery), the RB2_GPUS cluster, two NCCL ranks (CUDA IPC)
timeout 900 python -m pytest tests/test_sharded_nccl.py tests/test_cluster_gpu.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q -k "test_uniform_one_batch or test_varlen or test_medium" 2>&1 | tail -5
# bench at N=2: direct delivery (default) and the send/recv exchange
for p2p in 1 0; do
RB2_P2P=$p2p timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_n2_p2p$p2p.json 2> gpurun_out/r2_n2_p2p$p2p.err; echo "rc=$?"
cut -c1-1800 gpurun_out/r2_n2_p2p$p2p.json; tail -3 gpurun_out/r2_n2_p2p$p2p.err
done
