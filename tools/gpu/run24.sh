set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sharded_nccl.py tests/test_cluster_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -5
RB2_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_n2_ids.json 2> gpurun_out/r2_n2_ids.err; echo "rc=$?"
grep "rb2 trace" gpurun_out/r2_n2_ids.err | grep "rank 0" | tail -3
python - <<PY
import json
d=json.load(open("gpurun_out/r2_n2_ids.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k:round(v,1) for k,v in d["phases_ms_per_step"].items()})
print(d["parity"]["result"][:80])
PY
