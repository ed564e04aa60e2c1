set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
RB2_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2_n8_direct.json 2> gpurun_out/r2_n8_direct.err; echo "rc=$?"
grep "rb2 trace" gpurun_out/r2_n8_direct.err | grep "rank 0" | awk '{print $6, $10, $13, $16}' | tail -9
python - <<PY
import json
d=json.load(open("gpurun_out/r2_n8_direct.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], {k:round(v,1) for k,v in d["phases_ms_per_step"].items()})
print(d["parity"]["result"][:90]); print(d["exchange"])
PY
tail -3 gpurun_out/r2_n8_direct.err | cut -c1-300
