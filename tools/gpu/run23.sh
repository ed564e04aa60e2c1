set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for ser in 0 1; do
RB2_EXCH_SERIAL=$ser RB2_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2 --warmup 3 --no-verify > gpurun_out/r2_n2_ser$ser.json 2> gpurun_out/r2_n2_ser$ser.err; echo "rc=$?"
grep "rb2 trace" gpurun_out/r2_n2_ser$ser.err | grep "rank 0" | tail -2
python - <<PY
import json
d=json.load(open("gpurun_out/r2_n2_ser$ser.json"))
print("SER=$ser", d["value"], d["ms_per_step"], d["e2e"]["value"], {k:round(v,1) for k,v in d["phases_ms_per_step"].items()})
PY
done
