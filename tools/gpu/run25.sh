set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/r2_final_cfg2.json 2> gpurun_out/r2_final_cfg2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final_cfg2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"], d["parity"]["result"][:60])
PY
timeout 600 python bench.py --config cfg4 --steps 1 --warmup 1 > gpurun_out/r2_final_cfg4.json 2> gpurun_out/r2_final_cfg4.err; echo "rc=$?"
cut -c1-300 gpurun_out/r2_final_cfg4.json; python -c "
import json; d=json.load(open('gpurun_out/r2_final_cfg4.json')); print(d['parity'])"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
