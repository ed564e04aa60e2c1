set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 280 python bench.py > gpurun_out/r2_final2_cfg2.json 2> gpurun_out/r2_final2_cfg2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_final2_cfg2.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], {k:round(v,1) for k,v in d["phases_ms_per_step"].items()}, d["parity"]["result"][:40])
PY
timeout 330 python -m pytest tests -m gpu -x -q --deselect tests/test_parity_gpu.py::test_beyond_2_32_symbols_md5_vs_reference 2>&1 | tail -4
