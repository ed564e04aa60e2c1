set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python bench.py --config cfg3 > gpurun_out/r2_cfg3_final.json 2> gpurun_out/r2_cfg3_final.err; echo "cfg3 rc=$?" >> gpurun_out/r2_cfg3_final.err
cat gpurun_out/r2_cfg3_final.json | cut -c1-1500; tail -3 gpurun_out/r2_cfg3_final.err
python bench.py --impl reference --config cfg3 > gpurun_out/r2_cfg3_reference.json 2>&1; cut -c1-600 gpurun_out/r2_cfg3_reference.json
