set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
RB2_COLLOG= timeout 1500 python bench.py --config cfg3 --no-cpu-baseline > gpurun_out/r2_cfg3.json 2> gpurun_out/r2_cfg3.err; echo "cfg3 rc=$?" >> gpurun_out/r2_cfg3.err
cat gpurun_out/r2_cfg3.json | cut -c1-1500; tail -16 gpurun_out/r2_cfg3.err
timeout 900 python bench.py --config cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_cfg4.json 2> gpurun_out/r2_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/r2_cfg4.err
cat gpurun_out/r2_cfg4.json | cut -c1-1500; tail -5 gpurun_out/r2_cfg4.err
