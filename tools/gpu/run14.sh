set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_regime_gpu.py -m gpu -x -q > gpurun_out/r2_t14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t14.log; tail -3 gpurun_out/r2_t14.log
timeout 900 python bench.py --config cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_cfg4b.json 2> gpurun_out/r2_cfg4b.err; echo "cfg4 rc=$?" >> gpurun_out/r2_cfg4b.err
cat gpurun_out/r2_cfg4b.json | cut -c1-1800; tail -3 gpurun_out/r2_cfg4b.err
