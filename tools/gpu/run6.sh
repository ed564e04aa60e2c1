set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dense_regime_gpu.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2_t6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t6.log
tail -3 gpurun_out/r2_t6.log
timeout 900 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench6.json 2> gpurun_out/r2_bench6.err; echo "bench rc=$?" >> gpurun_out/r2_bench6.err
cat gpurun_out/r2_bench6.json | cut -c1-2800; tail -5 gpurun_out/r2_bench6.err
timeout 900 python bench.py --config cfg4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_cfg4.json 2> gpurun_out/r2_cfg4.err; echo "cfg4 rc=$?" >> gpurun_out/r2_cfg4.err
cat gpurun_out/r2_cfg4.json | cut -c1-2500; tail -5 gpurun_out/r2_cfg4.err
