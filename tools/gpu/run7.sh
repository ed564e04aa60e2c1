set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dense_regime_gpu.py -m gpu -x -q -k "beyond_2_32 or edge_cases or medium_multi or pipelined or resident or forced_dense" > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -3 gpurun_out/r2_t7.log
timeout 900 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?" >> gpurun_out/r2_bench7.err
cat gpurun_out/r2_bench7.json | cut -c1-1200; tail -5 gpurun_out/r2_bench7.err
timeout 300 python bench.py --config cfg3 --reads 2400000 --no-cpu-baseline > gpurun_out/r2_cfg3_small.json 2> gpurun_out/r2_cfg3_small.err; echo "cfg3 small rc=$?" >> gpurun_out/r2_cfg3_small.err
tail -3 gpurun_out/r2_cfg3_small.err
timeout 1500 python bench.py --config cfg3 --no-cpu-baseline > gpurun_out/r2_cfg3b.json 2> gpurun_out/r2_cfg3b.err; echo "cfg3 rc=$?" >> gpurun_out/r2_cfg3b.err
cat gpurun_out/r2_cfg3b.json | cut -c1-1500; tail -5 gpurun_out/r2_cfg3b.err
