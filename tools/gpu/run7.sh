set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "beyond_2_32 or edge_cases or medium_multi" > gpurun_out/r2_t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t7.log
tail -3 gpurun_out/r2_t7.log
timeout 900 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench7.json 2> gpurun_out/r2_bench7.err; echo "bench rc=$?" >> gpurun_out/r2_bench7.err
cat gpurun_out/r2_bench7.json | cut -c1-2800; tail -5 gpurun_out/r2_bench7.err
