set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt
timeout 900 python -m pytest tests/test_dense_regime_gpu.py tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2_t1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t1.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?" >> gpurun_out/r2_bench1.err
tail -3 gpurun_out/r2_t1.log; cat gpurun_out/r2_bench1.json | cut -c1-1500; tail -5 gpurun_out/r2_bench1.err
