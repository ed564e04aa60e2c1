set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t18.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t18.log; tail -3 gpurun_out/r2_t18.log
RB2_ASYNC=0 timeout 600 python -m pytest tests/test_dense_regime_gpu.py tests/test_parity_gpu.py -m gpu -x -q -k "not beyond_2_32 and not one_long" > gpurun_out/r2_t18_sync.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t18_sync.log; tail -2 gpurun_out/r2_t18_sync.log
RB2_FUSED=0 RB2_WIDE_RATIO=1 timeout 600 python -m pytest tests/test_dense_regime_gpu.py -m gpu -x -q > gpurun_out/r2_t18_nofused.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t18_nofused.log; tail -2 gpurun_out/r2_t18_nofused.log
for r in 96 384; do
  RB2_WIDE_RATIO=$r timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r2_bench18_w$r.json 2> gpurun_out/r2_bench18_w$r.err
  cut -c1-300 gpurun_out/r2_bench18_w$r.json
  RB2_WIDE_RATIO=$r timeout 600 python tools/prof_multibatch.py 4 100000000 cfg3 > gpurun_out/r2_multib18_w$r.log 2>&1; tail -2 gpurun_out/r2_multib18_w$r.log
done
