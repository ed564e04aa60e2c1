set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_regime_gpu.py tests/test_parity_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_t4.log
tail -3 gpurun_out/r2_t4.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?" >> gpurun_out/r2_bench4.err
cat gpurun_out/r2_bench4.json | cut -c1-2500; tail -5 gpurun_out/r2_bench4.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches4.csv python tools/prof_bench_scale.py > gpurun_out/r2_launches4.out 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_flat_merge" -s 90 -c 1 -o gpurun_out/r2_flat_merge4_c90 -f python tools/prof_bench_scale.py > gpurun_out/r2_ncu_full4.out 2>&1
ls -la gpurun_out | tail -6
