set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python bench.py --config cfg3 --no-cpu-baseline > gpurun_out/r2_cfg3c.json 2> gpurun_out/r2_cfg3c.err; echo "cfg3 rc=$?" >> gpurun_out/r2_cfg3c.err
cat gpurun_out/r2_cfg3c.json | cut -c1-1200; tail -3 gpurun_out/r2_cfg3c.err
timeout 900 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-verify > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?" >> gpurun_out/r2_bench9.err
cat gpurun_out/r2_bench9.json | cut -c1-1000; tail -3 gpurun_out/r2_bench9.err
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_column_fused" -s 40 -c 1 -o gpurun_out/r2_fused_c60 -f python tools/prof_bench_scale.py > gpurun_out/r2_ncu_fused.out 2>&1
ls -la gpurun_out | tail -4
