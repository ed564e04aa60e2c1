set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log; tail -5 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?" >> gpurun_out/r2_bench_final.err
cut -c1-800 gpurun_out/r2_bench_final.json; tail -3 gpurun_out/r2_bench_final.err
timeout 1700 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?" >> gpurun_out/r2_bench_reference.err
cat gpurun_out/r2_bench_reference.json | cut -c1-1500; tail -3 gpurun_out/r2_bench_reference.err
