set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python bench.py --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc=$?" >> gpurun_out/r2_bench11.err
cut -c1-600 gpurun_out/r2_bench11.json; tail -2 gpurun_out/r2_bench11.err
timeout 600 python tools/prof_multibatch.py 4 100000000 cfg3 > gpurun_out/r2_multib11.log 2>&1; tail -4 gpurun_out/r2_multib11.log
