set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for c in 2 4 8; do
  RB2_BUILD_DIR=$GRAFT_REPO_ROOT/ropebwt2_b200/_build_cpl$c RB2_NVCC_EXTRA="-DFS_CPL=$c" python -c "from ropebwt2_b200 import build; print(build.build(force=True))" > gpurun_out/r2_build_cpl$c.log 2>&1
  grep -A2 "k_flat_merge" ropebwt2_b200/_build_cpl$c/build.log | grep -E "registers|spill" | head -4
  RB2_BUILD_DIR=$GRAFT_REPO_ROOT/ropebwt2_b200/_build_cpl$c timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-verify > gpurun_out/r2_bench8_cpl$c.json 2> gpurun_out/r2_bench8_cpl$c.err; echo "rc=$?"
  cut -c1-400 gpurun_out/r2_bench8_cpl$c.json
  RB2_BUILD_DIR=$GRAFT_REPO_ROOT/ropebwt2_b200/_build_cpl$c timeout 600 python tools/prof_multibatch.py 4 100000000 cfg3 > gpurun_out/r2_multib_cpl$c.log 2>&1; echo "rc=$?"
  tail -4 gpurun_out/r2_multib_cpl$c.log
done
