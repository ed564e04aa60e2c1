set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_flat_merge" -s 90 -c 1 -o gpurun_out/r2_flat_merge12_c90 -f python tools/prof_bench_scale.py > gpurun_out/r2_ncu_full12.out 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches12.csv python tools/prof_bench_scale.py > gpurun_out/r2_launches12.out 2>&1
ls -la gpurun_out | tail -3
