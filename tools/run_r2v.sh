set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_coarse_gpu.py tests/test_dist_gpu.py tests/test_mask_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2v_pytest.log; cat gpurun_out/r2v_pytest.log
timeout 600 python bench.py --steps 5 --no-extras > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; tail -c 700 gpurun_out/r2v_bench_n1.json; tail -3 gpurun_out/r2v_bench_n1.err
# (1) launch list of the bench command itself (share of each kernel in the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2v_ncu_l.log 2>&1; tail -2 gpurun_out/r2v_ncu_l.log | cut -c1-300
# (2) --set full of every kernel of the path on C2 (index build + one round)
K='pack_kernel|minimizer_kernel|rs_hist|rs_scatter|bucket_bounds|seed_count|seed_fill|cluster_|rank_select|extend16_kernel|traceback_kernel|chain_kernel|copy_filter|copy_select|ins_max|layout|colmap|msa_clean|hb_judge|cons_compact'
timeout 1200 ncu --set full --import-source on --clock-control none -k regex:"$K" -c 90 -o gpurun_out/r2v_full python tools/prof_c2.py --config C2 --rounds 1 > gpurun_out/r2v_ncu_f.log 2>&1; tail -2 gpurun_out/r2v_ncu_f.log | cut -c1-300
ncu -i gpurun_out/r2v_full.ncu-rep --page raw --csv > gpurun_out/r2v_full.csv 2>/dev/null
ls -l gpurun_out/r2v_full.ncu-rep
[ $(stat -c %s gpurun_out/r2v_full.ncu-rep) -gt 45000000 ] && rm gpurun_out/r2v_full.ncu-rep
ls gpurun_out | head -5
