set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
# (1) launch list of the bench command itself (share of each kernel in the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2o_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2o_ncu_l.log 2>&1; tail -2 gpurun_out/r2o_ncu_l.log | cut -c1-400
# (2) --set full of every kernel of the path on C2 (index build + one round)
K='pack_kernel|minimizer_kernel|rs_hist|rs_scatter|bucket_bounds|sketch|seed_count|seed_fill|cluster_|rank_select|locus|extend16_kernel|traceback_kernel|chain_kernel|copy_filter|copy_select|ins_max|layout|colmap|msa_clean|hb_judge|compact'
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:"$K" -c 160 -o gpurun_out/r2o_full python tools/prof_c2.py --config C2 --rounds 1 > gpurun_out/r2o_ncu_f.log 2>&1; tail -2 gpurun_out/r2o_ncu_f.log | cut -c1-400
ncu -i gpurun_out/r2o_full.ncu-rep --page raw --csv > gpurun_out/r2o_full.csv 2>/dev/null
ls -l gpurun_out/r2o_full.ncu-rep
[ $(stat -c %s gpurun_out/r2o_full.ncu-rep) -gt 40000000 ] && rm gpurun_out/r2o_full.ncu-rep
# (3) small capture with source for the two dominant kernels
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'extend16_kernel|msa_clean' -c 4 -o gpurun_out/r2o_top python tools/prof_c2.py --config C2 --rounds 1 > gpurun_out/r2o_ncu_t.log 2>&1; tail -2 gpurun_out/r2o_ncu_t.log | cut -c1-300
ls -l gpurun_out/
