set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -c 1500 gpurun_out/r2a_bench_n1.json; tail -5 gpurun_out/r2a_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2a_launches.csv python tools/prof_c2.py --config C2 --rounds 2 > gpurun_out/r2a_ncu_l.log 2>&1; tail -2 gpurun_out/r2a_ncu_l.log
timeout 900 ncu --set full --clock-control none -k regex:'pack_kernel|minimizer_kernel|rs_hist|rs_scatter|seed_count_kernel|seed_fill_kernel|rank_select_kernel|extend16_kernel|traceback_kernel|msa_write_kernel|hb_clean_kernel|ins_max_kernel|copy_select_kernel|hb_judge' -c 70 -o gpurun_out/r2a_full python tools/prof_c2.py --config C2 > gpurun_out/r2a_ncu_f.log 2>&1; tail -2 gpurun_out/r2a_ncu_f.log
ncu -i gpurun_out/r2a_full.ncu-rep --page raw --csv > gpurun_out/r2a_full.csv 2>/dev/null; ls -la gpurun_out/r2a_full*
[ $(stat -c %s gpurun_out/r2a_full.ncu-rep) -gt 40000000 ] && rm gpurun_out/r2a_full.ncu-rep
timeout 600 python tools/recall_sweep.py --out gpurun_out/r2a_recall.md > gpurun_out/r2a_recall.log 2>&1; cat gpurun_out/r2a_recall.md
