set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2q_launches.csv python tools/prof_c2.py --config C2 --rounds 2 > gpurun_out/r2q_ncu_l.log 2>&1; tail -2 gpurun_out/r2q_ncu_l.log | cut -c1-300
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'msa_clean|hb_judge_[abcd]|traceback_kernel|copy_select' -c 7 -o gpurun_out/r2q_full python tools/prof_c2.py --config C2 --rounds 1 > gpurun_out/r2q_ncu_f.log 2>&1; tail -2 gpurun_out/r2q_ncu_f.log | cut -c1-300
ncu -i gpurun_out/r2q_full.ncu-rep --page raw --csv > gpurun_out/r2q_full.csv 2>/dev/null
ls -l gpurun_out/r2q_full.ncu-rep
