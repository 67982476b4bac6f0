set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2f_pytest.log; cat gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 5 --no-extras > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; tail -c 900 gpurun_out/r2f_bench_n1.json; tail -3 gpurun_out/r2f_bench_n1.err
timeout 900 python tools/recall_sweep.py --out gpurun_out/r2f_recall.md > gpurun_out/r2f_recall.log 2>&1; cat gpurun_out/r2f_recall.md; tail -3 gpurun_out/r2f_recall.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches.csv python tools/prof_c2.py --config C2 --rounds 2 > gpurun_out/r2f_ncu_l.log 2>&1; tail -2 gpurun_out/r2f_ncu_l.log
