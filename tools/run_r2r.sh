set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2r_pytest.log; cat gpurun_out/r2r_pytest.log
timeout 600 python bench.py --steps 5 --no-extras > gpurun_out/r2r_bench_n1.json 2> gpurun_out/r2r_bench_n1.err; tail -c 900 gpurun_out/r2r_bench_n1.json; tail -3 gpurun_out/r2r_bench_n1.err
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2r_launches_C5s8.csv python tools/prof_c2.py --config C5 --shard-of 8 --rounds 2 > gpurun_out/r2r_ncu_l.log 2>&1; tail -4 gpurun_out/r2r_ncu_l.log | cut -c1-300
