set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 120 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; tail -c 400 gpurun_out/r2z_bench_n1.json; tail -3 gpurun_out/r2z_bench_n1.err
# launch list of the bench command itself (share of each kernel in the step); numbers under ncu are not bench values
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --strong-leg off > gpurun_out/r2z_ncu_l.log 2>&1; tail -2 gpurun_out/r2z_ncu_l.log | cut -c1-300
ls -l gpurun_out | head
