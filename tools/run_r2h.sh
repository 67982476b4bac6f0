set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
timeout 900 python tools/index_bench.py --mbp 100 800 2300 --out gpurun_out/r2h_index_bench.json > gpurun_out/r2h_index_bench.log 2>&1; cat gpurun_out/r2h_index_bench.log
timeout 600 ncu --set full --clock-control none -k regex:'pack_kernel|minimizer_kernel|rs_hist|rs_scatter|bucket_count' -c 14 -o gpurun_out/r2h_index_full python tools/index_bench.py --mbp 100 > gpurun_out/r2h_ncu.log 2>&1; tail -2 gpurun_out/r2h_ncu.log
ncu -i gpurun_out/r2h_index_full.ncu-rep --page raw --csv > gpurun_out/r2h_index_full.csv 2>/dev/null
[ $(stat -c %s gpurun_out/r2h_index_full.ncu-rep) -gt 30000000 ] && rm gpurun_out/r2h_index_full.ncu-rep
timeout 600 python bench.py --steps 5 --no-extras > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; tail -c 600 gpurun_out/r2h_bench_n1.json
