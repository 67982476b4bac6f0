set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 5 --no-extras > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; tail -c 900 gpurun_out/r2g_bench_n1.json; tail -3 gpurun_out/r2g_bench_n1.err
timeout 600 python bench.py --steps 3 --no-extras --small > gpurun_out/r2g_bench_small.json 2> gpurun_out/r2g_bench_small.err; tail -c 300 gpurun_out/r2g_bench_small.json; tail -3 gpurun_out/r2g_bench_small.err
