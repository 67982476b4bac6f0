set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 240 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2w_pytest.log; cat gpurun_out/r2w_pytest.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2w_smoke.log 2>&1; tail -3 gpurun_out/r2w_smoke.log
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2w_bench_n1.json 2> gpurun_out/r2w_bench_n1.err; tail -c 900 gpurun_out/r2w_bench_n1.json; tail -3 gpurun_out/r2w_bench_n1.err
timeout 100 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_ref.json 2> gpurun_out/r2w_bench_ref.err; tail -c 400 gpurun_out/r2w_bench_ref.json
