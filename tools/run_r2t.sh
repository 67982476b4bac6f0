set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2t_pytest.log; cat gpurun_out/r2t_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1; tail -3 gpurun_out/r2t_smoke.log
timeout 900 python bench.py > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; tail -c 600 gpurun_out/r2t_bench_n1.json; tail -3 gpurun_out/r2t_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2t_bench_ref.json 2> gpurun_out/r2t_bench_ref.err; tail -c 600 gpurun_out/r2t_bench_ref.json; tail -3 gpurun_out/r2t_bench_ref.err
timeout 900 python bench.py --config C3 --steps 5 > gpurun_out/r2t_bench_n1_C3.json 2> gpurun_out/r2t_bench_n1_C3.err; tail -c 600 gpurun_out/r2t_bench_n1_C3.json; tail -3 gpurun_out/r2t_bench_n1_C3.err
