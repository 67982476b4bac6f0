set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_coarse_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2x_pytest.log; cat gpurun_out/r2x_pytest.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/r2x_bench_n1.json 2> gpurun_out/r2x_bench_n1.err; tail -c 1500 gpurun_out/r2x_bench_n1.json; tail -3 gpurun_out/r2x_bench_n1.err
timeout 120 python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/r2x_bench_n1_C3.json 2> gpurun_out/r2x_bench_n1_C3.err; tail -c 300 gpurun_out/r2x_bench_n1_C3.json; tail -3 gpurun_out/r2x_bench_n1_C3.err
