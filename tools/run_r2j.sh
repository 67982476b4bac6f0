set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_coarse_gpu.py tests/test_mask_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r2j_pytest.log; cat gpurun_out/r2j_pytest.log
free -g | head -2
( while true; do nvidia-smi --query-gpu=memory.used --format=csv,noheader; sleep 20; done ) > gpurun_out/r2j_mem.log 2>&1 &
MON=$!
timeout 1500 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2j_bench_n1_C5.json 2> gpurun_out/r2j_bench_n1_C5.err; tail -c 1200 gpurun_out/r2j_bench_n1_C5.json; tail -5 gpurun_out/r2j_bench_n1_C5.err
kill $MON
sort -n gpurun_out/r2j_mem.log | tail -1
