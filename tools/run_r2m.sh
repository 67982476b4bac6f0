set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2m_pytest.log; cat gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 5 --no-extras > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err; tail -c 900 gpurun_out/r2m_bench_n1.json; tail -3 gpurun_out/r2m_bench_n1.err
timeout 900 python tools/recall_sweep.py > gpurun_out/r2m_recall.txt 2> gpurun_out/r2m_recall.err; tail -30 gpurun_out/r2m_recall.txt; tail -3 gpurun_out/r2m_recall.err
timeout 1200 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_n1_C4.json 2> gpurun_out/r2m_bench_n1_C4.err; tail -c 900 gpurun_out/r2m_bench_n1_C4.json; tail -3 gpurun_out/r2m_bench_n1_C4.err
timeout 1500 python bench.py --config C5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_n1_C5.json 2> gpurun_out/r2m_bench_n1_C5.err; tail -c 900 gpurun_out/r2m_bench_n1_C5.json; tail -5 gpurun_out/r2m_bench_n1_C5.err
