set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_pipeline_gpu.py tests/test_itr_gpu.py tests/test_coarse_gpu.py -m gpu -q -x -k "chained or itr or sharded or tie_rule" > gpurun_out/r2e_sanitizer.log 2>&1; tail -15 gpurun_out/r2e_sanitizer.log
