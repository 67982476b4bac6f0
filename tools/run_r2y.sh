set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 200 python -m pytest tests/test_coarse_gpu.py tests/test_dist_gpu.py tests/test_boundary_gpu.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2y_pytest.log; cat gpurun_out/r2y_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2y_bench_n2.json 2> gpurun_out/r2y_bench_n2.err; tail -c 1800 gpurun_out/r2y_bench_n2.json; tail -3 gpurun_out/r2y_bench_n2.err | cut -c1-300
