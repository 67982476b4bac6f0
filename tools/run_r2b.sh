set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2b_smi.txt
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pytest.log
for cfg in "C2 weak" "C4 strong"; do set -- $cfg
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --config $1 --scaling $2 > gpurun_out/r2b_bench_n2_$1.json 2> gpurun_out/r2b_bench_n2_$1.err; tail -c 600 gpurun_out/r2b_bench_n2_$1.json; tail -3 gpurun_out/r2b_bench_n2_$1.err
done
timeout 900 python bench.py --config C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_n1_C4.json 2> gpurun_out/r2b_bench_n1_C4.err; tail -c 600 gpurun_out/r2b_bench_n1_C4.json
timeout 900 python bench.py --config C3 --steps 5 --warmup 3 > gpurun_out/r2b_bench_n1_C3.json 2> gpurun_out/r2b_bench_n1_C3.err; tail -c 600 gpurun_out/r2b_bench_n1_C3.json; tail -3 gpurun_out/r2b_bench_n1_C3.err
