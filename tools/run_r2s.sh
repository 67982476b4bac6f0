set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2s_pytest.log; cat gpurun_out/r2s_pytest.log
run() { n=$1; cfg=$2; sc=$3; st=$4; tag=$5; shift 5
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $n --steps $st --warmup 3 --config $cfg --scaling $sc --coarse off "$@" > gpurun_out/r2s_bench_n${n}_${tag}.json 2> gpurun_out/r2s_bench_n${n}_${tag}.err
  tail -c 1500 gpurun_out/r2s_bench_n${n}_${tag}.json; tail -2 gpurun_out/r2s_bench_n${n}_${tag}.err | cut -c1-300; }
run 2 C2 weak 10 C2w_thr_off --gather-thread off
run 2 C2 weak 10 C2w_thr_on --gather-thread on
