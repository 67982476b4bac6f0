set -x
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
run() { n=$1; cfg=$2; sc=$3; st=$4; tag=$5
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $n --steps $st --warmup 3 --config $cfg --scaling $sc > gpurun_out/r2n_bench_n${n}_${tag}.json 2> gpurun_out/r2n_bench_n${n}_${tag}.err
  tail -c 500 gpurun_out/r2n_bench_n${n}_${tag}.json; tail -2 gpurun_out/r2n_bench_n${n}_${tag}.err | cut -c1-300; }
run 8 C5 strong 3 C5s
run 8 C4 strong 5 C4s
run 8 C2 weak 10 C2w
