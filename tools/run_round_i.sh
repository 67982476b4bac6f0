set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_i.log; cat gpurun_out/pytest_gpu_i.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_i.json 2> gpurun_out/bench_ref_i.err
python bench.py > gpurun_out/bench_n1_i.json 2> gpurun_out/bench_n1_i.err; tail -c 400 gpurun_out/bench_n1_i.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_i.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_li.log 2>&1
