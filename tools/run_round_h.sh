set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_h.log; cat gpurun_out/pytest_gpu_h.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_h.json 2> gpurun_out/bench_ref_h.err
python bench.py > gpurun_out/bench_n1_h.json 2> gpurun_out/bench_n1_h.err; tail -c 600 gpurun_out/bench_n1_h.json
python tools/next_probe.py 30 3 2>&1 | tee gpurun_out/next_probe.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_h.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lh.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tsd_kernel|qhsp_kernel|mask_fill_kernel|export_kernel" -c 5 -o gpurun_out/next_full -f python tools/next_probe.py 30 1 > gpurun_out/ncu_next.log 2>&1; tail -2 gpurun_out/ncu_next.log
