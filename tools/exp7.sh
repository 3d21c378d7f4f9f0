#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r1_1gpu_v2.json 2> gpurun_out/bench_err.log; tail -c 3000 gpurun_out/bench_r1_1gpu_v2.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1_v2.csv python tools/prof_step.py 10000000 2 > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log
