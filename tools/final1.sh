#!/bin/bash
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r1_1gpu_v3.json 2> gpurun_out/bench_err.log; tail -c 1500 gpurun_out/bench_r1_1gpu_v3.json | head -c 1200; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref_v3.json 2>> gpurun_out/bench_err.log; tail -c 600 gpurun_out/bench_r1_ref_v3.json; echo
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1_v3.csv python tools/prof_step.py 10000000 2 > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log
ncu --set full --clock-control none --import-source on -k regex:k_subtree -c 1 -f -o gpurun_out/prof_subtree_r1 python tools/prof_tree.py 10000000 1 > gpurun_out/ncu_sub.log 2>&1; tail -1 gpurun_out/ncu_sub.log
ncu --set full --clock-control none --import-source on -k regex:k_big_level -s 4 -c 1 -f -o gpurun_out/prof_biglevel_r1 python tools/prof_tree.py 10000000 1 > gpurun_out/ncu_big.log 2>&1; tail -1 gpurun_out/ncu_big.log
ls -la gpurun_out/*.ncu-rep
