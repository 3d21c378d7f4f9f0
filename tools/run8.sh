#!/bin/bash
# 8-GPU runs of the BASELINE.json configurations (results land in gpurun_out/)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench8_grav3d_1e7.json
$TR --master-port 29522 tools/bench_physics.py vort3d dualtree 10000000 1.4 2>/dev/null | tail -1 > gpurun_out/bench8_vort3d_dtt.json
$TR --master-port 29523 tools/bench_physics.py vort3d boxwise 10000000 1.4 2>/dev/null | tail -1 > gpurun_out/bench8_vort3d_box.json
$TR --master-port 29524 tools/bench_physics.py vortgrad3d boxwise 10000000 1.4 2>/dev/null | tail -1 > gpurun_out/bench8_vortgrad3d_box.json
cat gpurun_out/bench8_*.json | cut -c1-900
