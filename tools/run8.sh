#!/bin/bash
# 8-GPU runs of the headline configurations (results land in gpurun_out/)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29521 bench.py --gpus 8 --steps 6 --warmup 3 2>gpurun_out/b8_1e7.err | tail -1 > gpurun_out/bench8_grav3d_1e7_v2.json
$TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 2 --particles 100000000 2>gpurun_out/b8_1e8.err | tail -1 > gpurun_out/bench8_grav3d_1e8_v2.json
$TR --master-port 29523 tools/bench_physics.py vort3d dualtree 10000000 1.4 2>gpurun_out/b8_v.err | tail -1 > gpurun_out/bench8_vort3d_dtt_v2.json
for f in gpurun_out/bench8_grav3d_1e7_v2.json gpurun_out/bench8_grav3d_1e8_v2.json; do python -c "
import json,sys
d=json.load(open('$f')); print(d['n_gpus'], d['config']['n_particles'], 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), {k: round(v,2) for k,v in d['phases_ms'].items()}, d['ms_steps'], d['e2e_ms_steps'])"; done
cat gpurun_out/bench8_vort3d_dtt_v2.json
$TR --master-port 29524 tools/check_multi.py 1000000 2>/dev/null | grep -c "True"; $TR --master-port 29525 tools/check_multi.py 1000000 2>/dev/null | grep CHECK_MULTI
