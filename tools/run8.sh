#!/bin/bash
# Round-2 multi-GPU session on one 8xB200 box: correctness over NCCL first, then the scaling points and BASELINE configs[4].
# Every command has its own timeout (a hung collective must not eat the box's limit).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nproc; free -g | head -2; df -h /dev/shm | tail -1; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
run() { n=$1; shift; timeout "$TMO" python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) "$@"; }
TMO=200
run 8 tools/check_multi.py 1000000 2>&1 | grep "CHECK_MULTI\|False\|rror" | tail -5
run 8 tools/check_multi.py 3000000 grav3d lean 2>&1 | grep "CHECK_MULTI\|False\|rror" | tail -5
TMO=240
run 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_8gpu_n1e7.json 2> gpurun_out/r2_bench_8gpu_n1e7.err; echo "8gpu 1e7 rc=$?"
run 4 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2_bench_4gpu_n1e7.json 2> gpurun_out/r2_bench_4gpu_n1e7.err; echo "4gpu 1e7 rc=$?"
TMO=400
run 8 bench.py --gpus 8 --particles 100000000 --steps 3 --warmup 2 --check-error > gpurun_out/r2_bench_8gpu_n1e8.json 2> gpurun_out/r2_bench_8gpu_n1e8.err; echo "8gpu 1e8 rc=$?"
TMO=900
run 8 bench.py --gpus 8 --particles 1000000000 --steps 2 --warmup 1 > gpurun_out/r2_bench_8gpu_n1e9.json 2> gpurun_out/r2_bench_8gpu_n1e9.err; echo "8gpu 1e9 rc=$?"
tail -4 gpurun_out/r2_bench_8gpu_n1e9.err
TMO=240
: > gpurun_out/r2_bench_8gpu_other_physics.jsonl
run 8 tools/bench_physics.py vort3d dualtree 10000000 1.4 | grep "^{" >> gpurun_out/r2_bench_8gpu_other_physics.jsonl
run 8 tools/bench_physics.py vort3d boxwise 10000000 1.4 | grep "^{" >> gpurun_out/r2_bench_8gpu_other_physics.jsonl
run 8 tools/bench_physics.py vortgrad3d boxwise 10000000 1.4 | grep "^{" >> gpurun_out/r2_bench_8gpu_other_physics.jsonl
cat gpurun_out/r2_bench_8gpu_other_physics.jsonl | cut -c1-400
timeout 300 ./onbody_b200/bin/ongrav3d -g=8 -n=100000000 -t=1.4 -o=4 > gpurun_out/r2_driver_8gpu_n1e8.txt 2>&1; echo "driver rc=$?"; tail -8 gpurun_out/r2_driver_8gpu_n1e8.txt
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*gpu_n1e*.json")):
    for line in open(f):
        if line.startswith("{"):
            r = json.loads(line)
            print(f, "ms/step %.2f e2e %.2f" % (r["ms_per_step"], r["e2e"]["ms_per_step"]), {k: round(v, 2) for k, v in r["phases_ms"].items()}, "mem %.1f GB" % (r["device_memory_peak_bytes_per_gpu"] / 1e9), r.get("accuracy"), r.get("exchange"))
PY
