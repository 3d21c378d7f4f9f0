#!/bin/bash
for per in 0.1 0.5 5; do
echo "clock period $per"
ONB_CLOCK_PERIOD=$per python bench.py --steps 10 --warmup 3 --no-cpu-baseline | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['ms_steps'], d['e2e_ms_steps'], d['clocks'])"
done
