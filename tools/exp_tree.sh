#!/bin/bash
echo "--- step, concurrent"; python tools/prof_step.py 10000000 4 | tail -3
echo "--- step, sequential"; ONB_SEQ_BUILDS=1 python tools/prof_step.py 10000000 4 | tail -3
echo "--- step, concurrent, 1 block/SM"; ONB_BIG_BLOCKS_PER_SM_CONC=1 python tools/prof_step.py 10000000 4 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'], d['e2e']['ms_per_step'], d['ms_steps'])"
