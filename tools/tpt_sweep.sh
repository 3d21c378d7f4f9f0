#!/bin/bash
for t in 2 4; do
  ONB_P2P_TPT=$t python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 > /tmp/tpt_$t.json
  python -c "import json; r=json.load(open('/tmp/tpt_$t.json')); print('tpt $t', r['ms_per_step'], r['phases_ms']['p2p'], r['roofline']['frac'])"
done
