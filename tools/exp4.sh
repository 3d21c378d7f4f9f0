#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in 8192 4096; do echo "ONB_SUB $s"; ONB_SUB=$s python tools/prof_trees2.py 10000000 | tail -1; done
ONB_SUB=4096 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
