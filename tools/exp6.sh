#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ONB_BIG_PROF=1 python tools/prof_tree.py 10000000 2 2>&1 | tail -10 | cut -c1-420
python tools/prof_trees2.py 10000000 | tail -1
