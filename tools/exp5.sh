#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/prof_trees2.py 10000000 | tail -1
python tools/prof_trees2.py 100000000 | tail -1
