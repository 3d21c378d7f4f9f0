#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in 0 32768 65536 131072; do
echo "SPLIT_TARGET $t"; ONB_P2P_SPLIT_TARGET=$t ONB_DTT_PROF=1 python tools/prof_step.py 10000000 3 2>&1 | tail -13 | awk '/dtt level/{printf "%s:%s ", $3, $9} /step 2/{print ""; print $0}'
done
