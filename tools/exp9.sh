#!/bin/bash
for cfg in "1184 4736" "4736 20000" "2400 20000" "1184 20000" "4736 40000" "10000 70000"; do set -- $cfg
echo "TPT1_MAX $1 TPT2_MAX $2"; ONB_TPT1_MAX=$1 ONB_TPT2_MAX=$2 ONB_DTT_PROF=1 python tools/prof_step.py 10000000 3 2>&1 | tail -11 | awk '/dtt level/{printf "%s:%s ", $3, $9} /step 2/{print ""; print $0}'
done
