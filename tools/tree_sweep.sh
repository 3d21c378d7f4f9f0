#!/bin/bash
for b in 1 2 4 8; do echo "blocks/SM $b"; ONB_BIG_BLOCKS_PER_SM=$b python tools/prof_tree.py 10000000 2 | tail -1 | cut -c1-16; done
for n in 8192 32768 65536 131072 262144; do echo "BIG_NODE $n"; ONB_BIG_NODE=$n python tools/prof_tree.py 10000000 2 | tail -1 | cut -c1-16; done
