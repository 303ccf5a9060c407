#!/bin/bash
# A/B of the inverse-CDF node tables on one box: B0 -> K*0(-> K pi) gamma kernel times with the tables (default) and with
# the solver / library inverse (PS_B200_NO_NODE_TABLES=1), alternating.   tools/ab_node_tables.sh [n_events] [reps]
N=${1:-100000000}; R=${2:-10}
for round in 1 2; do
  PS_B200_LIB= python tools/kbench_kstar.py $N $R | sed 's/^default/tables /'
  PS_B200_NO_NODE_TABLES=1 python tools/kbench_kstar.py $N $R | sed 's/^default/no tables/'
done
