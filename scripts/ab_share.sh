#!/bin/bash
# A/B (GPU box): co-resident kernels of two pipes — persistent grids capped below full occupancy so that the other pipe's
# kernel fits beside them. usage: scripts/ab_share.sh <probe> ; prints one line per (pipes, trace blocks/SM, wide blocks/SM)
probe=${1:-c64}
for cfg in "2 0 0" "2 4 4" "2 4 8" "2 6 4" "2 4 16" "3 4 4" "4 4 4" "2 5 6" "1 0 0"; do
  set -- $cfg
  export YK_PIPES=$1
  if [ "$2" != "0" ]; then export YK_TRACE_PER_SM=$2; else unset YK_TRACE_PER_SM; fi
  if [ "$3" != "0" ]; then export YK_WIDE_PER_SM=$3; else unset YK_WIDE_PER_SM; fi
  echo -n "pipes $1 trace/SM $2 wide/SM $3 | "
  YK_STAGE_TIMING=0 python scripts/perf_probe.py $probe 2>&1 | sed 's/scene [0-9.]*s tris [0-9]* nodes [0-9]* | //; s/ | launches.*//' | tail -1
done
