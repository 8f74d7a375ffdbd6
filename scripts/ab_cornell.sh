#!/bin/bash
# usage (GPU box): scripts/ab_cornell.sh <variant> ...   the Cornell-box line of perf_probe (one pipe) for each build/variants/libyuki_<variant>.so, twice
export YK_PIPES=${YK_PIPES:-1}
for rep in 1 2; do
for v in "$@"; do
  echo "== $v"
  YUKI_GPU_LIB=$PWD/build/variants/libyuki_$v.so python scripts/perf_probe.py c64 2>&1 | sed 's/scene [0-9.]*s tris [0-9]* nodes [0-9]* | //; s/Mrays\/s(closest) [0-9.]* Mrays\/s(total) //; s/ | launches.*roofline/ | roofline/'
done
done
