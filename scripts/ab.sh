#!/bin/bash
# usage (GPU box): scripts/ab.sh <variant> ...   runs perf_probe.py ab for each build/variants/libyuki_<variant>.so (one pipe)
export YK_PIPES=${YK_PIPES:-1}
for v in "$@"; do
  echo "== $v"
  YUKI_GPU_LIB=$PWD/build/variants/libyuki_$v.so python scripts/perf_probe.py ab 2>&1 | sed 's/scene [0-9.]*s tris [0-9]* nodes [0-9]* | //; s/Mrays\/s(closest) [0-9.]* Mrays\/s(total) //; s/ | launches.*roofline/ | roofline/'
done
