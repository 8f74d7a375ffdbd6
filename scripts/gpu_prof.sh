#!/bin/bash
# usage (on the GPU box): scripts/gpu_prof.sh <tag> <probe-config> <kernel-regex> [skip] [count]
# 1) plain timing run, 2) ncu --set full of the named kernel, 3) launch list. Outputs under gpurun_out/<tag>_*.
tag=$1; cfg=${2:-cornell16}; kre=${3:-k_trace}; skip=${4:-6}; cnt=${5:-2}
mkdir -p gpurun_out
python scripts/perf_probe.py $cfg > gpurun_out/${tag}_plain.log 2>&1 || { tail -20 gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log
ncu --set full --import-source on --clock-control none -k regex:$kre -s $skip -c $cnt -f -o gpurun_out/${tag}_full \
    python scripts/perf_probe.py $cfg > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_full_summary.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python scripts/perf_probe.py $cfg > gpurun_out/${tag}_ncu_list.log 2>&1
python scripts/launch_shares.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_shares.txt 2>&1
cat gpurun_out/${tag}_shares.txt
