#!/bin/bash
# On the GPU box: the bench line, then the ncu launch list and one --set full capture of the dominant kernel for the same command
# (each ncu pass only after the plain command exited 0). Outputs under gpurun_out/<tag>_*.
tag=${1:-bench}
mkdir -p gpurun_out
python bench.py > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err || { tail -20 gpurun_out/${tag}.err; exit 1; }
cat gpurun_out/${tag}.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_reference.json 2>> gpurun_out/${tag}.err
cat gpurun_out/${tag}_reference.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_short.json 2>> gpurun_out/${tag}.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_list.log 2>&1
python scripts/launch_shares.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_shares.txt 2>&1; cat gpurun_out/${tag}_shares.txt
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 8 -c 8 -f -o gpurun_out/${tag}_closest \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_closest.ncu-rep --page raw --csv > gpurun_out/${tag}_closest_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_closest_raw.csv > gpurun_out/${tag}_closest_summary.txt 2>&1
