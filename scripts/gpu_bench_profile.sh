#!/bin/bash
# On the GPU box: the bench lines (both arms), then the ncu launch list of the same command and `--set full` captures of the
# dominant kernel for the two workloads the line reports a roofline for (each ncu pass only after the plain command exited 0):
#   c2 = bench.py's main leg (Cornell box), c5 = the large-scene leg's render (scripts/perf_probe.py terrain16: same scene,
#   film, sampler, integrator, one pipe).
# Outputs under gpurun_out/<tag>_*; the raw CSVs and kernel_hash.txt come back so that scripts/ncu_traffic.py (run in the
# repository, where git knows the commit) can stamp profiles/<round>/traffic.json.
tag=${1:-bench}
mkdir -p gpurun_out
python -c "import bench; print(bench.kernel_source_hash())" > gpurun_out/${tag}_kernel_hash.txt
python bench.py > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err || { tail -20 gpurun_out/${tag}.err; exit 1; }
cat gpurun_out/${tag}.json
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_reference.json 2>> gpurun_out/${tag}.err
cat gpurun_out/${tag}_reference.json
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-large-scene > gpurun_out/${tag}_short.json 2>> gpurun_out/${tag}.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-large-scene > gpurun_out/${tag}_ncu_list.log 2>&1
python scripts/launch_shares.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_shares.txt 2>&1; cat gpurun_out/${tag}_shares.txt
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 8 -c 8 -f -o gpurun_out/${tag}_closest \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-large-scene > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_closest.ncu-rep --page raw --csv > gpurun_out/${tag}_closest_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_closest_raw.csv > gpurun_out/${tag}_closest_summary.txt 2>&1
python scripts/perf_probe.py terrain16 > gpurun_out/${tag}_c5_plain.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:k_trace_closest -s 0 -c 8 -f -o gpurun_out/${tag}_c5_closest \
    python scripts/perf_probe.py terrain16 > gpurun_out/${tag}_c5_ncu_full.log 2>&1
ncu -i gpurun_out/${tag}_c5_closest.ncu-rep --page raw --csv > gpurun_out/${tag}_c5_closest_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/${tag}_c5_closest_raw.csv > gpurun_out/${tag}_c5_closest_summary.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_c5_launches.csv \
    python scripts/perf_probe.py terrain16 > gpurun_out/${tag}_c5_ncu_list.log 2>&1
python scripts/launch_shares.py gpurun_out/${tag}_c5_launches.csv > gpurun_out/${tag}_c5_shares.txt 2>&1; cat gpurun_out/${tag}_c5_shares.txt
rm -f gpurun_out/${tag}_closest.ncu-rep gpurun_out/${tag}_c5_closest.ncu-rep   # gpurun brings back at most 64 MiB
