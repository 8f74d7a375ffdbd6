#!/bin/bash
# On the GPU box: `ncu --set full` captures of the kernels that dominate the multi-material room (BASELINE.json configs[3]):
# every k_shade<kind>, k_trace_shadow, k_classify, k_raygen on the room, k_tree_return on a Whitted render, and the
# closest-hit kernel on the 10 M-triangle terrain with the ray sort off and on. Each capture runs only after the plain command
# exited 0. Outputs under gpurun_out/<tag>_*. usage: scripts/gpu_prof_room.sh [tag]
tag=${1:-r02}
mkdir -p gpurun_out
cap() {  # name, probe-config, kernel-regex, skip, count, [env]
  local name=$1 cfg=$2 kre=$3 skip=$4 cnt=$5
  python scripts/perf_probe.py $cfg > gpurun_out/${tag}_${name}_plain.log 2>&1 || { echo "$name: plain run failed"; tail -5 gpurun_out/${tag}_${name}_plain.log; return; }
  ncu --set full --import-source on --clock-control none -k regex:$kre -s $skip -c $cnt -f -o gpurun_out/${tag}_${name} \
      python scripts/perf_probe.py $cfg > gpurun_out/${tag}_${name}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_${name}.ncu-rep --page raw --csv > gpurun_out/${tag}_${name}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/${tag}_${name}_raw.csv > gpurun_out/${tag}_${name}_summary.txt 2>&1
  # gpurun brings back at most 64 MiB: keep the summaries, drop the reports (KEEP_REP=name keeps one, with its source page)
  if [ "$KEEP_REP" = "$name" ]; then
    ncu -i gpurun_out/${tag}_${name}.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${tag}_${name}_source.csv 2>/dev/null
    python scripts/ncu_lines.py gpurun_out/${tag}_${name}_source.csv 0.7 > gpurun_out/${tag}_${name}_lines.txt 2>&1
    gzip -f gpurun_out/${tag}_${name}_source.csv
  fi
  rm -f gpurun_out/${tag}_${name}.ncu-rep gpurun_out/${tag}_${name}_raw.csv
  echo "$name: $(grep -c '^----' gpurun_out/${tag}_${name}_summary.txt) launches captured"
}
# room4: one batch of the room at 4 spp; launches per bounce: closest, classify, 4 x shade, shadow (+3 with the sort)
cap shade_room room4 k_shade 4 8          # bounces 1 and 2, all four kinds
cap shadow_room room4 k_trace_shadow 0 3
cap classify_room room4 k_classify 0 2
cap raygen_room room4 k_raygen 0 1
cap closest_room room4 k_trace_closest 0 3
cap tree_return whitted1 k_tree_return 1 2
cap closest_terrain terrain1 k_trace_closest 1 2
YK_SORT_KEY=1 cap closest_terrain_sorted terrain1 k_trace_closest 1 2
cap shadow_terrain terrain1 k_trace_shadow 0 2
rm -f gpurun_out/${tag}_*_raw.csv.tmp
ls -la gpurun_out | tail -30
