#!/bin/bash
# On the GPU box: compute-sanitizer memcheck and racecheck over the hand-written shared-memory / PTX paths
# (SURVEY.md §5 "race detection / sanitizers"): the smoke render (Path, Cornell box), the traversal-stack spill test,
# the sphere tests and the large-leaf (leaf table) tests. Logs under gpurun_out/<tag>_{memcheck,racecheck}_*.log; a
# summary line per run is printed. Usage: scripts/sanitize.sh [tag]
tag=${1:-sanitize}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name, tool, timeout, command...
  local name=$1 tool=$2 to=$3; shift 3
  local log=gpurun_out/${tag}_${tool}_${name}.log
  timeout $to $CS --tool $tool --print-limit 20 --error-exitcode 3 "$@" > $log 2>&1
  local rc=$?
  echo "$tool $name: exit $rc | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -E 'passed|failed|smoke ok' $log | tail -1)"
}
SPILL="tests/test_gpu_parity.py::test_deep_traversal_stack_spills_bit_exact"
LEAVES="tests/test_gpu_parity.py::test_leaf_sizes_counters_and_path_bit_exact"
SPHERE="tests/test_sphere.py"
WHITTED="tests/test_gpu_parity.py::test_whitted_deep_recursion"
for tool in memcheck racecheck; do
  run smoke $tool 900 python -c "import __graft_entry__ as g; g.smoke()"
  run spill $tool 1500 python -m pytest -x -q -m gpu $SPILL
  run sphere $tool 1500 python -m pytest -x -q -m gpu $SPHERE
  run leaves $tool 1500 python -m pytest -x -q -m gpu $LEAVES
  run whitted $tool 900 python -m pytest -x -q -m gpu $WHITTED
done
