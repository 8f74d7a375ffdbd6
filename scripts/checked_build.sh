#!/bin/bash
# Builds build/variants/libyuki_checked.so: the library with device-side bounds assertions on (-DYK_CHECKED, csrc/wf_common.cuh).
# On the GPU box: YUKI_GPU_LIB=$PWD/build/variants/libyuki_checked.so python -m pytest tests -q -m gpu
# (compute-sanitizer is closed on this pool; this is the substitute for its memcheck pass. Races: the traversal stack is
# s_stack[level][thread] — a thread only ever touches its own column — and the block-aggregated queue appends are separated by
# __syncthreads, so there is no inter-thread shared-memory hand-over to race on.)
cd "$(dirname "$0")/.."
bash scripts/build_variant.sh checked -DYK_CHECKED
