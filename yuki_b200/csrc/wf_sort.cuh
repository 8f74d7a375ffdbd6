// Ray-queue sort ("a ray-queue sort/compaction pass", BASELINE.json north_star): bounce rays leave the shading kernels in
// the order of the pixels they came from, with directions drawn from a hemisphere — neighbouring lanes walk unrelated parts
// of the BVH (profiles/r01: 15-18 of 32 lanes, L1 hit 40-52 % on the 10 M-triangle scene). The shading kernels therefore
// leave an 18-bit coherence key per surviving ray — where it starts (15 bits: the leaf slot of the shape it leaves, BVH order
// being a spatial order; or the Morton cell of its origin) and the octant it heads into (3 bits) — and a counting sort over
// the keys (histogram -> exclusive scan -> scatter, queue length read on the device) produces `perm`: sorted position ->
// queue slot. Nothing is moved: the next bounce's closest-hit kernel (and optionally the material sort, so that shading and
// shadow rays run in the same order) fetches its rays through `perm`. The order inside a bin is arbitrary; no result depends
// on the order rays are processed in (every path owns its state, DESIGN.md §4 "Determinism").
// Part of the single translation unit render.cu.
#pragma once
#include "wf_common.cuh"

namespace {

__device__ __forceinline__ uint32_t spread3_5(uint32_t v) {  // 5 bits -> every third bit
    v &= 0x1fu;
    v = (v | (v << 8)) & 0x100fu;
    v = (v | (v << 4)) & 0x10c3u;
    v = (v | (v << 2)) & 0x1249u;
    return v;
}
// The key of a ray leaving shape slot `hit_slot` from origin `o` in direction `d`.
__device__ __forceinline__ uint32_t ray_sort_key(const DevScene& sc, const RenderCfg& cfg, uint32_t hit_slot, float ox, float oy, float oz,
                                                 float dx, float dy, float dz) {
    const uint32_t octant = (dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u) | (dz < 0.0f ? 4u : 0u);
    uint32_t cell;
    if (cfg.sort_key_mode == 1u) {
        cell = hit_slot >> cfg.sort_slot_shift;
    } else {
        const float cx = (ox - sc.root_min[0]) * cfg.sort_cell_scale[0], cy = (oy - sc.root_min[1]) * cfg.sort_cell_scale[1];
        const float cz = (oz - sc.root_min[2]) * cfg.sort_cell_scale[2];
        const uint32_t ix = (uint32_t)fminf(fmaxf(cx, 0.0f), 31.0f), iy = (uint32_t)fminf(fmaxf(cy, 0.0f), 31.0f);
        const uint32_t iz = (uint32_t)fminf(fmaxf(cz, 0.0f), 31.0f);
        cell = spread3_5(ix) | (spread3_5(iy) << 1) | (spread3_5(iz) << 2);
    }
    return ((cell << 3) | octant) & (kSortBins - 1u);
}

// Histogram of the next queue's keys. One global atomic per distinct key of a warp.
__global__ void k_sort_hist(Wave w, const IterCounters* nxt) {
    const uint32_t n = nxt->n_active;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < n; base += n_warps * 32u) {  // warp-uniform
        const uint32_t i = base + lane;
        const uint32_t key = i < n ? w.sort_key[i] : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        YK_ASSERT(i >= n || key < kSortBins);
        if (i < n && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&w.sort_bins[key], (uint32_t)__popc(peers));
    }
}
// After the exclusive scan: bins[] are write cursors.
__global__ void k_sort_scatter(Wave w, const IterCounters* nxt) {
    const uint32_t n = nxt->n_active;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t base = warp * 32u; base < n; base += n_warps * 32u) {
        const uint32_t i = base + lane;
        const uint32_t key = i < n ? w.sort_key[i] : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        const int leader = __ffs(peers) - 1;
        uint32_t first = 0;
        if (i < n && lane == (uint32_t)leader) first = atomicAdd(&w.sort_bins[key], (uint32_t)__popc(peers));
        first = __shfl_sync(0xffffffffu, first, leader);
        YK_ASSERT(i >= n || first + __popc(peers & ((1u << lane) - 1u)) < n);
        if (i < n) w.perm[first + __popc(peers & ((1u << lane) - 1u))] = i;
    }
}

}  // namespace
