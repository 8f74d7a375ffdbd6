// Film kernels: per-pixel sample sums, the mean / overwrite of Film::update_tile, the accumulating film.
// Part of the single translation unit render.cu (compiled --fmad=false: every float op is the reference's un-fused IEEE op).
#pragma once
#include "wf_common.cuh"

namespace {

// ---- film --------------------------------------------------------------------------------------------
// `color += li` over ascending sample index (integrators/mod.rs:172), carried across batches in `accum`.
__global__ void k_film_accumulate(Wave w, Batch bt, float* accum, uint32_t res_x) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= bt.n_jobs) return;
    const Job job = bt.jobs[j];
    float* a = accum + ((size_t)job.y * res_x + job.x) * 3;
    float r = a[0], g = a[1], b = a[2];
    for (uint32_t s = 0; s < bt.n_samples; ++s) {
        const float4 L = ld_once(&w.L[(size_t)s * bt.n_jobs + j]);
        r = r + L.x; g = g + L.y; b = b + L.z;
    }
    a[0] = r; a[1] = g; a[2] = b;
}
// `color /= sample_count` + Film::update_tile overwrite (integrators/mod.rs:175-182, film.rs:274-279)
__global__ void k_film_store(const Job* jobs, uint32_t n_jobs, const float* accum, float* film, uint32_t res_x, float spp) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const Job job = jobs[j];
    const size_t p = ((size_t)job.y * res_x + job.x) * 3;
    film[p] = accum[p] / spp;
    film[p + 1] = accum[p + 1] / spp;
    film[p + 2] = accum[p + 2] / spp;
}
// Accumulating film: `*fc += c` per tile sample (film.rs:260-272); tiles of different samples may overlap.
__global__ void k_film_add(Wave w, Batch bt, float* film, uint32_t res_x) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bt.n_paths) return;
    const Job job = bt.jobs[bt.div_jobs.mod(i)];
    const float4 L = ld_once(&w.L[i]);
    float* f = film + ((size_t)job.y * res_x + job.x) * 3;
    atomicAdd(f, L.x); atomicAdd(f + 1, L.y); atomicAdd(f + 2, L.z);
}
__global__ void k_zero_jobs(const Job* jobs, uint32_t n_jobs, float* accum, uint32_t res_x) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const size_t p = ((size_t)jobs[j].y * res_x + jobs[j].x) * 3;
    accum[p] = 0.0f; accum[p + 1] = 0.0f; accum[p + 2] = 0.0f;
}
__global__ void k_fill_i32(int32_t* p, size_t n, int32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace
