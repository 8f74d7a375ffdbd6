// B200 (sm_100a) wavefront renderer behind yk_render: replaces Integrator::render + Film::update_tile
// (yuki/src/integrators/mod.rs:120-185, film.rs:210-282) for a list of film tiles.
//
// Pipeline per batch of (pixel, sample) paths, all state in HBM as SoA:
//   raygen -> [ trace_closest -> classify(compact by material) -> shade_{matte,glass,metal,glossy}
//               -> trace_any(shadow) -> resolve(+compact survivors) ]* -> film_accumulate
// Nothing here is a dense contraction, so no tensor cores: the hot kernel (trace_closest) is a
// dependent-load graph walk bounded by L2/HBM latency and bandwidth.
//
// Compiled with --fmad=false: every float op is the reference's un-fused IEEE op.
#include <cuda_runtime.h>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <memory>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "yk_device.cuh"
#include "yuki_gpu.h"

int yk_set_error(int code, const std::string& msg);  // host_scene.cpp

using namespace ykd;

namespace {

constexpr int kMaxLights = 16;
constexpr int kStackDepth = 64;  // bvh.rs:172
constexpr uint32_t kMiss = 0xffffffffu;
constexpr int kTraceThreads = 128;
#ifndef YK_SHADE_THREADS
#define YK_SHADE_THREADS 128
#endif
constexpr int kShadeThreads = YK_SHADE_THREADS;
#ifndef YK_TRACE_MIN_BLOCKS
#define YK_TRACE_MIN_BLOCKS 8
#endif
#ifndef YK_SHADOW_MIN_BLOCKS
#define YK_SHADOW_MIN_BLOCKS 8
#endif
#ifndef YK_SHADE_MIN_BLOCKS
#define YK_SHADE_MIN_BLOCKS (1024 / YK_SHADE_THREADS)
#endif

#define CUDA_TRY(expr)                                                                                          \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return yk_set_error(YK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));              \
    } while (0)

// ---- device-resident scene ------------------------------------------------------------------------
struct DevTexture {
    uint32_t kind, width, height, _pad;
    float value[3];
    float _pad2;
    const float* texels;
};
struct DevMaterial {
    uint32_t kind;
    int32_t tex[3];
    float eta;
    uint32_t remap;
    float const_alpha;  // >= 0: roughness texture is constant, alpha fully evaluated on the host
    uint32_t _pad;
};
struct DevScene {
    // One 64-byte record per *interior* node holding the boxes of its two children:
    //   (p_min child0, ref0) (p_max child0, -) (p_min child1, ref1) (p_max child1, -)
    // child0 = the node after its parent in the reference's pre-order array, child1 = second_child_index (bvh.rs:396-419).
    // A visit loads both boxes with one request; leaves have no record of their own (their shape range is in the ref).
    const float4* nodes2;
    const uint2* leaf_table;  // null: leaf refs are packed (count - 1) << 27 | first; else ref = index of (first, count)
    const float4* tris;    // 3 x float4 per triangle: (x0 x1 x2, area_light) (y0 y1 y2, material | flags<<24) (z0 z1 z2, orig_id)
    const float* normals;  // 9 per triangle or null
    const float* uvs;      // 6 per triangle or null
    const DevTexture* textures;
    const DevMaterial* materials;
    const yk_light* lights;
    const yk_sphere* spheres;  // sphere slots of `tris`: vertex lanes are NaN (so the triangle test "accepts" them and the
                               // rare hit path takes over), row 0's w = -2 - sphere index
    uint32_t n_lights, n_tris, n_nodes;
    float background[3];
    float root_min[3], root_max[3];  // the root's own box (tested once per ray)
    uint32_t root_ref;
};
// Child refs: interior = kRefInterior | split_axis << 29 | record index (the axis picks the near child before the record is
// loaded); leaf = bit 31 clear, see leaf_table. kNoNode (all ones) is the stack sentinel / "no node".
constexpr uint32_t kRefInterior = 0x80000000u;
constexpr uint32_t kRefIndexMask = 0x1fffffffu;
constexpr uint32_t kLeafFirstBits = 27;

// ---- per-iteration device counters ----------------------------------------------------------------
// Queue lengths never leave the device: every kernel of a bounce reads its element count from `cur` and appends to
// `nxt`, so a batch is one asynchronous launch sequence (no host round trip per bounce).
struct IterCounters {
    uint32_t n_active;      // rays of this bounce (length of the active queue)
    uint32_t mat[4];        // material queue lengths
    uint32_t work_closest;  // dynamic ray fetch cursors
    uint32_t work_shadow;
    uint32_t _pad;
};
struct Totals {
    unsigned long long closest_nodes, closest_tris, any_nodes, any_tris, hit_hash, shadow_rays, closest_rays;
};

struct Job {  // + the pixel's PCG stream, (SipHash13(x, y) << 1) | 1 (uniform.rs:77-81): one hash per pixel, not per sample
    uint16_t x, y;
    uint32_t sample_begin;
    unsigned long long rng_inc;
};
// Path i of a batch is sample (sample_begin + sample_off + i / n_jobs) of pixel jobs[i % n_jobs]: a warp holds
// 32 neighbouring pixels of one tile row at the same sample index.
struct Batch {
    const Job* jobs;
    uint32_t n_jobs, sample_off, n_samples, n_paths;
    FastDiv div_jobs;  // by n_jobs
};

// ---- wavefront state (SoA, capacity `cap` paths) --------------------------------------------------
// Per bounce a path touches: ray (32 B) + hit (8 B) in the traversal; ray, hit, rng state (8 B), beta (16 B) in
// shading, which writes the next ray / beta / rng state, the pending radiance terms (32 B) and 40 B per light that
// needs a shadow ray; the shadow kernel reads those back and does the one read-modify-write of L.
struct Wave {
    uint32_t cap, n_lights, stack_entries;
    // Per-bounce path state, streamed: bounce b reads st[b & 1] at the ray's queue slot and the shading kernels write the
    // survivors' state to st[(b + 1) & 1] at their position in the next queue. Every kernel therefore reads and writes
    // dense, (near-)coalesced arrays; nothing is gathered through a path index except L and the Whitted stack.
    struct Stream {
        float4* ray_o;   // o.xyz, t_max
        float4* ray_d;   // d.xyz, -
        float4* beta;    // throughput (path) / node weight (whitted); w = flags | sampler dimension << kDimShift
        unsigned long long* rng;
    } st[2];
    uint2* hit;         // per queue slot: t bits, shape slot (kMiss = none)
    uint2* bvh_counts;  // BVHIntersections: tests, hits
    float4* L;       // accumulated radiance
    // The five arrays below are the shading kernels' hand-over to the shadow kernel. They are indexed by the *shading
    // position* g (position in the concatenation of this bounce's four material queues), not by path: the shading
    // kernels write them fully coalesced and the shadow kernel streams them with no dependent gather.
    uint32_t* sh_path;   // path of shading position g
    float4* pend_beta;   // weight to apply to this bounce's radiance; w = clamp flag
    float4* pend_extra;  // emitted term of this bounce; w = bit mask of the lights whose shadow ray must be traced
    float4* lt_o;        // cap * n_lights: shadow ray o.xyz | contribution.r   (contribution = f * li * cos / pdf)
    float4* lt_d;        //                 shadow ray d.xyz | contribution.g
    float2* lt_c;        //                 contribution.b   | area light id of the sampled light (int bits, -1 = none)
    float4* stack;       // whitted: stack_entries * cap * 3 float4
    uint32_t* stack_top; // whitted
    uint32_t* q_active[2];
    uint32_t* q_mat;     // 4 * cap: paths per material kind
    uint32_t* q_mat_tri; // 4 * cap: the hit shape slot of each entry
    uint32_t* q_mat_slot; // 4 * cap: the entry's slot in this bounce's active queue (index of st[] / hit[])
    Totals* totals;
};
// beta.w flag word
constexpr uint32_t kFlagSpecular = 0x100u;   // path: specular_bounce / whitted: is_specular
constexpr uint32_t kFlagAlive = 0x200u;
constexpr uint32_t kDepthMask = 0xffu;       // path: bounces / whitted: depth
constexpr uint32_t kDimShift = 10;           // sampler dimension (stratified.rs:40) in the upper 22 bits
constexpr uint32_t kFlagMask = (1u << kDimShift) - 1u;

struct RenderCfg {
    SamplerCfg sampler;
    uint32_t integrator, max_depth, has_clamp;
    float clamp;
    float c2w[16], r2c[16];
    uint32_t res_x, res_y;
    uint32_t aux_sample;
    int32_t* hit_ids;  // device, or null
};

// ---- helpers --------------------------------------------------------------------------------------
__device__ __forceinline__ V3 f4v(float4 a) { return {a.x, a.y, a.z}; }

// Warp-aggregated append: one atomic per warp per queue.
__device__ __forceinline__ void queue_push(bool pred, uint32_t value, uint32_t* queue, uint32_t* counter) {
    const unsigned active = __activemask();
    const unsigned votes = __ballot_sync(active, pred);
    if (votes == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(votes) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(counter, (uint32_t)__popc(votes));
    base = __shfl_sync(active, base, leader);
    if (pred) queue[base + __popc(votes & ((1u << lane) - 1u))] = value;
}
__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
// Block-aggregated append to one of NQ queues: one global atomic per queue per block (same-address atomics were the
// bottleneck of classify / resolve with one atomic per warp, profiles/r01). `key` in [0, NQ) selects the queue, any
// other value appends nothing. Must be reached by every thread of the block (blockDim.x <= 256).
// Returns the slot the value was written to (undefined when nothing was appended).
// `K` items per thread share the block's atomics: K * blockDim.x items per global atomic and queue.
template <int NQ, int K>
__device__ __forceinline__ void block_scatter_multi(const int (&key)[K], const uint32_t (&value)[K], uint32_t* const (&queues)[NQ],
                                                    uint32_t* const (&counters)[NQ], uint32_t (&pos)[K]) {
    __shared__ uint32_t s_cnt[8][NQ];
    __shared__ uint32_t s_base[NQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = (blockDim.x + 31) >> 5;
    const unsigned lt = (1u << lane) - 1u;
    uint32_t run[NQ];  // warp-uniform: entries of this warp per queue so far
    uint32_t my_rank[K];
#pragma unroll
    for (int q = 0; q < NQ; ++q) run[q] = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        my_rank[k] = 0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const unsigned votes = __ballot_sync(0xffffffffu, key[k] == q);
            if (key[k] == q) my_rank[k] = run[q] + __popc(votes & lt);
            run[q] += __popc(votes);
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) s_cnt[warp][q] = run[q];
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        uint32_t total = 0;
        for (int wi = 0; wi < n_warps; ++wi) {
            const uint32_t c = s_cnt[wi][threadIdx.x];
            s_cnt[wi][threadIdx.x] = total;  // exclusive prefix over the block's warps
            total += c;
        }
        s_base[threadIdx.x] = total ? atomicAdd(counters[threadIdx.x], total) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        pos[k] = 0;
        if (key[k] >= 0 && key[k] < NQ) {
            pos[k] = s_base[key[k]] + s_cnt[warp][key[k]] + my_rank[k];
            queues[key[k]][pos[k]] = value[k];
        }
    }
    __syncthreads();  // the shared arrays are reused by the next call
}
template <int NQ>
__device__ __forceinline__ uint32_t block_scatter(int key, uint32_t value, uint32_t* const (&queues)[NQ], uint32_t* const (&counters)[NQ]) {
    const int keys[1] = {key};
    const uint32_t values[1] = {value};
    uint32_t pos[1];
    block_scatter_multi<NQ, 1>(keys, values, queues, counters, pos);
    return pos[0];
}
__device__ __forceinline__ unsigned long long mix_hit(uint32_t x, uint32_t y, uint32_t sample, uint32_t id) {
    unsigned long long h = ((unsigned long long)x << 48) ^ ((unsigned long long)y << 32) ^ ((unsigned long long)sample << 8) ^
                           (unsigned long long)id * 0x9E3779B97F4A7C15ULL;
    h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 29;
    return h;
}

// ---- pixel jobs of one pixel group, expanded on the device from the tile list -----------------------------------
// Job j of a render is pixel (j - off[t]) of tile t in row-major order (Bounds2 iteration, math/bounds.rs:102-126), tiles
// in list order; `off` is the prefix sum of the tile areas. Also hash_values!(pixel.x, pixel.y), the pixel's sampler
// stream (uniform.rs:77, stratified.rs:95). The host uploads 16 bytes per tile instead of 8 per pixel.
__global__ void k_jobs_expand(const yk_tile* tiles, const unsigned long long* off, uint32_t t_lo, uint32_t t_hi, unsigned long long j0,
                              uint32_t n, uint32_t accumulate, Job* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long j = j0 + i;
    uint32_t lo = t_lo, hi = t_hi;  // off[lo] <= j < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (off[mid] <= j) lo = mid;
        else hi = mid;
    }
    const yk_tile tl = tiles[lo];
    const uint32_t local = (uint32_t)(j - off[lo]), w = (uint32_t)tl.x1 - tl.x0;
    const uint32_t row = local / w;
    Job o;
    o.x = (uint16_t)(tl.x0 + (local - row * w));
    o.y = (uint16_t)(tl.y0 + row);
    o.sample_begin = accumulate ? tl.sample : 0u;
    o.rng_inc = (hash_pixel(o.x, o.y) << 1) | 1ULL;
    out[i] = o;
}

// hash_values!(pixel.x, pixel.y, dimension, seed) for every (dimension, pixel) of a pixel group (SamplerCfg::hash_table)
__global__ void k_dim_hashes(const Job* jobs, uint32_t n_jobs, uint32_t n_dims, unsigned long long seed, uint32_t* out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_jobs) return;
    const Job job = jobs[j];
    for (uint32_t dim = blockIdx.y; dim < n_dims; dim += gridDim.y)
        out[(size_t)dim * n_jobs + j] = (uint32_t)hash_pixel_dim_seed(job.x, job.y, dim, seed);
}

// ---- raygen: Integrator::render loop head (integrators/mod.rs:145-169) -----------------------------
// `rng.advance(sample_index * 65536)` (uniform.rs:81-83, stratified.rs:99-101) is an LCG jump: state' = M * state + inc * P
// with M, P functions of the distance only (the jump's additive term is linear in the stream increment). A tiny kernel
// (k_sample_jumps) tabulates (M, P) for the batch's consecutive sample indices, so seeking costs two multiplies instead of the
// O(log n) loop — which was most of this kernel's instructions.
constexpr uint32_t kMaxBatchSamples = 256;  // consecutive sample indices of a pixel per batch (size of the jump table)
struct SampleJump {
    unsigned long long mult, plus;
};
// jumps[k] = the (M, P) of Lcg64Xsh32::advance((first_sample + k) * 65536) for increment 1
__global__ void k_sample_jumps(uint32_t first_sample, uint32_t n, SampleJump* jumps) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    unsigned long long delta = (unsigned long long)(first_sample + k) * 65536ull, am = 1, ap = 0, cm = kPcgMult, cp = 1;
    while (delta) {
        if (delta & 1ull) { am *= cm; ap = ap * cm + cp; }
        cp = (cm + 1ull) * cp;
        cm *= cm;
        delta >>= 1;
    }
    jumps[k] = SampleJump{am, ap};
}
__global__ void k_raygen(Wave w, RenderCfg cfg, Batch bt, uint32_t first_sample, uint32_t n_jumps, const SampleJump* jumps, IterCounters* first) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bt.n_paths) return;
    if (i == 0) first->n_active = bt.n_paths;
    const uint32_t si = bt.div_jobs.div(i), ji = i - si * bt.n_jobs;
    const Job job = bt.jobs[ji];
    const uint32_t sample = job.sample_begin + bt.sample_off + si;
    SamplerState s;
    const uint32_t slot = sample - first_sample;
    if (slot < n_jumps) {  // (always, for the batches yk_render builds)
        s.px = job.x; s.py = job.y; s.index = sample; s.dim = 0; s.job = ji;
        s.rng.inc = job.rng_inc;
        const unsigned long long seeded = (cfg.sampler.seed + job.rng_inc) * kPcgMult + job.rng_inc;  // Lcg64Xsh32::new
        const SampleJump j = jumps[slot];
        s.rng.state = j.mult * seeded + job.rng_inc * j.plus;
    } else {
        s.start(cfg.sampler, job.x, job.y, sample, job.rng_inc, ji);
    }
    const V2 j = s.get_2d(cfg.sampler);
    // Camera::ray, camera.rs:105-114
    const V3 p_cam = xf_point(cfg.r2c, mk((float)job.x + j.x, (float)job.y + j.y, 0.0f));
    const V3 d_cam = unit(p_cam);
    const V3 o = xf_point(cfg.c2w, mk(0.0f, 0.0f, 0.0f));
    const V3 d = xf_vec(cfg.c2w, d_cam);
    w.st[0].ray_o[i] = make_float4(o.x, o.y, o.z, __int_as_float(0x7f800000));
    w.st[0].ray_d[i] = make_float4(d.x, d.y, d.z, 0.0f);
    w.st[0].rng[i] = s.rng.state;
    w.st[0].beta[i] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(kFlagAlive | (s.dim << kDimShift)));
    w.L[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (w.stack_top) w.stack_top[i] = 0;
}

// ---- BVH traversal (bvh.rs:160-302, math/bounds.rs:176-215, shapes/triangle.rs:49-139) --------------
// Persistent warps, one ray per lane. The kernels are issue-bound on small scenes and latency-bound on large ones
// (profiles/r01), so the design goal is: few instructions per step, and as many lanes as possible per instruction.
//  * Two phases per warp: box steps (N) and triangle steps (T). A lane that reaches a leaf parks until the warp
//    serves leaves; the warp keeps stepping boxes while at least kNodePhaseMin lanes want to, then drains every
//    parked leaf. A lane never walks past its own leaf, so each ray performs exactly the reference's sequence of
//    box and triangle tests (the counters are bit-exact). Policy chosen with scripts/sim_warp.py.
//  * Finished lanes are refilled from the ray queue once fewer than kRefillBelow lanes are live; a warp reserves
//    kChunk rays from the global cursor at a time.
//  * Both steps are branch-free apart from the rare f64 edge-function fallback. The traversal stack lives in shared
//    memory as s_stack[depth][thread] (conflict-free for any mix of depths) above a kNoNode sentinel, so a pop needs no
//    emptiness test; entries beyond kShortStack spill to local memory, up to the reference's 64.
//  * Triangles are stored transposed (x0 x1 x2 | y0 y1 y2 | z0 z1 z2), so the watertight test's axis permutation is
//    three index offsets instead of 18 selects.
constexpr uint32_t kNoNode = 0xffffffffu;
constexpr uint32_t kChunk = 64;
#ifndef YK_REFILL_BELOW
#define YK_REFILL_BELOW 22
#endif
#ifndef YK_NODE_PHASE_MIN
#define YK_NODE_PHASE_MIN 14
#endif
constexpr int kRefillBelow = YK_REFILL_BELOW;
constexpr int kNodePhaseMin = YK_NODE_PHASE_MIN;
#ifndef YK_SHORT_STACK
#define YK_SHORT_STACK 16
#endif
constexpr int kShortStack = YK_SHORT_STACK;
// s_stack[depth][thread] = (child ref, key = the child's clamped slab entry distance): one 64-bit access per push / pop,
// conflict-free for any mix of depths (a half-warp's 16 entries cover the 32 banks)
constexpr uint32_t kStackStride = kTraceThreads * 8;  // bytes between two levels of one lane's stack
constexpr int kDeepStack = kStackDepth + 1 - kShortStack;
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void lds_entry(uint32_t addr, uint32_t* ref, float* key) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(*ref), "=f"(*key) : "r"(addr));
}
__device__ __forceinline__ void sts_entry(uint32_t addr, uint32_t ref, float key) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ref), "f"(key) : "memory");
}

// Sphere slots are rare: the test lives behind a real call so that it costs the traversal loops no registers.
__device__ __noinline__ bool sphere_slot_test(const yk_sphere* spheres, int tag, float ox, float oy, float oz, float4 rd, float t_max,
                                              float* t_out) {
    V3 o_s, d_s;
    return sphere_test(spheres[-2 - tag], mk(ox, oy, oz), f4v(rd), t_max, t_out, &o_s, &d_s);
}

// Traversal state of one ray. The walk performs exactly the reference's sequence of box and shape tests, but is
// organised around the 64-byte two-child records:
//  * entering an interior node loads both children's boxes at once and slab-tests both. The near child (by the ray's
//    sign on the split axis, bvh.rs:186-194) is tested against the current t_max, as the reference does next. The far
//    child's test happens later in the reference, with whatever t_max is current *then* — but the slab arithmetic does
//    not depend on t_max except through the final `min(.., t_max)`, so the far child's clamped entry distance is kept
//    as the stack entry's `key` and the deferred test is `key <= t_max` at pop time: no memory access for a popped
//    node that misses, and only nodes whose box test passes are ever loaded (half the dependent loads of a one-node-
//    per-visit walk). A far child that can never pass (entry beyond its own exit) gets a NaN key.
//  * counters: the closest-hit walk always drains its stack, so both tests of a record are counted when it is loaded
//    and never-passing far children are not pushed; the any-hit walk ends early, so it counts a far child's test when
//    it is popped (or tested on the spot) and pushes NaN-key entries too.
struct TraceLane {
    float ox, oy, oz, ix, iy, iz, t_max;
    float okx, oky, okz, sx, sy, sz;  // watertight test: permuted origin, shear
    uint32_t kx, ky, kz, neg_mask;
    uint32_t cur;  // interior ref to enter next, or kNoNode
    uint32_t leaf_pos, leaf_end;
    uint32_t sp;  // shared-memory byte address of the lane's next free stack entry (level 0 holds the sentinel)
    uint32_t n_tests, n_tris;  // running totals of the lane (all of its rays): box tests, shape tests
    uint32_t n_hits;           // passed box tests of the current ray (COUNTS only)

    __device__ __forceinline__ void idle(uint32_t sbase) {
        cur = kNoNode; sp = sbase + kStackStride; leaf_pos = leaf_end = 0; n_tests = n_hits = n_tris = 0;
        ox = oy = oz = ix = iy = iz = t_max = okx = oky = okz = sx = sy = sz = 0.0f;
        kx = ky = kz = neg_mask = 0;
    }
    // Slab distances of one box (math/bounds.rs:176-215): lo = max(max_comp(min(t0, t1)), 0), hi = min_comp(max(t0, t1));
    // the reference's test is lo <= min(hi, t_max). NaN-ignoring min/max exactly like f32::min/max.
    __device__ __forceinline__ void slab(float ax, float ay, float az, float bx, float by, float bz, float* lo, float* hi) const {
        const float t0x = (ax - ox) * ix, t0y = (ay - oy) * iy, t0z = (az - oz) * iz;
        const float t1x = (bx - ox) * ix, t1y = (by - oy) * iy, t1z = (bz - oz) * iz;
        *lo = fmaxf(fmaxf(fminf(t0x, t1x), fmaxf(fminf(t0y, t1y), fminf(t0z, t1z))), 0.0f);
        *hi = fminf(fmaxf(t0x, t1x), fminf(fmaxf(t0y, t1y), fmaxf(t0z, t1z)));
    }
    // Makes `ref` the lane's next piece of work: an interior node to enter, or a leaf to park. Select-only (no branch
    // but the warp-uniform leaf-table one).
    template <bool GENERIC>
    __device__ __forceinline__ void enter(const DevScene& sc, uint32_t ref) {
        const bool interior = (int32_t)ref < 0;  // (the sentinel kNoNode counts as interior and ends the ray)
        cur = interior ? ref : kNoNode;
        uint32_t first, count;
        if (GENERIC && sc.leaf_table) {
            const uint2 l = interior ? make_uint2(0u, 0u) : __ldg(&sc.leaf_table[ref]);
            first = l.x; count = l.y;
        } else {
            first = ref & ((1u << kLeafFirstBits) - 1u);
            count = (ref >> kLeafFirstBits) + 1u;
        }
        leaf_pos = interior ? leaf_pos : first;
        leaf_end = interior ? leaf_end : first + count;
    }
    template <bool COUNTS, bool GENERIC>
    __device__ __forceinline__ void start(const DevScene& sc, uint32_t sbase, float o_x, float o_y, float o_z, float d_x, float d_y, float d_z,
                                          float tmax) {
        ox = o_x; oy = o_y; oz = o_z;
        t_max = tmax;
        ix = 1.0f / d_x; iy = 1.0f / d_y; iz = 1.0f / d_z;  // bvh.rs:164
        neg_mask = (ix < 0.0f ? 1u : 0u) | (iy < 0.0f ? 2u : 0u) | (iz < 0.0f ? 4u : 0u);
        // triangle.rs:58-80: permutation and shear depend on the ray only
        const float ax = fabsf(d_x), ay = fabsf(d_y), az = fabsf(d_z);
        kz = ax > ay ? (ax > az ? 0u : 2u) : (ay > az ? 1u : 2u);  // Vec3::max_dimension, math/vector.rs:188-202
        kx = kz < 2u ? kz + 1u : 0u;
        ky = kx < 2u ? kx + 1u : 0u;
        const float dkx = kx == 0 ? d_x : (kx == 1 ? d_y : d_z), dky = ky == 0 ? d_x : (ky == 1 ? d_y : d_z);
        const float dkz = kz == 0 ? d_x : (kz == 1 ? d_y : d_z);
        sx = -dkx / dkz; sy = -dky / dkz;
        sz = kz == 0 ? ix : (kz == 1 ? iy : iz);  // 1.0 / d[kz]: the same IEEE division as above
        okx = kx == 0 ? ox : (kx == 1 ? oy : oz);
        oky = ky == 0 ? ox : (ky == 1 ? oy : oz);
        okz = kz == 0 ? ox : (kz == 1 ? oy : oz);
        sp = sbase + kStackStride; n_tests += 1; n_hits = 0;
        leaf_pos = leaf_end = 0;
        cur = kNoNode;
        // the root's own box (bvh.rs:176-179 on node 0)
        float lo, hi;
        slab(sc.root_min[0], sc.root_min[1], sc.root_min[2], sc.root_max[0], sc.root_max[1], sc.root_max[2], &lo, &hi);
        if (lo <= fminf(hi, t_max)) {
            if (COUNTS) n_hits = 1;
            enter<GENERIC>(sc, sc.root_ref);
        }
    }
    __device__ __forceinline__ bool wants_box() const { return cur != kNoNode; }
    __device__ __forceinline__ bool wants_tri() const { return leaf_pos < leaf_end; }
    __device__ __forceinline__ void push(uint32_t sbase, uint32_t* deep_ref, float* deep_key, uint32_t ref, float key) {
        if (sp < sbase + (uint32_t)kShortStack * kStackStride) {
            sts_entry(sp, ref, key);
        } else {  // cold: the stack continues in local memory, up to the reference's 64 entries
            const uint32_t depth = (sp - sbase) / kStackStride - kShortStack;
            deep_ref[depth] = ref;
            deep_key[depth] = key;
        }
        sp += kStackStride;
    }
    // Pops until an entry passes its deferred box test (the sentinel's key 0 always does). Returns its ref.
    template <bool COUNTS, bool ANYHIT>
    __device__ __forceinline__ uint32_t pop_passing(uint32_t sbase, const uint32_t* deep_ref, const float* deep_key) {
        uint32_t ref;
        float key;
        if (sp > sbase + (uint32_t)kShortStack * kStackStride) {  // cold: the top of the stack is in local memory
            do {
                sp -= kStackStride;
                if (sp < sbase + (uint32_t)kShortStack * kStackStride) {
                    lds_entry(sp, &ref, &key);
                } else {
                    const uint32_t depth = (sp - sbase) / kStackStride - kShortStack;
                    ref = deep_ref[depth];
                    key = deep_key[depth];
                }
                if (ANYHIT) n_tests += ref != kNoNode ? 1u : 0u;
            } while (!(key <= t_max));
        } else {
            do {
                sp -= kStackStride;
                lds_entry(sp, &ref, &key);
                if (ANYHIT) n_tests += ref != kNoNode ? 1u : 0u;
            } while (!(key <= t_max));
        }
        if (COUNTS) n_hits += ref != kNoNode ? 1u : 0u;
        return ref;
    }
    // Enters the interior node `cur`: the box tests of its two children (bvh.rs:176-199, math/bounds.rs:176-215).
    template <bool COUNTS, bool ANYHIT, bool GENERIC>
    __device__ __forceinline__ void box_step(const DevScene& sc, uint32_t sbase, uint32_t* deep_ref, float* deep_key) {
        // near child first: the second child when the ray is negative on the split axis (bvh.rs:186-194)
        const uint32_t neg = (neg_mask >> ((cur >> 29) & 3u)) & 1u;
        const float4* rec = sc.nodes2 + 4 * (size_t)(cur & kRefIndexMask);
        const float4* near = rec + 2 * neg;
        const float4* far = rec + 2 * (neg ^ 1u);
        const float4 n0 = __ldg(near), n1 = __ldg(near + 1);
        const float4 f0 = __ldg(far), f1 = __ldg(far + 1);
        float lo_n, hi_n, lo_f, hi_f;
        slab(n0.x, n0.y, n0.z, n1.x, n1.y, n1.z, &lo_n, &hi_n);
        slab(f0.x, f0.y, f0.z, f1.x, f1.y, f1.z, &lo_f, &hi_f);
        const uint32_t ref_n = __float_as_uint(n0.w), ref_f = __float_as_uint(f0.w);
        const bool hit_n = lo_n <= fminf(hi_n, t_max);
        const bool ok_f = !(lo_f > hi_f);  // can the far child pass at all? (a NaN hi is ignored by the reference's min)
        const float key_f = ok_f ? lo_f : __int_as_float(0x7fc00000);
        // near missed: nothing happens before the far child's test, t_max is what the pop would see
        const bool hit_f = !hit_n && key_f <= t_max;
        n_tests += (ANYHIT && hit_n) ? 1u : 2u;
        if (COUNTS) n_hits += (hit_n || hit_f) ? 1u : 0u;
        const bool do_push = hit_n && (ANYHIT || ok_f);
        uint32_t take = hit_n ? ref_n : ref_f;
        if (sp >= sbase + (uint32_t)kShortStack * kStackStride) {  // cold: the stack continues in local memory
            if (do_push) push(sbase, deep_ref, deep_key, ref_f, key_f);
            if (!(hit_n || hit_f)) take = pop_passing<COUNTS, ANYHIT>(sbase, deep_ref, deep_key);
        } else {
            if (do_push) sts_entry(sp, ref_f, key_f);
            sp += do_push ? kStackStride : 0u;
            if (!(hit_n || hit_f)) {
                float key;
                do {
                    sp -= kStackStride;
                    lds_entry(sp, &take, &key);
                    if (ANYHIT) n_tests += take != kNoNode ? 1u : 0u;
                } while (!(key <= t_max));
                if (COUNTS) n_hits += take != kNoNode ? 1u : 0u;
            }
        }
        enter<GENERIC>(sc, take);
    }
    // One triangle test of the parked leaf (shapes/triangle.rs:62-130 on the permuted, origin-relative vertices).
    // Returns true on a hit with t in (0, t_max]; the caller decides what a hit means and then calls leaf_done().
    __device__ __forceinline__ bool tri_step(const DevScene& sc, uint32_t* tri, float* t_scaled_out, float* det_out, int* area_light) {
        const uint32_t s = leaf_pos++;
        const float4 A = __ldg(&sc.tris[3 * s + kx]);
        const float4 B = __ldg(&sc.tris[3 * s + ky]);
        const float4 C = __ldg(&sc.tris[3 * s + kz]);
        n_tris += 1;
        float ax = A.x - okx, bx = A.y - okx, cx = A.z - okx;
        float ay = B.x - oky, by = B.y - oky, cy = B.z - oky;
        const float az = C.x - okz, bz = C.y - okz, cz = C.z - okz;
        ax += sx * az; ay += sy * az;
        bx += sx * bz; by += sy * bz;
        cx += sx * cz; cy += sy * cz;
        float e0 = bx * cy - by * cx;
        float e1 = cx * ay - cy * ax;
        float e2 = ax * by - ay * bx;
        if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // f64 fallback, :98-105
            e0 = (float)((double)bx * (double)cy - (double)by * (double)cx);
            e1 = (float)((double)cx * (double)ay - (double)cy * (double)ax);
            e2 = (float)((double)ax * (double)by - (double)ay * (double)bx);
        }
        const float det = e0 + e1 + e2;
        const float t_scaled = e0 * (az * sz) + e1 * (bz * sz) + e2 * (cz * sz);
        const float lim = t_max * det;
        const bool mixed = (e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f);
        const bool out_neg = det < 0.0f && (t_scaled >= 0.0f || t_scaled < lim);
        const bool out_pos = det > 0.0f && (t_scaled <= 0.0f || t_scaled > lim);
        *tri = s;
        *t_scaled_out = t_scaled;
        *det_out = det;
        *area_light = __float_as_int(kx == 0 ? A.w : (ky == 0 ? B.w : C.w));
        return !mixed && det != 0.0f && !out_neg && !out_pos;
    }
    template <bool COUNTS, bool ANYHIT, bool GENERIC>
    __device__ __forceinline__ void leaf_done(const DevScene& sc, uint32_t sbase, const uint32_t* deep_ref, const float* deep_key) {
        if (leaf_pos == leaf_end) enter<GENERIC>(sc, pop_passing<COUNTS, ANYHIT>(sbase, deep_ref, deep_key));
    }
    // ends the ray: the next pop (leaf_done) takes the sentinel
    __device__ __forceinline__ void stop(uint32_t sbase) { cur = kNoNode; leaf_pos = leaf_end = 0; sp = sbase + kStackStride; }
};

// Runs box steps while enough lanes want one, then drains the parked leaves. `on_hit(tri, t_scaled, det, area_light)`
// is called for every accepted triangle. Returns when every lane of the warp is either finished or parked nowhere.
// Box steps per phase vote: a lane that parks or finishes in an earlier step would have idled until the phase ends
// anyway, so the extra steps only delay the phase decision and save their votes (~10 instructions each). Measured:
// 1 -> 2 -> 3 steps: -3 %, -6 % closest-hit time, 4 = 3; two triangle steps per vote: +2 % (not used).
#ifndef YK_BOX_STEPS_PER_VOTE
#define YK_BOX_STEPS_PER_VOTE 3
#endif
#define YK_TRACE_PHASES(LANE, LIVE, COUNTS, ANYHIT, GENERIC, ON_HIT)                                             \
    for (;;) {                                                                                                    \
        const bool want_n = (LANE).wants_box();                                                                   \
        const int n_n = __popc(__ballot_sync(0xffffffffu, want_n));                                               \
        if (n_n == 0) break;                                                                                      \
        if (n_n < kNodePhaseMin && __ballot_sync(0xffffffffu, (LIVE) && !want_n)) break;                          \
        if (want_n) (LANE).template box_step<COUNTS, ANYHIT, GENERIC>(sc, sbase, deep_ref, deep_key);                      \
        if (YK_BOX_STEPS_PER_VOTE > 1 && (LANE).wants_box()) (LANE).template box_step<COUNTS, ANYHIT, GENERIC>(sc, sbase, deep_ref, deep_key); \
        if (YK_BOX_STEPS_PER_VOTE > 2 && (LANE).wants_box()) (LANE).template box_step<COUNTS, ANYHIT, GENERIC>(sc, sbase, deep_ref, deep_key); \
    }                                                                                                             \
    while (__ballot_sync(0xffffffffu, (LANE).wants_tri())) {                                                      \
        if ((LANE).wants_tri()) {                                                                                 \
            uint32_t tri_; float ts_, det_; int al_;                                                              \
            if ((LANE).tri_step(sc, &tri_, &ts_, &det_, &al_)) { ON_HIT }                                         \
            (LANE).template leaf_done<COUNTS, ANYHIT, GENERIC>(sc, sbase, deep_ref, deep_key);                    \
        }                                                                                                         \
    }

// Closest hit: BoundingVolumeHierarchy::intersect (bvh.rs:160-232). One ray per queue entry.
template <bool COUNTS, bool SPHERES>
__global__ void __launch_bounds__(kTraceThreads, YK_TRACE_MIN_BLOCKS) k_trace_closest(DevScene sc, Wave w, int b, IterCounters* cur) {
    __shared__ uint2 s_stack[kShortStack][kTraceThreads];
    uint32_t deep_ref[kDeepStack];
    float deep_key[kDeepStack];
    const uint32_t n = cur->n_active;
    uint32_t* const cursor = &cur->work_closest;
    if (blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&w.totals->closest_rays, (unsigned long long)n);
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_stack[0][tid]);
    sts_entry(sbase, kNoNode, 0.0f);  // sentinel: popping it ends the ray (key 0 passes every deferred test)
    uint32_t tests_before = 0;               // COUNTS: the lane's running test count when its current ray started
    uint32_t chunk_next = 0, chunk_end = 0;  // warp-uniform
    bool exhausted = false;                  // warp-uniform: the global cursor ran past n

    TraceLane tl;
    tl.idle(sbase);
    bool live = false;
    uint32_t path = 0, hit_tri = kMiss;
    float hit_t = 0.0f;

    for (;;) {
        // ---- refill idle lanes -----------------------------------------------------------------------
        const unsigned idle = __ballot_sync(0xffffffffu, !live);
        if (idle && !exhausted) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                chunk_next = base;
                chunk_end = base + kChunk < n ? base + kChunk : n;
                if (base >= n) { exhausted = true; chunk_next = chunk_end = 0; }
            }
            if (!exhausted) {
                const uint32_t mine = chunk_next + __popc(idle & lt_mask);
                if (!live && mine < chunk_end) {
                    path = mine;  // the queue slot: rays, hits and counters of a bounce are all in queue order
                    const float4 ro = w.st[b].ray_o[path];
                    const float4 rd = w.st[b].ray_d[path];
                    if (COUNTS) tests_before = tl.n_tests;
                    tl.template start<COUNTS, SPHERES>(sc, sbase, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, ro.w);
                    hit_tri = kMiss; hit_t = 0.0f;
                    live = true;
                }
                const uint32_t taken = chunk_next + __popc(idle);
                chunk_next = taken < chunk_end ? taken : chunk_end;
            }
        }
        if (__ballot_sync(0xffffffffu, live) == 0) {
            if (exhausted) break;
            continue;
        }
        // ---- trace until too few lanes are live -----------------------------------------------------------
        for (;;) {
            YK_TRACE_PHASES(tl, live, COUNTS, false, SPHERES, {
                if (SPHERES && det_ != det_) {  // a sphere slot (NaN vertex lanes): shapes/sphere.rs:36-77
                    float t_s;
                    /* the direction is not kept in registers: re-read it on this rare path */
                    if (sphere_slot_test(sc.spheres, al_, tl.ox, tl.oy, tl.oz, w.st[b].ray_d[path], tl.t_max, &t_s)) {
                        hit_tri = tri_; hit_t = t_s; tl.t_max = t_s;
                    }
                } else {
                    const float inv_det = 1.0f / det_;  // triangle.rs:133-139
                    hit_tri = tri_; hit_t = ts_ * inv_det; tl.t_max = hit_t;  // later equal-t hit replaces (bvh.rs:204-207)
                }
            })
            if (live && !tl.wants_box()) {  // retire
                w.hit[path] = make_uint2(__float_as_uint(hit_t), hit_tri);
                if (COUNTS) w.bvh_counts[path] = make_uint2(tl.n_tests - tests_before, tl.n_hits);
                live = false;
            }
            const int busy = __popc(__ballot_sync(0xffffffffu, live));
            if (busy == 0 || (!exhausted && busy < kRefillBelow)) break;
        }
    }
    const unsigned long long sum_nodes = warp_sum((unsigned long long)tl.n_tests), sum_tris = warp_sum((unsigned long long)tl.n_tris);
    if (lane == 0 && (sum_nodes | sum_tris)) {
        atomicAdd(&w.totals->closest_nodes, sum_nodes);
        atomicAdd(&w.totals->closest_tris, sum_tris);
    }
}

// Shadow rays + radiance fold: BoundingVolumeHierarchy::any_intersect behind VisibilityTester (bvh.rs:235-302,
// visibility.rs) for every light the shading kernel queued, then the fold body `c + f*li*cos/pdf` in light order,
// `radiance += beta * Le`, the indirect clamp and `L += beta * radiance` (path.rs:113-129, whitted.rs:120-130).
// One *path* per queue entry (the four material queues, concatenated); a lane traces its path's shadow rays one after
// the other in light order, so the float sums associate exactly like the reference's fold.
template <bool SPHERES>
__global__ void __launch_bounds__(kTraceThreads, YK_SHADOW_MIN_BLOCKS) k_trace_shadow(DevScene sc, Wave w, RenderCfg cfg, IterCounters* cur) {
    uint32_t* const cursor = &cur->work_shadow;
    __shared__ uint2 s_stack[kShortStack][kTraceThreads];
    uint32_t deep_ref[kDeepStack];
    float deep_key[kDeepStack];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_stack[0][tid]);
    sts_entry(sbase, kNoNode, 0.0f);
    const uint32_t n0 = cur->mat[0], n1 = cur->mat[1], n2 = cur->mat[2], n3 = cur->mat[3];
    const uint32_t n = n0 + n1 + n2 + n3;
    uint32_t n_rays = 0;
    uint32_t chunk_next = 0, chunk_end = 0;
    bool exhausted = false;

    TraceLane tl;
    tl.idle(sbase);
    bool live = false;       // the lane owns a path whose fold is not finished
    bool need_ray = false;   // ... and must load the shadow ray of the lowest light in `mask`
    bool occluded = false;
    uint32_t path = 0, pos = 0, mask = 0;  // pos: the path's shading position (index of the hand-over arrays)
    int target_light = -1;
    RGB radiance = gray(0.0f), contribution = gray(0.0f);

    auto finish_path = [&]() {  // path.rs:121-129
        const float4 pe = w.pend_extra[pos], pb = w.pend_beta[pos];
        RGB r = radiance + rgb(pe.x, pe.y, pe.z);
        if (pb.w != 0.0f) r = rgb(fminf(r.r, cfg.clamp), fminf(r.g, cfg.clamp), fminf(r.b, cfg.clamp));
        float4 L = w.L[path];
        L.x = L.x + pb.x * r.r;
        L.y = L.y + pb.y * r.g;
        L.z = L.z + pb.z * r.b;
        w.L[path] = L;
    };

    for (;;) {
        // ---- refill: new paths for idle lanes (paths without shadow rays are folded on the spot) ------------
        for (int round = 0; round < 4; ++round) {
            const unsigned idle = __ballot_sync(0xffffffffu, !live);
            if (!idle || exhausted) break;
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
                chunk_next = base;
                chunk_end = base + kChunk < n ? base + kChunk : n;
                if (base >= n) { exhausted = true; chunk_next = chunk_end = 0; break; }
            }
            const uint32_t mine = chunk_next + __popc(idle & lt_mask);
            if (!live && mine < chunk_end) {
                pos = mine;
                path = w.sh_path[pos];
                mask = __float_as_uint(w.pend_extra[pos].w);
                radiance = gray(0.0f);
                n_rays += __popc(mask);
                if (mask) { live = true; need_ray = true; }
                else finish_path();
            }
            const uint32_t taken = chunk_next + __popc(idle);
            chunk_next = taken < chunk_end ? taken : chunk_end;
        }
        if (need_ray) {  // next light of this lane's path
            const uint32_t k = __ffs(mask) - 1;
            const size_t ref = (size_t)k * w.cap + pos;
            const float4 ro = w.lt_o[ref], rd = w.lt_d[ref];
            const float2 rc = w.lt_c[ref];
            contribution = rgb(ro.w, rd.w, rc.x);
            target_light = __float_as_int(rc.y);
            tl.template start<false, SPHERES>(sc, sbase, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, 0.9999f);  // interaction.rs:57-58
            occluded = false;
            need_ray = false;
        }
        if (__ballot_sync(0xffffffffu, live) == 0) {
            if (exhausted) break;
            continue;
        }
        for (;;) {
            YK_TRACE_PHASES(tl, live, false, true, SPHERES, {
                (void)tri_; (void)ts_;
                bool blocks = true;
                if (SPHERES && det_ != det_) {  // sphere slot: run the real test; spheres carry no area light
                    float t_s;
                    blocks = sphere_slot_test(sc.spheres, al_, tl.ox, tl.oy, tl.oz, w.lt_d[(size_t)(__ffs(mask) - 1) * w.cap + pos],
                                              tl.t_max, &t_s);
                } else if (target_light >= 0 && al_ >= 0 && al_ == target_light) {
                    blocks = false;  // bvh.rs:269-280: the target light's own emissive triangles do not occlude
                }
                if (blocks) { occluded = true; tl.stop(sbase); }
            })
            if (live && !need_ray && !tl.wants_box()) {  // this shadow ray is done
                if (!occluded) radiance = radiance + contribution;
                mask &= mask - 1;
                if (mask) need_ray = true;
                else { finish_path(); live = false; }
            }
            const int tracing = __popc(__ballot_sync(0xffffffffu, live && !need_ray));
            if (tracing == 0 || tracing < kRefillBelow) {
                // leave to reload unless nothing could be reloaded (queue exhausted and no lane waits for its next light)
                if (tracing == 0 || !exhausted || __ballot_sync(0xffffffffu, need_ray)) break;
            }
        }
    }
    const unsigned long long sum_nodes = warp_sum((unsigned long long)tl.n_tests), sum_tris = warp_sum((unsigned long long)tl.n_tris);
    const unsigned long long sum_rays = warp_sum((unsigned long long)n_rays);
    if (lane == 0 && (sum_nodes | sum_tris | sum_rays)) {
        atomicAdd(&w.totals->any_nodes, sum_nodes);
        atomicAdd(&w.totals->any_tris, sum_tris);
        atomicAdd(&w.totals->shadow_rays, sum_rays);
    }
}

// ---- whitted stack ------------------------------------------------------------------------------------
struct StackEntry {
    V3 o, d;
    RGB weight;
    uint32_t flags;  // depth | specular
};
__device__ __forceinline__ void stack_push(const Wave& w, uint32_t path, const StackEntry& e) {
    const uint32_t top = w.stack_top[path];
    float4* base = w.stack + ((size_t)top * w.cap + path) * 3;
    base[0] = make_float4(e.o.x, e.o.y, e.o.z, e.d.x);
    base[1] = make_float4(e.d.y, e.d.z, e.weight.r, e.weight.g);
    base[2] = make_float4(e.weight.b, __uint_as_float(e.flags), 0.0f, 0.0f);
    w.stack_top[path] = top + 1;
}
// Pops the next pending node of the path's tree. Returns false when the tree is done.
__device__ __forceinline__ bool stack_pop(const Wave& w, uint32_t path, StackEntry* e) {
    const uint32_t top = w.stack_top[path];
    if (top == 0) return false;
    const float4* base = w.stack + ((size_t)(top - 1) * w.cap + path) * 3;
    const float4 a = base[0], b = base[1], c = base[2];
    w.stack_top[path] = top - 1;
    e->o = mk(a.x, a.y, a.z);
    e->d = mk(a.w, b.x, b.y);
    e->weight = rgb(b.z, b.w, c.x);
    e->flags = __float_as_uint(c.y);
    return true;
}
// Writes a tree node as the path's next ray (the sampler dimension `dim` carries on: the reference shares one sampler
// through the recursion).
__device__ __forceinline__ void stream_node(const Wave::Stream& st, uint32_t pos, const StackEntry& e, uint32_t dim, unsigned long long rng) {
    st.ray_o[pos] = make_float4(e.o.x, e.o.y, e.o.z, __int_as_float(0x7f800000));
    st.ray_d[pos] = make_float4(e.d.x, e.d.y, e.d.z, 0.0f);
    st.beta[pos] = make_float4(e.weight.r, e.weight.g, e.weight.b, __uint_as_float(e.flags | kFlagAlive | (dim << kDimShift)));
    st.rng[pos] = rng;
}

// ---- classify: miss handling + compaction by material ("ray-queue sort/compaction pass") ---------------
// K items per thread and block round (K * blockDim.x rays per global atomic): the queue counters are single addresses, and
// same-address atomics, not bandwidth, bound this kernel. Whitted's tree walk re-queues rays here and runs with K = 1.
#ifndef YK_CLASSIFY_ITEMS
#define YK_CLASSIFY_ITEMS 8
#endif
template <int K, bool WHITTED>
__global__ void k_classify(DevScene sc, Wave w, RenderCfg cfg, Batch bt, const uint32_t* queue, int b, IterCounters* cur,
                           IterCounters* nxt, int first_iteration, uint32_t* q_next) {
    const uint32_t n = cur->n_active;
    const uint32_t per_round = gridDim.x * blockDim.x * K;
    const uint32_t rounds = (n + per_round - 1) / per_round;
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t block_first = (r * gridDim.x + blockIdx.x) * blockDim.x * K;
        if (block_first >= n) break;  // block-uniform
        uint32_t idx[K], path[K], hit_slot[K];
        int key[K];
        StackEntry node[WHITTED ? K : 1];
        uint32_t node_dim[WHITTED ? K : 1];
        unsigned long long hh = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t i = block_first + k * blockDim.x + threadIdx.x;
            idx[k] = i;
            path[k] = 0; hit_slot[k] = kMiss; key[k] = -1;
            if (i >= n) continue;
            path[k] = queue ? queue[i] : i;
            const uint2 h = w.hit[i];
            hit_slot[k] = h.y;
            uint32_t orig = 0xffffffffu;
            if (h.y != kMiss) {
                key[k] = (int)((__float_as_uint(__ldg(&sc.tris[3 * h.y + 1]).w) >> 28) & 3u);
                if (first_iteration) orig = __float_as_uint(__ldg(&sc.tris[3 * h.y + 2]).w);
            } else if (first_iteration != 2) {  // (2 = debug integrators: their li() returns no background)
                // path.rs:155-160 / whitted.rs:174: background weighted by the throughput / node weight
                const float4 bw = w.st[b].beta[i];
                float4 L = w.L[path[k]];
                L.x = L.x + bw.x * sc.background[0];
                L.y = L.y + bw.y * sc.background[1];
                L.z = L.z + bw.z * sc.background[2];
                w.L[path[k]] = L;
                if (WHITTED) {
                    node_dim[WHITTED ? k : 0] = __float_as_uint(bw.w) >> kDimShift;
                    if (stack_pop(w, path[k], &node[WHITTED ? k : 0])) key[k] = 4;
                }
            }
            if (first_iteration) {
                const uint32_t si = bt.div_jobs.div(path[k]);
                const Job job = bt.jobs[path[k] - si * bt.n_jobs];
                const uint32_t sample = job.sample_begin + bt.sample_off + si;
                hh += mix_hit(job.x, job.y, sample, orig);
                if (cfg.hit_ids && sample == cfg.aux_sample) cfg.hit_ids[(size_t)job.y * cfg.res_x + job.x] = (int32_t)orig;
            }
        }
        uint32_t* const queues[5] = {w.q_mat, w.q_mat + (size_t)w.cap, w.q_mat + (size_t)2 * w.cap, w.q_mat + (size_t)3 * w.cap, q_next};
        uint32_t* const counters[5] = {&cur->mat[0], &cur->mat[1], &cur->mat[2], &cur->mat[3], &nxt->n_active};
        uint32_t pos[K];
        block_scatter_multi<5, K>(key, path, queues, counters, pos);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (key[k] >= 0 && key[k] < 4) {
                w.q_mat_tri[(size_t)key[k] * w.cap + pos[k]] = hit_slot[k];
                w.q_mat_slot[(size_t)key[k] * w.cap + pos[k]] = idx[k];
            } else if (WHITTED && key[k] == 4) {
                stream_node(w.st[b ^ 1], pos[k], node[WHITTED ? k : 0], node_dim[WHITTED ? k : 0], w.st[b].rng[idx[k]]);
            }
        }
        if (first_iteration) {
            hh = warp_sum(hh);
            if ((threadIdx.x & 31) == 0 && hh) atomicAdd(&w.totals->hit_hash, hh);
        }
    }
}

// ---- surface set-up: Triangle::intersect's SurfaceInteraction part (triangle.rs:141-226) ----------------
__device__ __forceinline__ void make_surface(const DevScene& sc, uint32_t tri, V3 o, V3 d, Surface* si, uint32_t* material) {
    const float4 a4 = __ldg(&sc.tris[3 * tri]), b4 = __ldg(&sc.tris[3 * tri + 1]), c4 = __ldg(&sc.tris[3 * tri + 2]);
    const V3 p0 = mk(a4.x, b4.x, c4.x), p1 = mk(a4.y, b4.y, c4.y), p2 = mk(a4.z, b4.z, c4.z);  // stored transposed
    const uint32_t packed = __float_as_uint(b4.w);
    const uint32_t flags = (packed >> 24) & 0xfu;
    *material = packed & 0xffffffu;
    if (flags & YK_TRI_IS_SPHERE) {
        sphere_surface(sc.spheres[-2 - __float_as_int(a4.w)], o, d, si);
        return;
    }
    // Barycentrics: re-run the (deterministic) triangle test that the traversal accepted.
    TriRay tr;
    tr.setup(d);
    TriHit h{0, 0, 0, 0};
    tri_test(tr, o, __int_as_float(0x7f800000), p0, p1, p2, &h);
    V2 uv0{0.0f, 0.0f}, uv1{1.0f, 0.0f}, uv2{1.0f, 1.0f};  // triangle.rs:143-155
    if (flags & YK_TRI_HAS_UVS) {
        const float* u = sc.uvs + (size_t)tri * 6;
        uv0 = {u[0], u[1]}; uv1 = {u[2], u[3]}; uv2 = {u[4], u[5]};
    }
    const float du02 = uv0.x - uv2.x, dv02 = uv0.y - uv2.y, du12 = uv1.x - uv2.x, dv12 = uv1.y - uv2.y;
    const V3 dp02 = p0 - p2, dp12 = p1 - p2;
    const float uv_det = du02 * dv12 - dv02 * du12;
    V3 dpdu;
    if (uv_det == 0.0f) {
        V3 unused;
        frame_from(unit(cross64(p2 - p0, p1 - p0)), &dpdu, &unused);
    } else {
        const float inv = 1.0f / uv_det;
        dpdu = (dp02 * dv12 - dp12 * dv02) * inv;
    }
    si->p = p0 * h.b0 + p1 * h.b1 + p2 * h.b2;
    si->uv = {uv0.x * h.b0 + uv1.x * h.b1 + uv2.x * h.b2, uv0.y * h.b0 + uv1.y * h.b1 + uv2.y * h.b2};
    si->wo = -d;
    si->area_light = __float_as_int(a4.w);
    V3 n = unit(cross64(dp02, dp12));
    if (flags & YK_TRI_SWAPS_HANDEDNESS) n = -n;
    si->n = n;
    si->sh_n = n;
    si->sh_dpdu = dpdu;
    if (flags & YK_TRI_HAS_NORMALS) {  // triangle.rs:197-224 + set_shading_geometry, interaction.rs:126-132
        const float* nn = sc.normals + (size_t)tri * 9;
        const V3 n0 = mk(nn[0], nn[1], nn[2]), n1 = mk(nn[3], nn[4], nn[5]), n2 = mk(nn[6], nn[7], nn[8]);
        V3 ns = unit(n0 * h.b0 + n1 * h.b1 + n2 * h.b2);
        if (dot0(ns, ns) > 0.0f) ns = unit(ns);
        else ns = si->n;
        V3 ss = unit(dpdu);
        V3 ts = cross64(ss, ns);
        if (dot0(ts, ts) > 0.0f) {
            ts = unit(ts);
            ss = cross64(ts, ns);
        } else {
            frame_from(ns, &ss, &ts);
        }
        si->sh_n = unit(cross64(ss, ts));
        si->n = flip_toward_n(si->n, si->sh_n);
        si->sh_dpdu = ss;
    }
}

// textures/constant.rs:23-30, textures/image_texture.rs:81-111
__device__ __noinline__ RGB tex_image_eval(const DevTexture& t, V2 uv);
__device__ __forceinline__ RGB tex_eval(const DevScene& sc, int32_t index, V2 uv) {
    const DevTexture& t = sc.textures[index];
    if (t.kind == YK_TEX_CONSTANT) return rgb(t.value[0], t.value[1], t.value[2]);
    return tex_image_eval(t, uv);
}
__device__ __noinline__ RGB tex_image_eval(const DevTexture& t, V2 uv) {
    float sx = uv.x - truncf(uv.x), sy = uv.y - truncf(uv.y);
    if (sx < 0.0f) sx = 1.0f + sx;
    if (sy < 0.0f) sy = 1.0f + sy;
    sy = 1.0f - sy;
    sx = sx * (float)t.width - 0.5f;
    sy = sy * (float)t.height - 0.5f;
    const uint32_t ix = sx > 0.0f ? (uint32_t)sx : 0u, iy = sy > 0.0f ? (uint32_t)sy : 0u;
    const float* px = t.texels + ((size_t)iy * t.width + ix) * 3;
    return rgb(__ldg(px), __ldg(px + 1), __ldg(px + 2));
}

__device__ __forceinline__ float roughness_to_alpha(float r) {  // trowbridge_reitz.rs:22-30
    const float x = (float)log((double)fmaxf(r, 0.001f));
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

template <uint32_t KIND>
__device__ __forceinline__ void make_bsdf(const DevScene& sc, const DevMaterial& m, const Surface& si, Bsdf* b) {
    b->kind = KIND;
    b->empty = false;
    b->ng = si.n;
    b->ns = si.sh_n;
    b->ss = unit(si.sh_dpdu);
    b->ts = cross64(b->ns, b->ss);
    b->c1 = gray(0.0f);
    b->p0 = 0.0f;
    b->p1 = -1.0f;
    if (KIND == YK_MAT_MATTE) {  // matte.rs:22-40, oren_nayar.rs:18-25
        b->c0 = tex_eval(sc, m.tex[0], si.uv);
        const float sigma = tex_eval(sc, m.tex[1], si.uv).r;
        b->empty = black(b->c0);
        if (sigma != 0.0f) {
            const float s2 = sigma * sigma;
            b->p0 = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
            b->p1 = 0.45f * s2 / (s2 + 0.09f);
            if (b->p1 < 0.0f) b->p1 = 0.0f;  // cannot happen for real sigma; keeps the Lambertian tag (p1 < 0) unambiguous
        }
    } else if (KIND == YK_MAT_GLASS) {  // glass.rs:27-45
        b->c0 = tex_eval(sc, m.tex[0], si.uv);
        b->c1 = tex_eval(sc, m.tex[1], si.uv);
        b->p0 = m.eta;
    } else if (KIND == YK_MAT_METAL) {  // metal.rs:34-61, trowbridge_reitz.rs:16-20
        b->c0 = tex_eval(sc, m.tex[0], si.uv);
        b->c1 = tex_eval(sc, m.tex[1], si.uv);
        if (m.const_alpha >= 0.0f) b->p0 = m.const_alpha;
        else {
            float r = tex_eval(sc, m.tex[2], si.uv).r;
            if (m.remap) r = roughness_to_alpha(r);
            b->p0 = fmaxf(r, 0.001f);
        }
    } else {  // glossy.rs:32-58 (alpha = roughness^2)
        b->c0 = tex_eval(sc, m.tex[0], si.uv);
        if (m.const_alpha >= 0.0f) b->p0 = m.const_alpha;
        else {
            float r = tex_eval(sc, m.tex[1], si.uv).r;
            if (m.remap) r = roughness_to_alpha(r);
            b->p0 = fmaxf(r * r, 0.001f);
        }
    }
}

// Light::sample_li for the four light kinds (lights/*.rs). Returns false when no visibility test exists.
struct LightSample {
    V3 l;
    RGB li;
    float pdf;
    bool has_vis;
    Ray vis;
    int vis_light;
};
__device__ __forceinline__ void sample_light(const yk_light& L, int index, const Surface& si, V2 u, LightSample* s) {
    s->vis_light = -1;
    s->pdf = 1.0f;
    s->has_vis = true;
    const RGB I = rgb(L.i[0], L.i[1], L.i[2]);
    const V3 lp = mk(L.p[0], L.p[1], L.p[2]);
    if (L.kind == YK_LIGHT_POINT) {  // point_light.rs:27-49
        const V3 to = lp - si.p;
        const float d2 = dot0(to, to);
        s->li = I / d2;
        s->l = to / sqrtf(d2);
        s->vis = spawn_ray_to(si.p, si.n, lp);
    } else if (L.kind == YK_LIGHT_SPOT) {  // spot_light.rs:38-80
        const V3 to = lp - si.p;
        const float d2 = dot0(to, to);
        s->l = to / sqrtf(d2);
        const float ct = unit(xf_vec(L.world_to_light, -s->l)).z;
        float fall;
        if (ct < L.cos_total_width) fall = 0.0f;
        else if (ct > L.cos_falloff_start) fall = 1.0f;
        else {
            const float dl = (ct - L.cos_total_width) / (L.cos_falloff_start - L.cos_total_width);
            fall = (dl * dl) * (dl * dl);
        }
        s->li = I * fall / d2;
        s->has_vis = !black(s->li);
        s->vis = spawn_ray_to(si.p, si.n, lp);
    } else if (L.kind == YK_LIGHT_RECT) {  // rectangular_light.rs:46-72
        const V3 p = xf_point(L.sample_to_world, mk(u.x, 0.0f, u.y));
        const V3 n = xf_normal(L.sample_to_world_inv, mk(0.0f, -1.0f, 0.0f));
        const V3 wi = unit(p - si.p);
        const float c = dotn(n, -wi);
        s->li = c > 0.0f ? I : gray(0.0f);
        s->l = wi;
        s->vis = spawn_ray_to(si.p, si.n, p);
        s->vis_light = index;
        const V3 dp = si.p - p;
        s->pdf = dot0(dp, dp) / (fabsf(c) * L.area);
    } else {  // distant_light.rs:24-43
        s->li = I;
        s->l = lp;
        s->vis = spawn_ray_to(si.p, si.n, si.p + lp * 10000.0f);
    }
}

// ---- shading: one kernel instance per material kind ----------------------------------------------------
// Covers Material::compute_scattering_functions, the light fold (path.rs:102-119 / whitted.rs:109-126), the
// emitted term, BSDF sampling + throughput update + Russian roulette (path.rs:121-171), and the specular
// recursion of whitted.rs:132-170 flattened onto a per-sample DFS stack (children inherit weight * f * |cos|).
// Radiance is not summed here: each light that needs a visibility test leaves its shadow ray and contribution in
// lt_*, and k_trace_shadow adds the unoccluded terms in light order. Surviving paths are appended to the next
// active queue (one atomic per block).
template <uint32_t KIND, bool PATH>
__global__ void __launch_bounds__(kShadeThreads, YK_SHADE_MIN_BLOCKS) k_shade(DevScene sc, Wave w, RenderCfg cfg, Batch bt, const uint32_t* queue,
                                                                               const uint32_t* queue_tri, const uint32_t* queue_slot, int b,
                                                                               IterCounters* cur, IterCounters* nxt, uint32_t* q_next) {
    const uint32_t n = cur->mat[KIND];
    uint32_t g_base = 0;  // shading position of this kind's first queue entry (classify has finished: the counts are final)
#pragma unroll
    for (uint32_t k = 0; k < KIND; ++k) g_base += cur->mat[k];
    const uint32_t rounds = (n + gridDim.x * blockDim.x - 1) / (gridDim.x * blockDim.x);
    for (uint32_t r = 0; r < rounds; ++r) {
        const uint32_t block_first = (r * gridDim.x + blockIdx.x) * blockDim.x;
        if (block_first >= n) break;  // block-uniform
        const uint32_t i = block_first + threadIdx.x;
        bool alive = false;
        uint32_t path = 0;
        // the survivor's state for the next bounce, written after the compaction assigns its position
        float4 nx_o = make_float4(0, 0, 0, 0), nx_d = make_float4(0, 0, 0, 0), nx_beta = make_float4(0, 0, 0, 0);
        unsigned long long nx_rng = 0;
        if (i < n) {
            path = queue[i];
            const uint32_t g = g_base + i;
            w.sh_path[g] = path;
            const uint32_t hit_slot = queue_tri[i], slot = queue_slot[i];
            const float4 ro = w.st[b].ray_o[slot], rd = w.st[b].ray_d[slot];
            const V3 o = f4v(ro), d = f4v(rd);
            Surface si;
            uint32_t mat_index;
            make_surface(sc, hit_slot, o, d, &si, &mat_index);
            Bsdf bsdf;
            make_bsdf<KIND>(sc, sc.materials[mat_index], si, &bsdf);

            const float4 beta4 = w.st[b].beta[slot];
            RGB beta = rgb(beta4.x, beta4.y, beta4.z);
            const uint32_t flags = __float_as_uint(beta4.w) & kFlagMask;
            const uint32_t depth = flags & kDepthMask;  // path: bounces so far; whitted: node depth
            const bool was_specular = (flags & kFlagSpecular) != 0;

            const uint32_t sample_i = bt.div_jobs.div(path), job_i = path - sample_i * bt.n_jobs;
            const Job job = bt.jobs[job_i];
            SamplerState smp;
            smp.rng.state = w.st[b].rng[slot];
            smp.rng.inc = job.rng_inc;
            smp.dim = __float_as_uint(beta4.w) >> kDimShift;
            smp.px = job.x;
            smp.py = job.y;
            smp.index = job.sample_begin + bt.sample_off + sample_i;
            smp.job = job_i;

            // Light fold: every light consumes one get_2d whether it is used or not (path.rs:103).
            uint32_t shadow_mask = 0;
            for (uint32_t k = 0; k < sc.n_lights; ++k) {
                const V2 u = smp.get_2d(cfg.sampler);
                LightSample ls;
                sample_light(sc.lights[k], (int)k, si, u, &ls);
                if (!black(ls.li)) {
                    const RGB f = bsdf.f(si.wo, ls.l);
                    if (ls.has_vis && !black(f)) {
                        const RGB c = f * ls.li * clamp01ish(dotn(si.sh_n, ls.l), 0.0f, 1.0f) / ls.pdf;
                        const size_t ref = (size_t)k * w.cap + g;
                        w.lt_o[ref] = make_float4(ls.vis.o.x, ls.vis.o.y, ls.vis.o.z, c.r);
                        w.lt_d[ref] = make_float4(ls.vis.d.x, ls.vis.d.y, ls.vis.d.z, c.g);
                        w.lt_c[ref] = make_float2(c.b, __int_as_float(ls.vis_light));
                        shadow_mask |= 1u << k;
                    }
                }
            }

            // Emitted radiance: interaction.rs:134-138 + rectangular_light.rs:74-81
            RGB le = gray(0.0f);
            // The integrators pass -ray.d here and (Path) to sample_f, but si.wo to Bsdf::f; the two differ for spheres, whose
            // si.wo went through object_to_world once more (sphere.rs:116, interaction.rs:155).
            const V3 wo_ray = -d;
            if (si.area_light >= 0 && dotn(si.n, wo_ray) > 0.0f) {
                const yk_light& al = sc.lights[si.area_light];
                le = rgb(al.i[0], al.i[1], al.i[2]);
            }
            const bool add_le = depth == 0 || was_specular;

            uint32_t new_flags = 0;
            if (PATH) {
                // path.rs:121-129 — beta multiplies the emitted term here and again in the fold (reference quirk)
                const RGB extra = add_le ? beta * le : gray(0.0f);
                w.pend_extra[g] = make_float4(extra.r, extra.g, extra.b, __uint_as_float(shadow_mask));
                w.pend_beta[g] = make_float4(beta.r, beta.g, beta.b, (depth > 0 && cfg.has_clamp) ? 1.0f : 0.0f);
                const Bsdf::Sample s = bsdf.sample_f(wo_ray, smp.get_2d(cfg.sampler), BX_ALL);  // path.rs:131-137 (wo = -ray.d)
                if (!(black(s.f) || s.pdf == 0.0f)) {
                    alive = true;
                    const bool spec = (s.type & BX_SPECULAR) != 0;
                    beta = beta * (s.f * fabsf(dotn(s.wi, si.sh_n)) / s.pdf);
                    const Ray nr = spawn_ray(si.p, si.n, s.wi);
                    if (depth > 3) {  // Russian roulette, path.rs:163-169
                        const float q = fmaxf(1.0f - beta.g, 0.05f);
                        if (smp.get_1d(cfg.sampler) < q) alive = false;
                        else beta = beta * (gray(1.0f) / (1.0f - q));
                    }
                    const uint32_t bounces = depth + 1;
                    if (bounces >= cfg.max_depth) alive = false;
                    new_flags = (bounces & kDepthMask) | (spec ? kFlagSpecular : 0u);
                    nx_o = make_float4(nr.o.x, nr.o.y, nr.o.z, nr.t_max);
                    nx_d = make_float4(nr.d.x, nr.d.y, nr.d.z, 0.0f);
                    nx_beta = make_float4(beta.r, beta.g, beta.b, __uint_as_float(new_flags | kFlagAlive | (smp.dim << kDimShift)));
                }
            } else {
                // whitted.rs:128-170
                const RGB extra = add_le ? le : gray(0.0f);
                w.pend_extra[g] = make_float4(extra.r, extra.g, extra.b, __uint_as_float(shadow_mask));
                w.pend_beta[g] = make_float4(beta.r, beta.g, beta.b, 0.0f);
                StackEntry child[2];
                int n_child = 0;
                if (KIND == YK_MAT_GLASS && depth + 1 < cfg.max_depth) {  // only Glass owns SPECULAR lobes
                    const uint32_t wants[2] = {BX_SPECULAR | BX_REFLECTION, BX_SPECULAR | BX_TRANSMISSION};
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const Bsdf::Sample s = bsdf.sample_f(si.wo, V2{0.0f, 0.0f}, wants[c]);
                        if (s.type == 0u) continue;  // BxdfType::NONE: no ray, no radiance
                        const Ray nr = spawn_ray(si.p, si.n, s.wi);
                        StackEntry e;
                        e.o = nr.o;
                        e.d = nr.d;
                        e.weight = beta * s.f * fabsf(dotn(s.wi, si.sh_n));
                        e.flags = ((depth + 1) & kDepthMask) | ((s.type & BX_SPECULAR) ? kFlagSpecular : 0u);
                        child[n_child++] = e;
                    }
                }
                if (n_child == 2) stack_push(w, path, child[1]);  // transmission waits until the reflection subtree is done
                StackEntry e = child[0];
                alive = n_child >= 1 || stack_pop(w, path, &e);
                nx_o = make_float4(e.o.x, e.o.y, e.o.z, __int_as_float(0x7f800000));
                nx_d = make_float4(e.d.x, e.d.y, e.d.z, 0.0f);
                nx_beta = make_float4(e.weight.r, e.weight.g, e.weight.b, __uint_as_float(e.flags | kFlagAlive | (smp.dim << kDimShift)));
            }
            nx_rng = smp.rng.state;
        }
        uint32_t* const queues[1] = {q_next};
        uint32_t* const counters[1] = {&nxt->n_active};
        const uint32_t npos = block_scatter<1>(alive ? 0 : -1, path, queues, counters);
        if (alive) {  // a finished path's ray / throughput / sampler state is never read again
            const Wave::Stream& out = w.st[b ^ 1];
            out.ray_o[npos] = nx_o;
            out.ray_d[npos] = nx_d;
            out.beta[npos] = nx_beta;
            out.rng[npos] = nx_rng;
        }
    }
}

// ---- debug integrators (bvh_heatmap.rs, geometry_normals.rs, shading_normals.rs, shading_uvs.rs) --------
__global__ void k_debug_shade(DevScene sc, Wave w, RenderCfg cfg, uint32_t n) {
    const uint32_t path = blockIdx.x * blockDim.x + threadIdx.x;
    if (path >= n) return;
    const uint2 h = w.hit[path];
    RGB c = gray(0.0f);
    if (cfg.integrator == YK_INTEGRATOR_BVH_INTERSECTIONS) {
        const uint2 cnt = w.bvh_counts[path];
        c = rgb((float)cnt.x, (float)cnt.y, h.y != kMiss ? (float)cnt.y : 0.0f);
    } else if (h.y != kMiss) {
        Surface si;
        uint32_t m;
        make_surface(sc, h.y, f4v(w.st[0].ray_o[path]), f4v(w.st[0].ray_d[path]), &si, &m);  // first bounce: slot == path
        if (cfg.integrator == YK_INTEGRATOR_GEOMETRY_NORMALS) c = rgb(si.n.x, si.n.y, si.n.z) / 2.0f + gray(0.5f);
        else if (cfg.integrator == YK_INTEGRATOR_SHADING_NORMALS) c = rgb(si.sh_n.x, si.sh_n.y, si.sh_n.z) / 2.0f + gray(0.5f);
        else c = rgb(si.uv.x, si.uv.y, 0.0f);
    }
    w.L[path] = make_float4(c.r, c.g, c.b, 0.0f);
}

// ---- film --------------------------------------------------------------------------------------------
// `color += li` over ascending sample index (integrators/mod.rs:172), carried across batches in `accum`.
__global__ void k_film_accumulate(Wave w, Batch bt, float* accum, uint32_t res_x) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= bt.n_jobs) return;
    const Job job = bt.jobs[j];
    float* a = accum + ((size_t)job.y * res_x + job.x) * 3;
    float r = a[0], g = a[1], b = a[2];
    for (uint32_t s = 0; s < bt.n_samples; ++s) {
        const float4 L = w.L[(size_t)s * bt.n_jobs + j];
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
    const float4 L = w.L[i];
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

// ---- scene packing (yk_scene_create): the reference-layout arrays are repacked on the device ---------------------------
__global__ void k_scene_interior_flags(const yk_bvh_node* nodes, uint32_t n, uint32_t* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = nodes[i].is_leaf ? 0u : 1u;
}
// rec[i] = number of interior nodes before node i: an interior node's record index; i - rec[i] = a leaf's table index.
__global__ void k_scene_records(const yk_bvh_node* nodes, const uint32_t* rec, uint32_t n, int packed_leaves, uint2* leaf_table, float4* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const yk_bvh_node nd = nodes[i];
    if (nd.is_leaf) return;
    const uint32_t kids[2] = {i + 1, nd.offset};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const yk_bvh_node ch = nodes[kids[k]];
        uint32_t ref;
        if (!ch.is_leaf) ref = kRefInterior | ((uint32_t)ch.split_axis << 29) | rec[kids[k]];
        else if (packed_leaves) ref = ((uint32_t)(ch.shape_count - 1) << kLeafFirstBits) | ch.offset;
        else {
            ref = kids[k] - rec[kids[k]];
            leaf_table[ref] = make_uint2(ch.offset, ch.shape_count);
        }
        out[(size_t)rec[i] * 4 + 2 * k] = make_float4(ch.p_min[0], ch.p_min[1], ch.p_min[2], __uint_as_float(ref));
        out[(size_t)rec[i] * 4 + 2 * k + 1] = make_float4(ch.p_max[0], ch.p_max[1], ch.p_max[2], 0.0f);
    }
}
__global__ void k_scene_tris(const float* verts, const uint32_t* orig, const uint32_t* mat, const int32_t* alight, const uint8_t* flags,
                             const int32_t* sphere, const uint8_t* mat_kind, uint32_t n, float4* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = mat[i], f = flags[i];
    // material index | YK_TRI_* flags << 24 | material kind << 28 (so the material sort needs no material table look-up)
    const uint32_t packed = m | ((f & 0xfu) << 24) | (((uint32_t)mat_kind[m] & 3u) << 28);
    const float fp = __uint_as_float(packed), fi = __uint_as_float(orig[i]);
    if (f & YK_TRI_IS_SPHERE) {  // NaN vertex lanes + (-2 - sphere index) where triangles keep their area light
        const float qnan = __int_as_float(0x7fc00000);
        out[3 * (size_t)i] = make_float4(qnan, qnan, qnan, __int_as_float(-2 - sphere[i]));
        out[3 * (size_t)i + 1] = make_float4(qnan, qnan, qnan, fp);
        out[3 * (size_t)i + 2] = make_float4(qnan, qnan, qnan, fi);
        return;
    }
    const float* v = verts + (size_t)i * 9;
    out[3 * (size_t)i] = make_float4(v[0], v[3], v[6], __int_as_float(alight[i]));
    out[3 * (size_t)i + 1] = make_float4(v[1], v[4], v[7], fp);
    out[3 * (size_t)i + 2] = make_float4(v[2], v[5], v[8], fi);
}

}  // namespace

// =====================================================================================================
// Host side: context, scene upload, wavefront driver.
// =====================================================================================================
constexpr int kMaxPipes = 2;   // batches in flight on separate streams (their kernels overlap on the SMs)
constexpr int kRing = 2;       // batches queued per pipe before the host waits for the oldest
constexpr int kTimedStages = 5;  // events per bounce: before/after closest, after classify, after shading, after shadow

// One asynchronous wavefront lane: its own stream, path state, bounce counters and timing events.
struct Pipe {
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    Wave wave{};
    std::vector<void*> wave_allocs;
    uint32_t wave_cap = 0, wave_lights = 0, wave_stack = 0;
    IterCounters* d_ctr = nullptr;   // two entries, alternating per bounce
    IterCounters* h_ctr = nullptr;   // pinned: read-back for the integrators whose bounce count is unbounded (Whitted)
    Totals* h_totals = nullptr;      // pinned
    SampleJump* d_jumps = nullptr;   // sampler seek table of the batch being queued (kMaxBatchSamples entries)
    Job* d_jobs = nullptr;           // the pixel jobs of the pipe's current pixel group
    size_t jobs_cap = 0;
    uint32_t* d_dim_hash = nullptr;  // SamplerCfg::hash_table of the pipe's current pixel group
    size_t dim_hash_cap = 0;
    size_t hash_group = (size_t)-1;
    struct Slot {
        cudaEvent_t done = nullptr;
        std::vector<cudaEvent_t> ev;  // kTimedStages per bounce
        uint32_t n_iters = 0;
        uint64_t n_paths = 0;
        bool busy = false;
    } slot[kRing];
    uint64_t n_batches = 0;
    int timing = 2;  // the context's stage_timing when the queued batches were recorded
};

struct yk_context {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {};
    Pipe pipe[kMaxPipes];
    yk_tile* d_tiles = nullptr;              // the tile list of the current render and the prefix sum of its tile areas
    unsigned long long* d_tile_off = nullptr;
    size_t tiles_cap = 0;
    float* d_accum = nullptr;
    float* d_film = nullptr;
    int32_t* d_hit_ids = nullptr;
    size_t film_cap = 0;
    int occ_trace_closest = 0, occ_trace_any = 0;
    int n_pipes_env = 0;  // YK_PIPES override (development)
    uint64_t mem_budget = 0;  // bytes of wavefront state per pipe the default batch size may use (set at the first render)
    int stage_timing = 1;  // CUDA events per bounce: 1 = around the closest-hit kernel (the roofline figure), 2 = every stage
                           // (costs ~1.5 % of a Cornell render), 0 = none; environment variable YK_STAGE_TIMING
};

struct yk_scene {
    yk_context* ctx = nullptr;
    int device = 0;
    uint32_t material_kinds = 0;  // bit k set: some triangle's material has kind k
    DevScene dev{};
    std::vector<void*> allocs;
};

namespace {

template <class T>
int dev_alloc(std::vector<void*>& bag, T** out, size_t count) {
    void* p = nullptr;
    CUDA_TRY(cudaMalloc(&p, std::max<size_t>(count, 1) * sizeof(T)));
    bag.push_back(p);
    *out = (T*)p;
    return YK_OK;
}
template <class T>
int dev_upload(std::vector<void*>& bag, const T** out, const T* src, size_t count) {
    T* p = nullptr;
    int rc = dev_alloc(bag, &p, count);
    if (rc != YK_OK) return rc;
    if (count) CUDA_TRY(cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return YK_OK;
}
void free_bag(std::vector<void*>& bag) {
    for (void* p : bag) cudaFree(p);
    bag.clear();
}

// trowbridge_reitz.rs:22-30 on the host (libm logf, as the reference's f32::ln)
float host_roughness_to_alpha(float r) {
    const float x = logf(fmaxf(r, 0.001f));
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}

int ensure_wave(Pipe* p, uint32_t cap, uint32_t n_lights, uint32_t stack_entries) {
    if (p->wave_cap == cap && p->wave_lights == n_lights && p->wave_stack == stack_entries) return YK_OK;
    free_bag(p->wave_allocs);
    p->wave_cap = 0;
    Wave w{};
    w.cap = cap;
    w.n_lights = n_lights;
    w.stack_entries = stack_entries;
    std::vector<void*>& bag = p->wave_allocs;
    const size_t nl = std::max<uint32_t>(n_lights, 1);
    int rc = YK_OK;
#define WAVE_ALLOC(field, count) \
    if ((rc = dev_alloc(bag, &w.field, (size_t)(count))) != YK_OK) { free_bag(bag); return rc; }
    for (int k = 0; k < 2; ++k) {
        WAVE_ALLOC(st[k].ray_o, cap) WAVE_ALLOC(st[k].ray_d, cap) WAVE_ALLOC(st[k].beta, cap) WAVE_ALLOC(st[k].rng, cap)
    }
    WAVE_ALLOC(hit, cap) WAVE_ALLOC(bvh_counts, cap) WAVE_ALLOC(L, cap)
    WAVE_ALLOC(sh_path, cap) WAVE_ALLOC(pend_beta, cap) WAVE_ALLOC(pend_extra, cap)
    WAVE_ALLOC(lt_o, cap * nl) WAVE_ALLOC(lt_d, cap * nl) WAVE_ALLOC(lt_c, cap * nl)
    WAVE_ALLOC(q_active[0], cap) WAVE_ALLOC(q_active[1], cap) WAVE_ALLOC(q_mat, (size_t)4 * cap) WAVE_ALLOC(q_mat_tri, (size_t)4 * cap) WAVE_ALLOC(q_mat_slot, (size_t)4 * cap)
    WAVE_ALLOC(totals, 1)
    if (stack_entries) {
        WAVE_ALLOC(stack, (size_t)stack_entries * cap * 3)
        WAVE_ALLOC(stack_top, cap)
    }
#undef WAVE_ALLOC
    p->wave = w;
    p->wave_cap = cap;
    p->wave_lights = n_lights;
    p->wave_stack = stack_entries;
    return YK_OK;
}

struct Timers {
    double closest = 0, any = 0, shade = 0;
    uint64_t launches = 0, closest_launches = 0;
};

int grid_for(uint32_t n, int threads, int max_blocks) {
    const uint32_t need = (n + threads - 1) / threads;
    return (int)std::max<uint32_t>(1u, std::min<uint32_t>(need, (uint32_t)max_blocks));
}

// Waits for the batch in ring slot `k` of the pipe and folds its stage timings into `tm`.
int retire_slot(Pipe* p, int k, Timers* tm, uint64_t* done_paths) {
    Pipe::Slot& sl = p->slot[k];
    if (!sl.busy) return YK_OK;
    CUDA_TRY(cudaEventSynchronize(sl.done));
    for (uint32_t i = 0; i < sl.n_iters; ++i) {
        cudaEvent_t* e = &sl.ev[(size_t)i * kTimedStages];
        float ms = 0;
        if (p->timing > 0) { cudaEventElapsedTime(&ms, e[0], e[1]); tm->closest += ms; }
        if (p->timing > 1) {
            cudaEventElapsedTime(&ms, e[2], e[3]); tm->shade += ms;
            cudaEventElapsedTime(&ms, e[3], e[4]); tm->any += ms;
        }
    }
    *done_paths += sl.n_paths;
    sl.busy = false;
    return YK_OK;
}

// One batch on one pipe: raygen, the bounce loop, and the per-batch film step, all asynchronous on the pipe's stream.
// Path tracing runs exactly max_depth bounces (every queue length stays on the device); Whitted's tree walk has no
// such bound, so it reads the next bounce's ray count back once per bounce.
int run_batch(yk_context* c, Pipe* p, const yk_scene* sc, const RenderCfg& cfg, const Batch& bt, uint32_t first_sample, bool accumulate_film,
              float* d_film, Timers* tm, uint64_t* done_paths) {
    cudaStream_t s = p->stream;
    Wave& w = p->wave;
    const int k = (int)(p->n_batches % kRing);
    int rc = retire_slot(p, k, tm, done_paths);
    if (rc != YK_OK) return rc;
    Pipe::Slot& sl = p->slot[k];
    if (!sl.done) CUDA_TRY(cudaEventCreate(&sl.done));
    sl.n_iters = 0;
    sl.n_paths = bt.n_paths;
    p->timing = c->stage_timing;
    auto stage_event = [&](uint32_t iter, int stage) -> cudaEvent_t {
        const size_t idx = (size_t)iter * kTimedStages + stage;
        while (sl.ev.size() <= idx) {
            cudaEvent_t e = nullptr;
            cudaEventCreate(&e);
            sl.ev.push_back(e);
        }
        return sl.ev[idx];
    };

    const int T = 256;
    CUDA_TRY(cudaMemsetAsync(p->d_ctr, 0, 2 * sizeof(IterCounters), s));
    k_sample_jumps<<<1, kMaxBatchSamples, 0, s>>>(first_sample, bt.n_samples, p->d_jumps);
    tm->launches += 1;
    k_raygen<<<(bt.n_paths + T - 1) / T, T, 0, s>>>(w, cfg, bt, first_sample, bt.n_samples, p->d_jumps, &p->d_ctr[0]);
    tm->launches += 1;
    const bool debug = cfg.integrator >= YK_INTEGRATOR_BVH_INTERSECTIONS;
    const bool sync_loop = cfg.integrator == YK_INTEGRATOR_WHITTED;
    const int trace_blocks_closest = c->sm_count * std::max(1, c->occ_trace_closest);
    const int trace_blocks_shadow = c->sm_count * std::max(1, c->occ_trace_any);
    const int wide_blocks = c->sm_count * 16;
    const int classify_items = cfg.integrator == YK_INTEGRATOR_WHITTED ? 1 : YK_CLASSIFY_ITEMS;
    const int classify_blocks = grid_for(bt.n_paths, T * classify_items, c->sm_count * 8);
    const int shade_blocks = grid_for(bt.n_paths, kShadeThreads, wide_blocks);
    const int closest_blocks = grid_for(bt.n_paths, kTraceThreads, trace_blocks_closest);
    const int shadow_blocks = grid_for(bt.n_paths, kTraceThreads, trace_blocks_shadow);
    uint32_t max_iters = 1;
    if (cfg.integrator == YK_INTEGRATOR_PATH) max_iters = cfg.max_depth;
    else if (sync_loop) max_iters = 0xffffffffu;
    uint32_t* q_cur = nullptr;
    int flip = 0;
    for (uint32_t iter = 0; iter < max_iters; ++iter) {
        const int b = (int)(iter & 1);  // this bounce reads stream b and writes stream b ^ 1
        IterCounters* cur = &p->d_ctr[iter & 1];
        IterCounters* nxt = &p->d_ctr[(iter + 1) & 1];
        if (iter > 0) CUDA_TRY(cudaMemsetAsync(nxt, 0, sizeof(IterCounters), s));
        if (c->stage_timing > 0) CUDA_TRY(cudaEventRecord(stage_event(iter, 0), s));
        const bool spheres = sc->dev.spheres != nullptr || sc->dev.leaf_table != nullptr;  // the generic instantiations: sphere slots, leaf table
        if (cfg.integrator == YK_INTEGRATOR_BVH_INTERSECTIONS) {
            if (spheres) k_trace_closest<true, true><<<closest_blocks, kTraceThreads, 0, s>>>(sc->dev, w, b, cur);
            else k_trace_closest<true, false><<<closest_blocks, kTraceThreads, 0, s>>>(sc->dev, w, b, cur);
        } else {
            if (spheres) k_trace_closest<false, true><<<closest_blocks, kTraceThreads, 0, s>>>(sc->dev, w, b, cur);
            else k_trace_closest<false, false><<<closest_blocks, kTraceThreads, 0, s>>>(sc->dev, w, b, cur);
        }
        if (c->stage_timing > 0) CUDA_TRY(cudaEventRecord(stage_event(iter, 1), s));
        tm->launches += 1;
        tm->closest_launches += 1;
        uint32_t* q_next = w.q_active[flip];
        if (debug) {
            k_debug_shade<<<(bt.n_paths + T - 1) / T, T, 0, s>>>(sc->dev, w, cfg, bt.n_paths);
            // primary-hit digest / id image for the debug integrators too
            k_classify<YK_CLASSIFY_ITEMS, false><<<classify_blocks, T, 0, s>>>(sc->dev, w, cfg, bt, q_cur, b, cur, nxt, 2, q_next);
            tm->launches += 2;
            for (int st = 2; st < kTimedStages; ++st) if (c->stage_timing > 1) CUDA_TRY(cudaEventRecord(stage_event(iter, st), s));
            sl.n_iters = iter + 1;
            break;
        }
        if (cfg.integrator == YK_INTEGRATOR_WHITTED) k_classify<1, true><<<classify_blocks, T, 0, s>>>(sc->dev, w, cfg, bt, q_cur, b, cur, nxt, iter == 0 ? 1 : 0, q_next);
        else k_classify<YK_CLASSIFY_ITEMS, false><<<classify_blocks, T, 0, s>>>(sc->dev, w, cfg, bt, q_cur, b, cur, nxt, iter == 0 ? 1 : 0, q_next);
        if (c->stage_timing > 1) CUDA_TRY(cudaEventRecord(stage_event(iter, 2), s));
        tm->launches += 1;
        const bool is_path = cfg.integrator == YK_INTEGRATOR_PATH;
        for (uint32_t kind = 0; kind < 4; ++kind) {
            if (!(sc->material_kinds & (1u << kind))) continue;  // no triangle of the scene has this material kind
            uint32_t* q = w.q_mat + (size_t)kind * w.cap;
            uint32_t* qt = w.q_mat_tri + (size_t)kind * w.cap;
            uint32_t* qs = w.q_mat_slot + (size_t)kind * w.cap;
            switch (kind) {
                case YK_MAT_MATTE:
                    if (is_path) k_shade<YK_MAT_MATTE, true><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    else k_shade<YK_MAT_MATTE, false><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    break;
                case YK_MAT_GLASS:
                    if (is_path) k_shade<YK_MAT_GLASS, true><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    else k_shade<YK_MAT_GLASS, false><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    break;
                case YK_MAT_METAL:
                    if (is_path) k_shade<YK_MAT_METAL, true><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    else k_shade<YK_MAT_METAL, false><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    break;
                default:
                    if (is_path) k_shade<YK_MAT_GLOSSY, true><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    else k_shade<YK_MAT_GLOSSY, false><<<shade_blocks, kShadeThreads, 0, s>>>(sc->dev, w, cfg, bt, q, qt, qs, b, cur, nxt, q_next);
                    break;
            }
            tm->launches += 1;
        }
        if (c->stage_timing > 1) CUDA_TRY(cudaEventRecord(stage_event(iter, 3), s));
        if (spheres) k_trace_shadow<true><<<shadow_blocks, kTraceThreads, 0, s>>>(sc->dev, w, cfg, cur);
        else k_trace_shadow<false><<<shadow_blocks, kTraceThreads, 0, s>>>(sc->dev, w, cfg, cur);
        if (c->stage_timing > 1) CUDA_TRY(cudaEventRecord(stage_event(iter, 4), s));
        tm->launches += 1;
        sl.n_iters = iter + 1;
        q_cur = q_next;
        flip ^= 1;
        if (sync_loop) {
            CUDA_TRY(cudaMemcpyAsync(p->h_ctr, nxt, sizeof(IterCounters), cudaMemcpyDeviceToHost, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            if (p->h_ctr->n_active == 0) break;
            if (iter > (1u << 20)) return yk_set_error(YK_ERR_INVALID, "yk_render: bounce loop did not terminate");
        }
    }
    if (accumulate_film) {
        k_film_add<<<(bt.n_paths + T - 1) / T, T, 0, s>>>(w, bt, d_film, cfg.res_x);
    } else {
        k_film_accumulate<<<(bt.n_jobs + T - 1) / T, T, 0, s>>>(w, bt, c->d_accum, cfg.res_x);
    }
    tm->launches += 1;
    CUDA_TRY(cudaEventRecord(sl.done, s));
    CUDA_TRY(cudaGetLastError());
    sl.busy = true;
    p->n_batches += 1;
    return YK_OK;
}

}  // namespace

extern "C" {

int yk_context_create(int device_id, yk_context** out) {
    if (!out) return yk_set_error(YK_ERR_INVALID, "yk_context_create: null output");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return yk_set_error(YK_ERR_CUDA, std::string("yk_context_create: no CUDA device (") + cudaGetErrorString(e) +
                                             "); this backend has no CPU fallback");
    if (device_id < 0 || device_id >= n_dev) return yk_set_error(YK_ERR_INVALID, "yk_context_create: device id out of range");
    CUDA_TRY(cudaSetDevice(device_id));
    auto* c = new yk_context();
    c->device = device_id;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device_id));
    c->sm_count = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& ev : c->ev) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaEventDestroy(c->ev[2]));
    CUDA_TRY(cudaEventDestroy(c->ev[3]));
    CUDA_TRY(cudaEventCreate(&c->ev[2]));  // timing pair around the whole render
    CUDA_TRY(cudaEventCreate(&c->ev[3]));
    if (const char* np = getenv("YK_PIPES")) c->n_pipes_env = std::max(1, std::min(kMaxPipes, atoi(np)));
    if (const char* st = getenv("YK_STAGE_TIMING")) c->stage_timing = std::max(0, std::min(2, atoi(st)));
    for (int i = 0; i < kMaxPipes; ++i) {
        Pipe& p = c->pipe[i];
        if (i == 0) p.stream = c->stream;
        else {
            CUDA_TRY(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking));
            p.owns_stream = true;
        }
        CUDA_TRY(cudaMalloc((void**)&p.d_ctr, 2 * sizeof(IterCounters)));
        CUDA_TRY(cudaMalloc((void**)&p.d_jumps, kMaxBatchSamples * sizeof(SampleJump)));
        CUDA_TRY(cudaMallocHost((void**)&p.h_ctr, sizeof(IterCounters)));
        CUDA_TRY(cudaMallocHost((void**)&p.h_totals, sizeof(Totals)));
    }
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_trace_closest, k_trace_closest<false, false>, kTraceThreads, 0));
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c->occ_trace_any, k_trace_shadow<false>, kTraceThreads, 0));
    *out = c;
    return YK_OK;
}

void yk_context_destroy(yk_context* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (Pipe& p : c->pipe) {
        free_bag(p.wave_allocs);
        cudaFree(p.d_ctr);
        cudaFree(p.d_dim_hash);
        cudaFree(p.d_jobs);
        cudaFree(p.d_jumps);
        cudaFreeHost(p.h_ctr);
        cudaFreeHost(p.h_totals);
        for (auto& sl : p.slot) {
            if (sl.done) cudaEventDestroy(sl.done);
            for (cudaEvent_t e : sl.ev) cudaEventDestroy(e);
        }
        if (p.owns_stream) cudaStreamDestroy(p.stream);
    }
    cudaFree(c->d_tiles);
    cudaFree(c->d_tile_off);
    cudaFree(c->d_accum);
    cudaFree(c->d_film);
    cudaFree(c->d_hit_ids);
    for (auto& ev : c->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(c->stream);
    delete c;
}

void* yk_context_stream(yk_context* c) { return c ? (void*)c->stream : nullptr; }

int yk_scene_create(yk_context* c, const yk_scene_desc* d, yk_scene** out) {
    if (!c || !d || !out) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: null argument");
    if (!d->n_nodes || !d->nodes || !d->n_tris || !d->tri_vertices || !d->tri_orig_id || !d->tri_material || !d->tri_area_light ||
        !d->tri_flags)
        return yk_set_error(YK_ERR_INVALID, "yk_scene_create: missing node / triangle arrays");
    if (d->n_lights > (uint32_t)kMaxLights) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: more than 16 lights");
    if (d->n_materials > 0xffffffu) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: too many materials");
    CUDA_TRY(cudaSetDevice(c->device));
    auto sc = std::make_unique<yk_scene>();
    sc->ctx = c;
    sc->device = c->device;
    int rc;
    // The reference-layout arrays go to the device as they are and are repacked there (k_scene_*): no host-side copy of
    // the scene is built. Meanwhile host threads validate the same arrays (indices, ranges, flags).
    struct Check {
        const char* error = nullptr;
        uint32_t n_interior = 0, kinds = 0;
        bool small_leaves = true;
    };
    const unsigned n_workers = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    auto validate = [d, n_workers](unsigned wi) -> Check {
        Check ck;
        auto fail = [&ck](const char* m) { if (!ck.error) ck.error = m; };
        const uint64_t n0 = (uint64_t)d->n_nodes * wi / n_workers, n1 = (uint64_t)d->n_nodes * (wi + 1) / n_workers;
        for (uint64_t i = n0; i < n1; ++i) {
            const yk_bvh_node& n = d->nodes[i];
            if (!n.is_leaf) {
                ck.n_interior += 1;
                if (n.split_axis > 2) fail("yk_scene_create: split axis out of range");
                // children follow their parent (pre-order, bvh.rs:396-419): the walk cannot cycle
                if (n.offset >= d->n_nodes || i + 1 >= d->n_nodes || n.offset <= i + 1) fail("yk_scene_create: child index out of range");
            } else {
                if ((uint64_t)n.offset + n.shape_count > d->n_tris) fail("yk_scene_create: leaf range out of range");
                if (n.shape_count < 1 || n.shape_count > 16) ck.small_leaves = false;
            }
        }
        const uint64_t t0 = (uint64_t)d->n_tris * wi / n_workers, t1 = (uint64_t)d->n_tris * (wi + 1) / n_workers;
        for (uint64_t i = t0; i < t1; ++i) {
            const uint32_t m = d->tri_material[i];
            if (m >= d->n_materials) { fail("yk_scene_create: material index out of range"); continue; }
            if (d->materials[m].kind <= YK_MAT_GLOSSY) ck.kinds |= 1u << d->materials[m].kind;
            const int32_t al = d->tri_area_light[i];
            if (al >= (int32_t)d->n_lights) fail("yk_scene_create: area light out of range");
            else if (al >= 0 && d->lights[al].kind != YK_LIGHT_RECT) fail("yk_scene_create: area light must be rectangular");
            const uint8_t f = d->tri_flags[i];
            if ((f & YK_TRI_HAS_NORMALS) && !d->tri_normals) fail("yk_scene_create: normals flagged but absent");
            if ((f & YK_TRI_HAS_UVS) && !d->tri_uvs) fail("yk_scene_create: uvs flagged but absent");
            if ((f & YK_TRI_IS_SPHERE) && (!d->tri_sphere || !d->spheres || d->tri_sphere[i] < 0 || (uint32_t)d->tri_sphere[i] >= d->n_spheres))
                fail("yk_scene_create: sphere slot without a valid sphere index");
        }
        return ck;
    };
    for (uint32_t i = 0; i < d->n_materials; ++i)
        if (d->materials[i].kind > YK_MAT_GLOSSY) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: unknown material kind");
    for (uint32_t i = 0; i < d->n_lights; ++i)
        if (d->lights[i].kind > YK_LIGHT_DISTANT) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: unknown light kind");
    std::vector<std::future<Check>> checks;
    for (unsigned wi = 0; wi < n_workers; ++wi) checks.push_back(std::async(std::launch::async, validate, wi));
    auto join_checks = [&](Check* total) {
        for (auto& f : checks) {
            const Check ck = f.get();
            if (ck.error && !total->error) total->error = ck.error;
            total->n_interior += ck.n_interior;
            total->kinds |= ck.kinds;
            total->small_leaves = total->small_leaves && ck.small_leaves;
        }
        checks.clear();
    };
    struct Cleanup {  // on an error return: wait for the validators, release what was uploaded
        std::function<void()> fn;
        bool armed = true;
        ~Cleanup() { if (armed) fn(); }
    };
    std::vector<void*> temps;
    Cleanup cleanup{[&] {
        Check ignore;
        join_checks(&ignore);
        cudaDeviceSynchronize();
        free_bag(temps);
        free_bag(sc->allocs);
    }};
    const yk_bvh_node* r_nodes = nullptr;
    const float* r_verts = nullptr;
    const uint32_t *r_orig = nullptr, *r_mat = nullptr;
    const int32_t *r_alight = nullptr, *r_sphere = nullptr;
    const uint8_t *r_flags = nullptr, *r_kinds = nullptr;
    std::vector<uint8_t> kinds(std::max(d->n_materials, 1u), 0);
    for (uint32_t i = 0; i < d->n_materials; ++i) kinds[i] = (uint8_t)d->materials[i].kind;
    if ((rc = dev_upload(temps, &r_nodes, d->nodes, d->n_nodes)) != YK_OK) return rc;
    if ((rc = dev_upload(temps, &r_verts, d->tri_vertices, (size_t)d->n_tris * 9)) != YK_OK) return rc;
    if ((rc = dev_upload(temps, &r_orig, d->tri_orig_id, d->n_tris)) != YK_OK) return rc;
    if ((rc = dev_upload(temps, &r_mat, d->tri_material, d->n_tris)) != YK_OK) return rc;
    if ((rc = dev_upload(temps, &r_alight, d->tri_area_light, d->n_tris)) != YK_OK) return rc;
    if ((rc = dev_upload(temps, &r_flags, d->tri_flags, d->n_tris)) != YK_OK) return rc;
    if (d->tri_sphere && (rc = dev_upload(temps, &r_sphere, d->tri_sphere, d->n_tris)) != YK_OK) return rc;
    if ((rc = dev_upload(temps, &r_kinds, kinds.data(), kinds.size())) != YK_OK) return rc;
    if (d->tri_normals && (rc = dev_upload(sc->allocs, &sc->dev.normals, d->tri_normals, (size_t)d->n_tris * 9)) != YK_OK) return rc;
    if (d->tri_uvs && (rc = dev_upload(sc->allocs, &sc->dev.uvs, d->tri_uvs, (size_t)d->n_tris * 6)) != YK_OK) return rc;
    Check total;
    join_checks(&total);
    if (total.error) return yk_set_error(YK_ERR_INVALID, total.error);
    if (total.n_interior > kRefIndexMask) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: more than 2^29 interior nodes");
    sc->material_kinds = total.kinds;
    const uint32_t n_interior = total.n_interior, n_leaves = d->n_nodes - n_interior;
    const bool packed_leaves = total.small_leaves && d->n_tris <= (1u << kLeafFirstBits);

    // Nodes: one 64-byte record per interior node with the boxes of both children (DevScene::nodes2). Interior records are
    // numbered in the reference's pre-order (an exclusive scan of the interior flags), so a first child's record follows
    // its parent's; leaves are numbered the same way for the leaf table.
    {
        cudaStream_t st = c->stream;
        const uint32_t n = d->n_nodes;
        uint32_t *d_flag = nullptr, *d_rec = nullptr;
        float4* d_rec_out = nullptr;
        uint2* d_leaf = nullptr;
        if ((rc = dev_alloc(temps, &d_flag, n)) != YK_OK || (rc = dev_alloc(temps, &d_rec, n)) != YK_OK) return rc;
        if ((rc = dev_alloc(sc->allocs, &d_rec_out, (size_t)std::max(n_interior, 1u) * 4)) != YK_OK) return rc;
        if (!packed_leaves && (rc = dev_alloc(sc->allocs, &d_leaf, n_leaves)) != YK_OK) return rc;
        k_scene_interior_flags<<<(n + 255) / 256, 256, 0, st>>>(r_nodes, n, d_flag);
        size_t scan_bytes = 0;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, d_flag, d_rec, (int)n, st));
        unsigned char* d_scan = nullptr;
        if ((rc = dev_alloc(temps, &d_scan, scan_bytes)) != YK_OK) return rc;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(d_scan, scan_bytes, d_flag, d_rec, (int)n, st));
        k_scene_records<<<(n + 255) / 256, 256, 0, st>>>(r_nodes, d_rec, n, packed_leaves ? 1 : 0, d_leaf, d_rec_out);
        sc->dev.nodes2 = d_rec_out;
        sc->dev.leaf_table = d_leaf;
        const yk_bvh_node& root = d->nodes[0];
        if (!root.is_leaf) sc->dev.root_ref = kRefInterior | ((uint32_t)root.split_axis << 29);  // record 0
        else if (packed_leaves) sc->dev.root_ref = ((uint32_t)(root.shape_count - 1) << kLeafFirstBits) | root.offset;
        else {
            sc->dev.root_ref = 0;  // leaf 0 of the table; nobody's child, so it is written here
            const uint2 entry = make_uint2(root.offset, root.shape_count);
            CUDA_TRY(cudaMemcpyAsync(d_leaf, &entry, sizeof entry, cudaMemcpyHostToDevice, st));
        }
        std::memcpy(sc->dev.root_min, root.p_min, 12);
        std::memcpy(sc->dev.root_max, root.p_max, 12);
        // Triangles: three 16-byte words, vertices pre-gathered in leaf order and transposed (x0 x1 x2 | y0 y1 y2 | z0 z1 z2);
        // the w lanes carry the per-triangle ids.
        float4* d_tris = nullptr;
        if ((rc = dev_alloc(sc->allocs, &d_tris, (size_t)d->n_tris * 3)) != YK_OK) return rc;
        k_scene_tris<<<(d->n_tris + 255) / 256, 256, 0, st>>>(r_verts, r_orig, r_mat, r_alight, r_flags, r_sphere, r_kinds, d->n_tris, d_tris);
        sc->dev.tris = d_tris;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(st));
        free_bag(temps);
    }

    std::vector<DevTexture> tex(d->n_textures);
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const yk_texture_desc& t = d->textures[i];
        DevTexture dt{};
        dt.kind = t.kind;
        dt.width = t.width;
        dt.height = t.height;
        std::memcpy(dt.value, t.value, 12);
        if (t.kind == YK_TEX_IMAGE) {
            if (!t.texels || !t.width || !t.height) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: empty image texture");
            if ((rc = dev_upload(sc->allocs, &dt.texels, t.texels, (size_t)t.width * t.height * 3)) != YK_OK) return rc;
        } else if (t.kind != YK_TEX_CONSTANT) {
            return yk_set_error(YK_ERR_INVALID, "yk_scene_create: unknown texture kind");
        }
        tex[i] = dt;
    }
    std::vector<DevMaterial> mats(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const yk_material_desc& m = d->materials[i];
        if (m.kind > YK_MAT_GLOSSY) return yk_set_error(YK_ERR_INVALID, "yk_scene_create: unknown material kind");
        const int n_tex = m.kind == YK_MAT_METAL ? 3 : 2;
        for (int k = 0; k < n_tex; ++k)
            if (m.tex[k] < 0 || (uint32_t)m.tex[k] >= d->n_textures)
                return yk_set_error(YK_ERR_INVALID, "yk_scene_create: texture index out of range");
        DevMaterial dm{};
        dm.kind = m.kind;
        std::memcpy(dm.tex, m.tex, 12);
        dm.eta = m.eta;
        dm.remap = m.remap_roughness;
        dm.const_alpha = -1.0f;
        if (m.kind == YK_MAT_METAL || m.kind == YK_MAT_GLOSSY) {
            const yk_texture_desc& rt = d->textures[m.tex[m.kind == YK_MAT_METAL ? 2 : 1]];
            if (rt.kind == YK_TEX_CONSTANT) {  // metal.rs:41-45 / glossy.rs:39-49 + trowbridge_reitz.rs:16-20
                float r = rt.value[0];
                if (m.remap_roughness) r = host_roughness_to_alpha(r);
                dm.const_alpha = fmaxf(m.kind == YK_MAT_GLOSSY ? r * r : r, 0.001f);
            }
        }
        mats[i] = dm;
    }
    if ((rc = dev_upload(sc->allocs, &sc->dev.textures, tex.data(), tex.size())) != YK_OK) return rc;
    if ((rc = dev_upload(sc->allocs, &sc->dev.materials, mats.data(), mats.size())) != YK_OK) return rc;
    if ((rc = dev_upload(sc->allocs, &sc->dev.lights, d->lights, d->n_lights)) != YK_OK) return rc;
    if (d->n_spheres && (rc = dev_upload(sc->allocs, &sc->dev.spheres, d->spheres, d->n_spheres)) != YK_OK) return rc;
    sc->dev.n_lights = d->n_lights;
    sc->dev.n_tris = d->n_tris;
    sc->dev.n_nodes = d->n_nodes;
    std::memcpy(sc->dev.background, d->background, 12);
    cleanup.armed = false;
    *out = sc.release();
    return YK_OK;
}

void yk_scene_destroy(yk_scene* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    free_bag(s->allocs);
    delete s;
}

int yk_render(yk_context* c, const yk_scene* sc, const yk_camera* cam, const yk_film_settings* fs, const yk_sampler* sm,
              const yk_integrator* in, const yk_tile* tiles, uint32_t n_tiles, const yk_render_opts* opts, float* film_rgb,
              yk_stats* stats) {
    const auto wall0 = std::chrono::steady_clock::now();
    if (!c || !sc || !cam || !fs || !sm || !in || !film_rgb) return yk_set_error(YK_ERR_INVALID, "yk_render: null argument");
    if (sc->ctx != c) return yk_set_error(YK_ERR_INVALID, "yk_render: scene belongs to another context");
    if (!fs->res_x || !fs->res_y || fs->res_x > 0xffffu || fs->res_y > 0xffffu)
        return yk_set_error(YK_ERR_INVALID, "yk_render: film resolution must fit u16 (integrators/mod.rs:140-141)");
    if (in->kind > YK_INTEGRATOR_SHADING_UVS) return yk_set_error(YK_ERR_INVALID, "yk_render: unknown integrator");
    if (sm->kind > YK_SAMPLER_STRATIFIED) return yk_set_error(YK_ERR_INVALID, "yk_render: unknown sampler");
    const uint32_t spp = sm->kind == YK_SAMPLER_UNIFORM ? sm->nx : sm->nx * sm->ny;
    if (spp == 0 || spp > 0x10000u) return yk_set_error(YK_ERR_INVALID, "yk_render: samples per pixel must be in 1..65536");
    if (in->max_depth > 255) return yk_set_error(YK_ERR_INVALID, "yk_render: max_depth above 255");
    if (in->kind == YK_INTEGRATOR_WHITTED && in->max_depth > 24)
        return yk_set_error(YK_ERR_INVALID, "yk_render: whitted max_depth above 24 is not supported");
    if (n_tiles && !tiles) return yk_set_error(YK_ERR_INVALID, "yk_render: null tile list");
    CUDA_TRY(cudaSetDevice(c->device));
    (void)cudaGetLastError();  // do not inherit a stale error from an unrelated earlier call
    cudaStream_t s = c->stream;
    const uint32_t flags = opts ? opts->flags : 0u;
    const bool on_device = (flags & YK_RENDER_FILM_ON_DEVICE) != 0;
    const bool accumulate = fs->accumulate != 0;

    // Pixel jobs: tile order, row-major inside a tile. Only the tile list and the prefix sum of the tile areas go to the
    // device; each pixel group's jobs are expanded there (k_jobs_expand).
    std::vector<unsigned long long> tile_off((size_t)n_tiles + 1, 0ull);
    for (uint32_t t = 0; t < n_tiles; ++t) {
        const yk_tile& tl = tiles[t];
        if (tl.x0 >= tl.x1 || tl.y0 >= tl.y1 || tl.x1 > fs->res_x || tl.y1 > fs->res_y)
            return yk_set_error(YK_ERR_INVALID, "yk_render: tile outside the film (film.rs:224-231)");
        tile_off[t + 1] = tile_off[t] + (unsigned long long)(tl.x1 - tl.x0) * (tl.y1 - tl.y0);
    }
    const unsigned long long n_jobs_total = tile_off[n_tiles], area = n_jobs_total;
    const size_t n_pixels = (size_t)fs->res_x * fs->res_y;
    const uint32_t samples_per_job = accumulate ? 1u : spp;

    RenderCfg cfg{};
    cfg.sampler = SamplerCfg{};
    cfg.sampler.kind = sm->kind;
    cfg.sampler.nx = sm->nx;
    cfg.sampler.ny = sm->kind == YK_SAMPLER_UNIFORM ? 1u : sm->ny;
    cfg.sampler.jitter = sm->jitter;
    cfg.sampler.seed = sm->seed;
    cfg.sampler.div_nx = FastDiv::make(cfg.sampler.nx);
    cfg.sampler.div_ny = FastDiv::make(cfg.sampler.ny);
    cfg.sampler.div_n = FastDiv::make(cfg.sampler.nx * cfg.sampler.ny);
    cfg.integrator = in->kind;
    cfg.max_depth = in->max_depth;
    cfg.has_clamp = in->has_clamp;
    cfg.clamp = in->indirect_clamp;
    std::memcpy(cfg.c2w, cam->camera_to_world, 64);
    std::memcpy(cfg.r2c, cam->raster_to_camera, 64);
    cfg.res_x = fs->res_x;
    cfg.res_y = fs->res_y;
    cfg.aux_sample = opts ? opts->aux_sample : 0u;

    // Device film / accumulators.
    if (c->film_cap < n_pixels) {
        cudaFree(c->d_accum); cudaFree(c->d_film); cudaFree(c->d_hit_ids);
        c->d_accum = nullptr; c->d_film = nullptr; c->d_hit_ids = nullptr;
        c->film_cap = 0;
        CUDA_TRY(cudaMalloc((void**)&c->d_accum, n_pixels * 3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&c->d_film, n_pixels * 3 * sizeof(float)));
        CUDA_TRY(cudaMalloc((void**)&c->d_hit_ids, n_pixels * sizeof(int32_t)));
        c->film_cap = n_pixels;
    }
    float* d_film = on_device ? film_rgb : c->d_film;
    const bool full_cover = !accumulate && area == n_pixels;
    if (!on_device) {
        if (full_cover) CUDA_TRY(cudaMemsetAsync(d_film, 0, n_pixels * 3 * sizeof(float), s));
        else CUDA_TRY(cudaMemcpyAsync(d_film, film_rgb, n_pixels * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    int32_t* d_ids = nullptr;
    if (opts && opts->hit_ids) {
        d_ids = on_device ? opts->hit_ids : c->d_hit_ids;
        k_fill_i32<<<(unsigned)((n_pixels + 255) / 256), 256, 0, s>>>(d_ids, n_pixels, -1);
    }
    cfg.hit_ids = d_ids;

    yk_stats st{};
    Timers tm;
    Totals totals{};
    CUDA_TRY(cudaEventRecord(c->ev[2], s));
    if (n_jobs_total) {
        if (c->tiles_cap < n_tiles) {
            cudaFree(c->d_tiles);
            cudaFree(c->d_tile_off);
            c->d_tiles = nullptr;
            c->d_tile_off = nullptr;
            c->tiles_cap = 0;
            CUDA_TRY(cudaMalloc((void**)&c->d_tiles, (size_t)n_tiles * sizeof(yk_tile)));
            CUDA_TRY(cudaMalloc((void**)&c->d_tile_off, ((size_t)n_tiles + 1) * sizeof(unsigned long long)));
            c->tiles_cap = n_tiles;
        }
        CUDA_TRY(cudaMemcpyAsync(c->d_tiles, tiles, (size_t)n_tiles * sizeof(yk_tile), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(c->d_tile_off, tile_off.data(), tile_off.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
        // Wavefront capacity: paths in flight per batch.
        // Default: as many as ~6 GB of wavefront state per pipe hold, at most 16 Mi (measured on the Cornell bench: 4 Mi -> 8 Mi ->
        // 16 Mi -> 32 Mi paths = +8 %, +12 %, +14 %: fewer, longer launches amortise the kernels' tails and the launch gaps),
        // and never more than a quarter of the free device memory.
        const uint32_t stack_entries = in->kind == YK_INTEGRATOR_WHITTED ? std::max(in->max_depth, 1u) : 0u;
        uint32_t cap = opts ? opts->wavefront_paths : 0u;
        if (!cap) {
            const uint64_t bytes_per_path = 320ull + 40ull * std::max(sc->dev.n_lights, 1u) + 52ull * stack_entries;
            if (!c->mem_budget) {  // asked once per context: cudaMemGetInfo can take milliseconds
                size_t free_b = 0, total_b = 0;
                c->mem_budget = 6ull << 30;
                if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) c->mem_budget = std::min<uint64_t>(c->mem_budget, (uint64_t)free_b / 8);
            }
            cap = (uint32_t)std::min<uint64_t>(1u << 24, std::max<uint64_t>(1u << 20, c->mem_budget / bytes_per_path));
        }
        const uint64_t total_paths = (uint64_t)n_jobs_total * samples_per_job;
        if (cap > total_paths) cap = (uint32_t)total_paths;
        cap = std::max(cap, 32u);
        // Samples of one pixel per batch: enough to amortise per-batch fixed costs, few enough that many pixels
        // (>= 64 Ki when available) share a batch.
        uint32_t m = std::min(samples_per_job, kMaxBatchSamples);
        while (m > 1 && (uint64_t)m * std::min<uint64_t>(n_jobs_total, 65536) > cap) m >>= 1;
        uint32_t jobs_per_batch = std::max(1u, cap / m);
        // Pipes: pixel groups alternate between the streams, so one group's latency-bound shading overlaps the other's
        // issue-bound traversal. A pixel's samples stay on one pipe, in order (the film sum is order-dependent).
        // The accumulating film adds with atomics across tiles, so it keeps to one stream.
        // Default: a second pipe only where a bounce is many small launches (several material kinds / lights), measured
        // +17 % on the config-4 room and +5 % on the Cornell box (profiles/r01); YK_PIPES / opts->pipes override.
        int n_pipes = c->n_pipes_env > 0 ? c->n_pipes_env
                                         : ((__builtin_popcount(sc->material_kinds) + (int)sc->dev.n_lights >= 5) ? 2 : 1);
        if (opts && opts->pipes) n_pipes = (int)opts->pipes;
        n_pipes = std::max(1, std::min(kMaxPipes, n_pipes));
        if (accumulate) n_pipes = 1;
        if (n_pipes > 1 && n_jobs_total <= jobs_per_batch) {  // one group only: split it if that leaves decent batches
            if ((uint64_t)(n_jobs_total / 2) * m >= (1u << 20)) jobs_per_batch = (uint32_t)((n_jobs_total + 1) / 2);
            else n_pipes = 1;
        }
        const uint32_t wave_cap = (uint32_t)std::min<uint64_t>(cap, (uint64_t)jobs_per_batch * m);
        int rc = YK_OK;
        CUDA_TRY(cudaEventRecord(c->ev[0], s));
        for (int pi = 0; pi < n_pipes; ++pi) {
            Pipe& p = c->pipe[pi];
            if ((rc = ensure_wave(&p, wave_cap, sc->dev.n_lights, stack_entries)) != YK_OK) return rc;
            if (pi > 0) CUDA_TRY(cudaStreamWaitEvent(p.stream, c->ev[0], 0));
            CUDA_TRY(cudaMemsetAsync(p.wave.totals, 0, sizeof(Totals), p.stream));
            p.hash_group = (size_t)-1;
        }
        // Pixel groups: consecutive runs of at most jobs_per_batch jobs. The accumulating film adds every batch into the
        // film in stream order, so a group never spans a change of `tile.sample`: within one sample index the reference's
        // tiles are disjoint, each pixel appears once per batch and the per-pixel sum runs in tile-list order exactly
        // like the reference's `*fc += c` (film.rs:260-272).
        struct Group { unsigned long long first; uint32_t count, t_lo, t_hi; };  // jobs [first, first + count) lie in tiles [t_lo, t_hi)
        std::vector<Group> groups;
        {
            uint32_t seg_t0 = 0;
            auto flush = [&](uint32_t seg_t1) {  // the jobs of tiles [seg_t0, seg_t1)
                uint32_t t = seg_t0;
                for (unsigned long long j = tile_off[seg_t0]; j < tile_off[seg_t1]; j += jobs_per_batch) {
                    const unsigned long long end = std::min<unsigned long long>(j + jobs_per_batch, tile_off[seg_t1]);
                    while (tile_off[t + 1] <= j) ++t;
                    uint32_t t_hi = t + 1;
                    while (tile_off[t_hi] < end) ++t_hi;
                    groups.push_back(Group{j, (uint32_t)(end - j), t, t_hi});
                }
                seg_t0 = seg_t1;
            };
            if (accumulate)
                for (uint32_t t = 1; t < n_tiles; ++t)
                    if (tiles[t].sample != tiles[t - 1].sample) flush(t);
            flush(n_tiles);
        }
        const size_t n_groups = groups.size();
        uint64_t done = 0;
        bool cancelled = false;
        for (size_t g0 = 0; g0 < n_groups && !cancelled; g0 += n_pipes) {
            for (uint32_t s0 = 0; s0 < samples_per_job && !cancelled; s0 += m) {
                for (int pi = 0; pi < n_pipes && g0 + pi < n_groups && !cancelled; ++pi) {
                    Pipe& p = c->pipe[pi];
                    const size_t group = g0 + pi;
                    const uint32_t nj = groups[group].count;
                    if (p.hash_group != group) {  // a new pixel group on this pipe: its jobs, and a zeroed accumulator
                        if (p.jobs_cap < nj) {
                            CUDA_TRY(cudaStreamSynchronize(p.stream));
                            cudaFree(p.d_jobs);
                            p.d_jobs = nullptr;
                            p.jobs_cap = 0;
                            CUDA_TRY(cudaMalloc((void**)&p.d_jobs, (size_t)std::max(nj, jobs_per_batch) * sizeof(Job)));
                            p.jobs_cap = std::max(nj, jobs_per_batch);
                        }
                        k_jobs_expand<<<(nj + 255) / 256, 256, 0, p.stream>>>(c->d_tiles, c->d_tile_off, groups[group].t_lo, groups[group].t_hi,
                                                                               groups[group].first, nj, accumulate ? 1u : 0u, p.d_jobs);
                        tm.launches += 1;
                        if (!accumulate) {
                            k_zero_jobs<<<(nj + 255) / 256, 256, 0, p.stream>>>(p.d_jobs, nj, c->d_accum, fs->res_x);
                            tm.launches += 1;
                        }
                    }
                    RenderCfg gcfg = cfg;
                    if (sm->kind == YK_SAMPLER_STRATIFIED) {
                        // tabulate the (pixel, dimension) hashes of this pixel group once for all of its samples
                        uint64_t dims = 2;
                        if (in->kind == YK_INTEGRATOR_PATH) dims = 2 + (uint64_t)in->max_depth * (2ull * sc->dev.n_lights + 3);
                        else if (in->kind == YK_INTEGRATOR_WHITTED)
                            dims = 2 + 2ull * sc->dev.n_lights * ((1ull << std::min(in->max_depth, 8u)) - 1);
                        dims = std::min<uint64_t>(dims, std::min<uint64_t>(4096, (256ull << 20) / (4ull * nj)));  // the rest is hashed on the fly
                        if (p.hash_group != group) {
                            const size_t need = (size_t)dims * nj;
                            if (p.dim_hash_cap < need) {
                                CUDA_TRY(cudaStreamSynchronize(p.stream));
                                cudaFree(p.d_dim_hash);
                                p.d_dim_hash = nullptr;
                                p.dim_hash_cap = 0;
                                CUDA_TRY(cudaMalloc((void**)&p.d_dim_hash, need * sizeof(uint32_t)));
                                p.dim_hash_cap = need;
                            }
                            k_dim_hashes<<<dim3((nj + 255) / 256, (unsigned)std::min<uint64_t>(dims, 64)), 256, 0, p.stream>>>(
                                p.d_jobs, nj, (uint32_t)dims, sm->seed, p.d_dim_hash);
                            tm.launches += 1;
                        }
                        gcfg.sampler.hash_table = p.d_dim_hash;
                        gcfg.sampler.n_hash_dims = (uint32_t)dims;
                        gcfg.sampler.hash_stride = nj;
                    }
                    p.hash_group = group;
                    Batch bt;
                    bt.jobs = p.d_jobs;
                    bt.n_jobs = nj;
                    bt.div_jobs = FastDiv::make(nj);
                    bt.sample_off = s0;
                    bt.n_samples = std::min(m, samples_per_job - s0);
                    bt.n_paths = nj * bt.n_samples;
                    // the batch's first sample index: s0, or the sample of the group's tiles (accumulating films: one per group)
                    const uint32_t first_sample = accumulate ? (uint32_t)tiles[groups[group].t_lo].sample : s0;
                    rc = run_batch(c, &p, sc, gcfg, bt, first_sample, accumulate, d_film, &tm, &done);
                    if (rc != YK_OK) { cudaDeviceSynchronize(); return rc; }
                    if (!accumulate && s0 + m >= samples_per_job) {
                        // the group's last samples are queued: `color /= sample_count` + Film::update_tile for its pixels
                        k_film_store<<<(nj + 255) / 256, 256, 0, p.stream>>>(p.d_jobs, nj, c->d_accum, d_film, fs->res_x, (float)spp);
                        tm.launches += 1;
                    }
                    if (opts && opts->progress && opts->progress(opts->progress_user, done, total_paths)) cancelled = true;
                }
            }
        }
        for (int pi = 0; pi < n_pipes; ++pi) {
            Pipe& p = c->pipe[pi];
            CUDA_TRY(cudaMemcpyAsync(p.h_totals, p.wave.totals, sizeof(Totals), cudaMemcpyDeviceToHost, p.stream));
            for (int k = 0; k < kRing; ++k)
                if ((rc = retire_slot(&p, k, &tm, &done)) != YK_OK) return rc;
            CUDA_TRY(cudaStreamSynchronize(p.stream));
            const Totals& t = *p.h_totals;
            totals.closest_nodes += t.closest_nodes; totals.closest_tris += t.closest_tris;
            totals.any_nodes += t.any_nodes; totals.any_tris += t.any_tris;
            totals.hit_hash += t.hit_hash; totals.shadow_rays += t.shadow_rays; totals.closest_rays += t.closest_rays;
        }
        if (cancelled) return yk_set_error(YK_ERR_CANCELLED, "yk_render: cancelled by the progress callback");
        if (opts && opts->progress) opts->progress(opts->progress_user, done, total_paths);
        st.samples = total_paths;
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], s));
    if (!on_device) {
        CUDA_TRY(cudaMemcpyAsync(film_rgb, d_film, n_pixels * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (opts && opts->hit_ids)
            CUDA_TRY(cudaMemcpyAsync(opts->hit_ids, d_ids, n_pixels * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaGetLastError());
    if (stats) {
        st.closest_nodes = totals.closest_nodes;
        st.closest_tris = totals.closest_tris;
        st.any_nodes = totals.any_nodes;
        st.any_tris = totals.any_tris;
        st.primary_hit_hash = totals.hit_hash;
        st.shadow_rays = totals.shadow_rays;
        st.ray_count = totals.closest_rays;
        float ms = 0;
        cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]);
        st.device_ms = ms;
        st.trace_closest_ms = tm.closest;
        st.trace_any_ms = tm.any;
        st.shade_ms = tm.shade;
        st.kernel_launches = tm.launches;
        st.trace_closest_launches = tm.closest_launches;
        st.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        *stats = st;
    }
    return YK_OK;
}

}  // extern "C"
