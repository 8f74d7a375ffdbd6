// BVH traversal kernels: closest hit (bvh.rs:160-232) and the shadow-ray + radiance-fold kernel (bvh.rs:235-302).
// Part of the single translation unit render.cu (compiled --fmad=false: every float op is the reference's un-fused IEEE op).
#pragma once
#include "wf_common.cuh"

namespace {

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
constexpr uint32_t kKeyNever = 0xffffffffu;  // stack key of a child whose slab entry lies beyond its exit
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
__device__ __forceinline__ void lds_entry(uint32_t addr, uint32_t* ref, float* key) {
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(*ref), "=f"(*key) : "r"(addr));
}
__device__ __forceinline__ void sts_entry(uint32_t addr, uint32_t ref, float key) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(ref), "f"(key) : "memory");
}

#ifndef YK_LD256
#define YK_LD256 1  // measured: closest-hit time -0.5 % (Cornell) ... -3.4 % (10 M-triangle terrain) against four LDG.128
#endif
// 256-bit read-only load (sm_100+, LDG.E.256): two consecutive float4 of a 32-byte aligned address in one instruction.
__device__ __forceinline__ void ldg256(const float4* p, float4* a, float4* b) {
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w)
        : "l"(p));
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
//    per-visit walk). A far child that can never pass (entry beyond its own exit) gets the key kKeyNever.
//  * counters: the closest-hit walk always drains its stack, so both tests of a record are counted when it is loaded
//    and never-passing far children are not pushed; the any-hit walk ends early, so it counts a far child's test when
//    it is popped (or tested on the spot) and pushes kKeyNever entries too.
struct TraceLane {
    float ox, oy, oz, ix, iy, iz, t_max;
    float okx, oky, okz, sx, sy, sz;  // watertight test: permuted origin, shear
    uint32_t kx, ky, kz, neg_mask;
    uint32_t cur;  // interior ref to enter next, or kNoNode
    uint32_t leaf_pos, leaf_end;
    uint32_t sp;  // shared-memory byte address of the lane's next free stack entry (level 0 holds the sentinel)
    uint32_t n_tests, n_tris;  // running totals of the lane (all of its rays): box tests, shape tests
    uint32_t n_hits;           // passed box tests of the current ray (COUNTS only)
    uint32_t pend_ref;         // POST only: a second leaf waiting behind the current one (kNoNode = none) and its box key
    float pend_key;

    __device__ __forceinline__ void idle(uint32_t sbase) {
        cur = kNoNode; sp = sbase + kStackStride; leaf_pos = leaf_end = 0; n_tests = n_hits = n_tris = 0;
        pend_ref = kNoNode; pend_key = 0.0f;
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
        pend_ref = kNoNode;
        // the root's own box (bvh.rs:176-179 on node 0)
        float lo, hi;
        slab(sc.root_min[0], sc.root_min[1], sc.root_min[2], sc.root_max[0], sc.root_max[1], sc.root_max[2], &lo, &hi);
        if (lo <= fminf(hi, t_max)) {
            if (COUNTS) n_hits = 1;
            enter<GENERIC>(sc, sc.root_ref);
        }
    }
    // The deferred box test of a stacked child: the reference's `lo <= min(hi, t_max)` (f32::min ignores a NaN) with
    // `lo <= hi` already folded into the key. Compared as unsigned bit patterns: lo >= +0 and t_max >= +0 are ordered like
    // their bits; a NaN t_max (a hit whose t is NaN: zero-length direction, inf * 0 in the edge functions) is 0x7fffffff
    // and, as in the reference, stops constraining the boxes; kKeyNever fails against every t_max; the sentinel's key 0
    // passes against every t_max, so a pop can never run below the stack.
#ifdef YK_KEY_FLOAT_COMPARE  // A/B only: the float compare this replaced (runs below the stack when t_max becomes NaN)
    __device__ __forceinline__ bool key_fails(float key) const { return !(key <= t_max); }
#else
    __device__ __forceinline__ bool key_fails(float key) const { return __float_as_uint(key) > __float_as_uint(t_max); }
#endif
    __device__ __forceinline__ bool wants_box() const { return cur != kNoNode; }
    __device__ __forceinline__ bool wants_tri() const { return leaf_pos < leaf_end; }
    __device__ __forceinline__ void push(uint32_t sbase, uint32_t* deep_ref, float* deep_key, uint32_t ref, float key) {
        if (sp < sbase + (uint32_t)kShortStack * kStackStride) {
            sts_entry(sp, ref, key);
        } else {  // cold: the stack continues in local memory, up to the reference's 64 entries
            const uint32_t depth = (sp - sbase) / kStackStride - kShortStack;
            YK_ASSERT(depth < (uint32_t)kDeepStack);  // bvh.rs:172-174: 64 entries
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
            } while (key_fails(key));
        } else {
            do {
                sp -= kStackStride;
                YK_ASSERT(sp >= sbase && sp < sbase + (uint32_t)kShortStack * kStackStride);
                lds_entry(sp, &ref, &key);
                if (ANYHIT) n_tests += ref != kNoNode ? 1u : 0u;
            } while (key_fails(key));
        }
        if (COUNTS) n_hits += ref != kNoNode ? 1u : 0u;
        return ref;
    }
    // Enters the interior node `cur`: the box tests of its two children (bvh.rs:176-199, math/bounds.rs:176-215).
    template <bool COUNTS, bool ANYHIT, bool GENERIC>
    __device__ __forceinline__ void box_step(const DevScene& sc, uint32_t sbase, uint32_t* deep_ref, float* deep_key) {
        // near child first: the second child when the ray is negative on the split axis (bvh.rs:186-194)
        const uint32_t neg = (neg_mask >> ((cur >> 29) & 3u)) & 1u;
        YK_ASSERT((cur & kRefIndexMask) < sc.n_nodes);
        const float4* rec = sc.nodes2 + 4 * (size_t)(cur & kRefIndexMask);
        const float4* near = rec + 2 * neg;
        const float4* far = rec + 2 * (neg ^ 1u);
#if YK_LD256
        float4 n0, n1, f0, f1;  // one child's half of the record (32 bytes, 32-byte aligned) per LDG.256
        ldg256(near, &n0, &n1);
        ldg256(far, &f0, &f1);
#else
        const float4 n0 = __ldg(near), n1 = __ldg(near + 1);
        const float4 f0 = __ldg(far), f1 = __ldg(far + 1);
#endif
        float lo_n, hi_n, lo_f, hi_f;
        slab(n0.x, n0.y, n0.z, n1.x, n1.y, n1.z, &lo_n, &hi_n);
        slab(f0.x, f0.y, f0.z, f1.x, f1.y, f1.z, &lo_f, &hi_f);
        const uint32_t ref_n = __float_as_uint(n0.w), ref_f = __float_as_uint(f0.w);
        const bool hit_n = lo_n <= fminf(hi_n, t_max);
        const bool ok_f = !(lo_f > hi_f);  // can the far child pass at all? (a NaN hi is ignored by the reference's min)
        const float key_f = ok_f ? lo_f : __uint_as_float(kKeyNever);
        // near missed: nothing happens before the far child's test, t_max is what the pop would see
        const bool hit_f = !hit_n && !key_fails(key_f);
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
                    YK_ASSERT(sp >= sbase && sp < sbase + (uint32_t)kShortStack * kStackStride);
                    lds_entry(sp, &take, &key);
                    if (ANYHIT) n_tests += take != kNoNode ? 1u : 0u;
                } while (key_fails(key));
                if (COUNTS) n_hits += take != kNoNode ? 1u : 0u;
            }
        }
        enter<GENERIC>(sc, take);
    }
    // One triangle test of the parked leaf (shapes/triangle.rs:62-130 on the permuted, origin-relative vertices).
    // Returns true on a hit with t in (0, t_max]; the caller decides what a hit means and then calls leaf_done().
    __device__ __forceinline__ bool tri_step(const DevScene& sc, uint32_t* tri, float* t_scaled_out, float* det_out, int* area_light) {
        const uint32_t s = leaf_pos++;
        YK_ASSERT(s < sc.n_tris && kx < 3 && ky < 3 && kz < 3);
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

    // ---- postponed leaves (YK_POSTPONE, closest hit without per-ray counters) ---------------------------------------
    // A lane that reaches a leaf during the box phase keeps walking with that leaf pending; a second leaf parks it. Pending
    // leaves are tested in the order they were found, the second one behind a re-test of its own box key against the t_max
    // the first one left — a child's slab entry is never below its parent's, so a leaf the reference would have culled
    // through any ancestor fails its own key — hence exactly the reference's sequence of shape tests and the same hit, ties
    // included; only the number of *box* tests grows (those taken with the stale t_max). scripts/sim_spec.py is the model.
    template <bool GENERIC>
    __device__ __forceinline__ void set_leaf(const DevScene& sc, uint32_t ref) {
        uint32_t first, count;
        if (GENERIC && sc.leaf_table) {
            const uint2 l = __ldg(&sc.leaf_table[ref]);
            first = l.x; count = l.y;
        } else {
            first = ref & ((1u << kLeafFirstBits) - 1u);
            count = (ref >> kLeafFirstBits) + 1u;
        }
        leaf_pos = first; leaf_end = first + count;
    }
    // pop_passing that also returns the key and leaves the sentinel in place (it may be popped again after the pending leaves)
    __device__ __forceinline__ uint32_t pop_keyed(uint32_t sbase, const uint32_t* deep_ref, const float* deep_key, float* key_out) {
        uint32_t ref;
        float key;
        do {
            sp -= kStackStride;
            if (sp < sbase + (uint32_t)kShortStack * kStackStride) {
                lds_entry(sp, &ref, &key);
            } else {
                const uint32_t depth = (sp - sbase) / kStackStride - kShortStack;
                ref = deep_ref[depth];
                key = deep_key[depth];
            }
        } while (key_fails(key));
        sp += ref == kNoNode ? kStackStride : 0u;
        *key_out = key;
        return ref;
    }
    template <bool GENERIC>
    __device__ __forceinline__ void enter_post(const DevScene& sc, uint32_t sbase, const uint32_t* deep_ref, const float* deep_key, uint32_t ref, float key) {
        if ((int32_t)ref < 0) { cur = ref; return; }  // interior, or the sentinel
        cur = kNoNode;
        if (leaf_pos < leaf_end) { pend_ref = ref; pend_key = key; return; }  // a second leaf: parked
        set_leaf<GENERIC>(sc, ref);
        float k2;
        const uint32_t r2 = pop_keyed(sbase, deep_ref, deep_key, &k2);  // walk on (t_max is stale until the leaf is tested)
        if ((int32_t)r2 < 0) cur = r2;
        else { pend_ref = r2; pend_key = k2; }
    }
    template <bool GENERIC>
    __device__ __forceinline__ void leaf_done_post(const DevScene& sc, uint32_t sbase, const uint32_t* deep_ref, const float* deep_key) {
        if (leaf_pos != leaf_end) return;
        if (pend_ref != kNoNode) {
            const uint32_t r = pend_ref;
            const float k = pend_key;
            pend_ref = kNoNode;
            if (!key_fails(k)) { set_leaf<GENERIC>(sc, r); return; }  // the pending leaf's deferred box test, with the current t_max
        }
        if (cur == kNoNode) {  // parked (or finished): next entry; no speculation inside the shape phase
            float k2;
            const uint32_t r2 = pop_keyed(sbase, deep_ref, deep_key, &k2);
            if ((int32_t)r2 < 0) cur = r2;
            else set_leaf<GENERIC>(sc, r2);
        }
    }
    template <bool GENERIC>
    __device__ __forceinline__ void box_step_post(const DevScene& sc, uint32_t sbase, uint32_t* deep_ref, float* deep_key) {
        const uint32_t neg = (neg_mask >> ((cur >> 29) & 3u)) & 1u;
        const float4* rec = sc.nodes2 + 4 * (size_t)(cur & kRefIndexMask);
        const float4* near = rec + 2 * neg;
        const float4* far = rec + 2 * (neg ^ 1u);
        float4 n0, n1, f0, f1;
        ldg256(near, &n0, &n1);
        ldg256(far, &f0, &f1);
        float lo_n, hi_n, lo_f, hi_f;
        slab(n0.x, n0.y, n0.z, n1.x, n1.y, n1.z, &lo_n, &hi_n);
        slab(f0.x, f0.y, f0.z, f1.x, f1.y, f1.z, &lo_f, &hi_f);
        const uint32_t ref_n = __float_as_uint(n0.w), ref_f = __float_as_uint(f0.w);
        const bool hit_n = lo_n <= fminf(hi_n, t_max);
        const bool ok_f = !(lo_f > hi_f);
        const float key_f = ok_f ? lo_f : __uint_as_float(kKeyNever);
        const bool hit_f = !hit_n && !key_fails(key_f);
        n_tests += 2u;
        if (hit_n && ok_f) push(sbase, deep_ref, deep_key, ref_f, key_f);
        uint32_t take = hit_n ? ref_n : ref_f;
        float take_key = hit_n ? lo_n : key_f;
        if (!(hit_n || hit_f)) take = pop_keyed(sbase, deep_ref, deep_key, &take_key);
        enter_post<GENERIC>(sc, sbase, deep_ref, deep_key, take, take_key);
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


// The same two phases with postponed leaves (closest hit, no per-ray counters): see TraceLane::enter_post.
#define YK_TRACE_PHASES_POST(LANE, LIVE, GENERIC, ON_HIT)                                                         \
    for (;;) {                                                                                                    \
        const bool want_n = (LANE).wants_box();                                                                   \
        const int n_n = __popc(__ballot_sync(0xffffffffu, want_n));                                               \
        if (n_n == 0) break;                                                                                      \
        if (n_n < kNodePhaseMin && __ballot_sync(0xffffffffu, (LIVE) && !want_n)) break;                          \
        if (want_n) (LANE).template box_step_post<GENERIC>(sc, sbase, deep_ref, deep_key);                        \
        if (YK_BOX_STEPS_PER_VOTE > 1 && (LANE).wants_box()) (LANE).template box_step_post<GENERIC>(sc, sbase, deep_ref, deep_key); \
        if (YK_BOX_STEPS_PER_VOTE > 2 && (LANE).wants_box()) (LANE).template box_step_post<GENERIC>(sc, sbase, deep_ref, deep_key); \
    }                                                                                                             \
    while (__ballot_sync(0xffffffffu, (LANE).wants_tri())) {                                                      \
        if ((LANE).wants_tri()) {                                                                                 \
            uint32_t tri_; float ts_, det_; int al_;                                                              \
            if ((LANE).tri_step(sc, &tri_, &ts_, &det_, &al_)) { ON_HIT }                                         \
            (LANE).template leaf_done_post<GENERIC>(sc, sbase, deep_ref, deep_key);                               \
        }                                                                                                         \
    }
#ifndef YK_POSTPONE
#define YK_POSTPONE 0
#endif

// Closest hit: BoundingVolumeHierarchy::intersect (bvh.rs:160-232). One ray per queue entry.
template <bool COUNTS, bool SPHERES>
__global__ void __launch_bounds__(kTraceThreads, YK_TRACE_MIN_BLOCKS) k_trace_closest(DevScene sc, Wave w, int b, IterCounters* cur, const uint32_t* perm) {
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
                    // the queue slot: rays, hits and counters of a bounce are all in queue order (sorted renders fetch it through perm)
                    path = perm ? __ldg(&perm[mine]) : mine;
                    YK_ASSERT(path < n);
                    const float4 ro = ld_once(&w.st[b].ray_o[path]);
                    const float4 rd = ld_once(&w.st[b].ray_d[path]);
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
            // What a closest-hit walk does with an accepted shape test.
            //  * A sphere slot is recognised by its tag (row 0's w = -2 - sphere index; triangles carry an area light >= -1), not
            //    by the NaN determinant its NaN vertex lanes produce: a real triangle gives a NaN determinant too (zero-length or
            //    NaN direction, non-finite vertex) and must stay on the triangle branch, where the reference returns a NaN-t hit.
            //    The sphere test (shapes/sphere.rs:36-77) re-reads the direction, which is not kept in registers, on this rare path.
            //  * Triangle (triangle.rs:133-139): a later equal-t hit replaces the earlier one (bvh.rs:204-207).
            //  * A NaN hit distance (NaN determinant: zero-length or NaN direction) makes the reference's later box tests
            //    `lo <= min(hi, NaN)`: `lo <= hi` for a box with a numeric exit distance — what the stacked keys encode — but false
            //    for a box whose own exit distance is NaN too, i.e. every box when all three direction components are NaN. Those
            //    rays drop their stack here (their popped nodes are already counted as tested, none would pass). Only the counting
            //    instantiations carry the test: they serve yk_trace's caller rays and the debug integrator; the integrators' own
            //    rays have finite directions, and the test costs the hot instantiation 1.7 % (profiles/r02/ab_cornell_regression.txt).
#define YK_CLOSEST_ON_HIT {                                                                                                     \
                if (SPHERES && al_ <= -2) {                                                                                      \
                    float t_s;                                                                                                   \
                    if (sphere_slot_test(sc.spheres, al_, tl.ox, tl.oy, tl.oz, w.st[b].ray_d[path], tl.t_max, &t_s)) {           \
                        hit_tri = tri_; hit_t = t_s; tl.t_max = t_s;                                                            \
                    }                                                                                                            \
                } else {                                                                                                         \
                    const float inv_det = 1.0f / det_;                                                                           \
                    hit_tri = tri_; hit_t = ts_ * inv_det; tl.t_max = hit_t;                                                     \
                    if (COUNTS && hit_t != hit_t && tl.ix != tl.ix && tl.iy != tl.iy && tl.iz != tl.iz) tl.sp = sbase + kStackStride; \
                }                                                                                                                \
            }
            if constexpr (YK_POSTPONE != 0 && !COUNTS) {
                YK_TRACE_PHASES_POST(tl, live, SPHERES, YK_CLOSEST_ON_HIT)
            } else {
                YK_TRACE_PHASES(tl, live, COUNTS, false, SPHERES, YK_CLOSEST_ON_HIT)
            }
#undef YK_CLOSEST_ON_HIT
            if (live && !tl.wants_box()) {  // retire
                st_once(&w.hit[path], make_uint2(__float_as_uint(hit_t), hit_tri));
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
#ifndef YK_SHADOW_PREFETCH
#define YK_SHADOW_PREFETCH 1  // measured: shadow / fold kernel -2.6 %, Cornell render +0.7 %; reserving one chunk ahead (2) loses it to spills
#endif
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

    // The kernel's inputs are streamed once, 16 - 40 B per array and path, and a lane needs them the moment it picks a path up:
    // ncu shows the kernel waiting on exactly these loads (long-scoreboard ~6 warps per issue, issue-active 60 %,
    // profiles/r01/ncu_shadow_cornell.txt). So a warp pulls a chunk's lines (the hand-over arrays of light 0 and the pending
    // terms: 38 lines of 128 B for 64 paths) into the L2 when it reserves the chunk.
    auto prefetch_chunk = [&](uint32_t base) {
        if (base >= n) return;
        const int a = lane >> 3;  // 0 pend_extra, 1 lt_o, 2 lt_d, 3 pend_beta: 8 lines each
        const char* p = a == 0 ? (const char*)(w.pend_extra + base) : a == 1 ? (const char*)(w.lt_o + base)
                      : a == 2 ? (const char*)(w.lt_d + base) : (const char*)(w.pend_beta + base);
        // (no per-line bounds test: the arrays are allocated with a chunk of slack, ensure_wave)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + (lane & 7) * 128));
        if (lane < 6) {
            const char* q = lane < 4 ? (const char*)(w.lt_c + base) + lane * 128 : (const char*)(w.sh_path + base) + (lane - 4) * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
    };
#if YK_SHADOW_PREFETCH == 2
    uint32_t ahead = 0;
    {
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(cursor, kChunk);
        ahead = __shfl_sync(0xffffffffu, b0, 0);
        prefetch_chunk(ahead);
    }
#endif
    auto finish_path = [&]() {  // path.rs:121-129
        const float4 pe = w.pend_extra[pos], pb = w.pend_beta[pos];
        RGB r = radiance + rgb(pe.x, pe.y, pe.z);
        if (pb.w > 0.0f) r = rgb(fminf(r.r, cfg.clamp), fminf(r.g, cfg.clamp), fminf(r.b, cfg.clamp));
        if (pb.w < 0.0f) {  // whitted.rs:109-130 (w = -1): the node's own radiance, summed up the tree by k_tree_return
            w.pend_extra[pos] = make_float4(r.r, r.g, r.b, pe.w);
            return;
        }
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
#if YK_SHADOW_PREFETCH == 2
                uint32_t base = ahead;  // reserved (and prefetched) while the previous chunk was being consumed
                uint32_t nxt_base = 0;
                if (lane == 0) nxt_base = atomicAdd(cursor, kChunk);
                ahead = __shfl_sync(0xffffffffu, nxt_base, 0);
                prefetch_chunk(ahead);
#else
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(cursor, kChunk);
                base = __shfl_sync(0xffffffffu, base, 0);
#if YK_SHADOW_PREFETCH == 1
                prefetch_chunk(base);
#endif
#endif
                chunk_next = base;
                chunk_end = base + kChunk < n ? base + kChunk : n;
                if (base >= n) { exhausted = true; chunk_next = chunk_end = 0; break; }
            }
            const uint32_t mine = chunk_next + __popc(idle & lt_mask);
            if (!live && mine < chunk_end) {
                pos = mine;
                YK_ASSERT(pos < w.cap);
                path = w.sh_path[pos];
                YK_ASSERT(path < w.cap);
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
                if (SPHERES && al_ <= -2) {  // sphere slot (tag, see k_trace_closest): run the real test; spheres carry no area light
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
// ---- shadow rays, one per lane ---------------------------------------------------------------------------------------
// The alternative to k_trace_shadow's one-path-per-lane walk (a path's lights traced back to back by one lane: with L lights
// the lane's dependent chain and the warp's refill imbalance grow L-fold). Here the work item is one (shading position g,
// light k) pair. Items are enumerated light-major in chunks of kChunk consecutive shading positions, so the lanes of a warp
// trace rays towards the same light from neighbouring hit points — coherent any-hit walks, coalesced hand-over reads. A lane
// whose item's bit is clear in sh_mask[g] (no shadow ray for that light) takes another item. An occluded ray clears its bit
// (only the lane that owns (g, k) ever touches bit k, so reading it needs no ordering); k_shadow_fold then adds the
// contributions of the bits that are left, in light order, exactly like the reference's fold (path.rs:102-119).
template <bool SPHERES>
__global__ void __launch_bounds__(kTraceThreads, YK_SHADOW_MIN_BLOCKS) k_trace_shadow_rays(DevScene sc, Wave w, uint32_t n_lights, IterCounters* cur) {
    uint32_t* const cursor = &cur->work_shadow;
    __shared__ uint2 s_stack[kShortStack][kTraceThreads];
    uint32_t deep_ref[kDeepStack];
    float deep_key[kDeepStack];
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(&s_stack[0][tid]);
    sts_entry(sbase, kNoNode, 0.0f);
    const uint32_t n = cur->mat[0] + cur->mat[1] + cur->mat[2] + cur->mat[3];
    const uint32_t chunks_per_light = (n + kChunk - 1) / kChunk;
    const uint32_t n_chunks = chunks_per_light * n_lights;
    uint32_t n_rays = 0;
    uint32_t chunk_next = 0, chunk_end = 0, chunk_light = 0;  // warp-uniform
    bool exhausted = false;

    TraceLane tl;
    tl.idle(sbase);
    bool live = false, occluded = false;
    uint32_t my_g = 0, my_k = 0;
    int target_light = -1;

    for (;;) {
        for (int round = 0; round < 4; ++round) {
            const unsigned idle = __ballot_sync(0xffffffffu, !live);
            if (!idle || exhausted) break;
            if (chunk_next >= chunk_end) {
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(cursor, 1u);
                c = __shfl_sync(0xffffffffu, c, 0);
                if (c >= n_chunks) { exhausted = true; chunk_next = chunk_end = 0; break; }
                chunk_light = c / chunks_per_light;
                chunk_next = (c - chunk_light * chunks_per_light) * kChunk;
                chunk_end = chunk_next + kChunk < n ? chunk_next + kChunk : n;
            }
            const uint32_t mine = chunk_next + __popc(idle & lt_mask);
            if (!live && mine < chunk_end && ((w.sh_mask[mine] >> chunk_light) & 1u)) {
                my_g = mine; my_k = chunk_light;
                YK_ASSERT(my_g < w.cap && my_k < n_lights);
                const size_t ref = (size_t)my_k * w.cap + my_g;
                const float4 ro = ld_once(&w.lt_o[ref]), rd = ld_once(&w.lt_d[ref]);
                target_light = __float_as_int(ld_once(&w.lt_c[ref]).y);
                tl.template start<false, SPHERES>(sc, sbase, ro.x, ro.y, ro.z, rd.x, rd.y, rd.z, 0.9999f);  // interaction.rs:57-58
                occluded = false;
                live = true;
                n_rays += 1;
            }
            const uint32_t taken = chunk_next + __popc(idle);
            chunk_next = taken < chunk_end ? taken : chunk_end;
        }
        if (__ballot_sync(0xffffffffu, live) == 0) {
            if (exhausted) break;
            continue;
        }
        for (;;) {
            YK_TRACE_PHASES(tl, live, false, true, SPHERES, {
                (void)tri_; (void)ts_; (void)det_;
                bool blocks = true;
                if (SPHERES && al_ <= -2) {  // sphere slot: run the real test; spheres carry no area light
                    float t_s;
                    blocks = sphere_slot_test(sc.spheres, al_, tl.ox, tl.oy, tl.oz, w.lt_d[(size_t)my_k * w.cap + my_g], tl.t_max, &t_s);
                } else if (target_light >= 0 && al_ >= 0 && al_ == target_light) {
                    blocks = false;  // bvh.rs:269-280: the target light's own emissive triangles do not occlude
                }
                if (blocks) { occluded = true; tl.stop(sbase); }
            })
            if (live && !tl.wants_box()) {  // this shadow ray is done
                if (occluded) atomicAnd(&w.sh_mask[my_g], ~(1u << my_k));
                live = false;
            }
            const int busy = __popc(__ballot_sync(0xffffffffu, live));
            if (busy == 0 || (!exhausted && busy < kRefillBelow)) break;
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

// The fold after k_trace_shadow_rays: `c + f*li*cos/pdf` over the unoccluded lights in light order, `radiance += beta * Le`,
// the indirect clamp and `L += beta * radiance` (path.rs:113-129); Whitted leaves the node's radiance for k_tree_return
// (whitted.rs:120-130). One shading position per thread, streamed.
__global__ void k_shadow_fold(Wave w, RenderCfg cfg, const IterCounters* cur) {
    const uint32_t n = cur->mat[0] + cur->mat[1] + cur->mat[2] + cur->mat[3];
    for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
        uint32_t mask = ld_once(&w.sh_mask[g]);
        RGB radiance = gray(0.0f);
        while (mask) {
            const size_t ref = (size_t)(__ffs(mask) - 1) * w.cap + g;
            radiance = radiance + rgb(ld_once(&w.lt_o[ref]).w, ld_once(&w.lt_d[ref]).w, ld_once(&w.lt_c[ref]).x);
            mask &= mask - 1;
        }
        const float4 pe = ld_once(&w.pend_extra[g]), pb = ld_once(&w.pend_beta[g]);
        RGB r = radiance + rgb(pe.x, pe.y, pe.z);
        if (pb.w > 0.0f) r = rgb(fminf(r.r, cfg.clamp), fminf(r.g, cfg.clamp), fminf(r.b, cfg.clamp));
        if (pb.w < 0.0f) {  // whitted.rs:109-130 (w = -1): the node's own radiance, summed up the tree by k_tree_return
            w.pend_extra[g] = make_float4(r.r, r.g, r.b, pe.w);
            continue;
        }
        const uint32_t path = ld_once(&w.sh_path[g]);
        float4 L = w.L[path];
        L.x = L.x + pb.x * r.r;
        L.y = L.y + pb.y * r.g;
        L.z = L.z + pb.z * r.b;
        w.L[path] = L;
    }
}

}  // namespace
