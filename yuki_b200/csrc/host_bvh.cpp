// Host BVH builder for the B200 backend: produces, directly in pre-order, the same 32-byte linear
// nodes and the same leaf ordering as yuki's BoundingVolumeHierarchy::new (yuki/src/bvh.rs:39-115,
// recursive_build :305-390, split_* :422-523, flatten_tree :396-419), so that device traversal visits
// exactly the nodes the reference visits (its intersection_test_count, bvh.rs:177, is a parity gate).
//
// Not a port of the reference's two-phase build (arena tree + flatten): subtrees are emitted straight
// into a flat vector, large subtrees are built concurrently and spliced (their leaf ranges are
// disjoint slices of the primitive array, so the result does not depend on the schedule), and the
// passes over the biggest nodes (bounds, SAH buckets, the partition) run on several threads.
#include <algorithm>
#include <atomic>
#include <future>
#include <thread>
#include <vector>

#include "host_math.h"
#include "yuki_gpu.h"

namespace ykh {

struct Prim {
    uint32_t index;
    box3 box;
    f3 key;  // the reference's "centroid": p_min + diagonal / 0.5 (bvh.rs:56) — twice the diagonal, reproduced
};

struct Builder {
    Prim* prims;
    uint32_t leaf_max;
    uint32_t n_total = 0;
    uint32_t method;
    std::atomic<bool> failed{false};

    static constexpr int kBuckets = 12;
    static constexpr size_t kParallelCutoff = 1u << 16;

    // centroid_bounds.offset(c)[axis] (yuki_derive/src/impl_bounds.rs:230-234), then the bucket index of
    // bvh.rs:476-479: (12*o).max(0) as usize, min 11. `as usize` truncates and saturates, NaN -> 0.
    static int bucket_index(const box3& kb, const Prim& p, int axis) {
        float o = p.key.get(axis) - kb.lo.get(axis);
        const float ext_hi = kb.hi.get(axis), ext_lo = kb.lo.get(axis);
        if (ext_hi != ext_lo) o /= ext_hi - ext_lo;
        const float scaled = fmaxf((float)kBuckets * o, 0.0f);
        if (!(scaled == scaled)) return 0;
        if (scaled >= (float)kBuckets) return kBuckets - 1;
        return std::min((int)scaled, kBuckets - 1);
    }

    // itertools::partition (third-party, itertools 0.10): two-ended scan with swaps; deterministic.
    template <class Keep>
    static size_t stable_front_partition(Prim* a, size_t n, Keep keep) {
        size_t lo = 0, hi = n, kept = 0;
        while (lo < hi) {
            Prim& front = a[lo++];
            if (!keep(front)) {
                bool swapped = false;
                while (lo < hi) {
                    Prim& back = a[--hi];
                    if (keep(back)) {
                        std::swap(front, back);
                        swapped = true;
                        break;
                    }
                }
                if (!swapped) break;
            }
            ++kept;
        }
        return kept;
    }

    // ---- the top of the tree: one node's passes over millions of primitives, spread over threads -------------------
    // (below kWideNode the concurrently built subtrees provide the parallelism). Every pass is order-independent — min /
    // max unions and counts — except the partition, whose parallel form reproduces the sequential permutation exactly.
    static constexpr size_t kWideNode = 1u << 19;
    unsigned workers = 1;

    template <class F>
    void for_chunks(size_t n, F&& f) const {  // f(chunk, begin, end) on `workers` threads; chunk < workers
        const unsigned w = (unsigned)std::min<size_t>(workers, std::max<size_t>(1, n / 65536));
        std::vector<std::future<void>> futs;
        for (unsigned c = 1; c < w; ++c) futs.push_back(std::async(std::launch::async, [&f, c, w, n] { f(c, n * c / w, n * (c + 1) / w); }));
        f(0u, (size_t)0, n / w);
        for (auto& x : futs) x.get();
    }
    // itertools::partition's result, computed in parallel: the sequential two-ended scan swaps the k-th failing element
    // met from the front with the k-th passing element met from the back, and the two scans meet at K = the number of
    // passing elements; so the failing positions below K pair, in order, with the passing positions at or above K taken
    // from the back. Elements outside those two sets never move.
    template <class Keep>
    size_t wide_front_partition(Prim* a, size_t n, Keep keep) const {
        std::vector<uint8_t> pass(n);
        std::vector<size_t> cnt(workers + 1, 0);
        for_chunks(n, [&](unsigned c, size_t b, size_t e) {
            size_t k = 0;
            for (size_t i = b; i < e; ++i) k += (pass[i] = keep(a[i]) ? 1 : 0);
            cnt[c + 1] = k;
        });
        size_t K = 0;
        for (size_t c : cnt) K += c;
        std::vector<size_t> fail_front, pass_back;  // increasing positions
        {
            std::vector<std::vector<size_t>> ff(workers), pb(workers);
            for_chunks(n, [&](unsigned c, size_t b, size_t e) {
                for (size_t i = b; i < e; ++i) {
                    if (i < K && !pass[i]) ff[c].push_back(i);
                    else if (i >= K && pass[i]) pb[c].push_back(i);
                }
            });
            for (auto& v : ff) fail_front.insert(fail_front.end(), v.begin(), v.end());
            for (auto& v : pb) pass_back.insert(pass_back.end(), v.begin(), v.end());
        }
        const size_t m = fail_front.size();  // == pass_back.size()
        for_chunks(m, [&](unsigned, size_t b, size_t e) {
            for (size_t k = b; k < e; ++k) std::swap(a[fail_front[k]], a[pass_back[m - 1 - k]]);
        });
        return K;
    }
    template <class Keep>
    size_t front_partition(Prim* a, size_t n, Keep keep) const {
        return (workers > 1 && n >= kWideNode) ? wide_front_partition(a, n, keep) : stable_front_partition(a, n, keep);
    }
    box3 prim_bounds(size_t lo, size_t hi) const {
        if (workers > 1 && hi - lo >= kWideNode) {
            std::vector<box3> part(workers, empty_box());
            for_chunks(hi - lo, [&](unsigned c, size_t b, size_t e) {
                box3 x = empty_box();
                for (size_t i = lo + b; i < lo + e; ++i) x = merge(x, prims[i].box);
                part[c] = x;
            });
            box3 x = empty_box();
            for (const box3& p : part) x = merge(x, p);
            return x;
        }
        box3 x = empty_box();
        for (size_t i = lo; i < hi; ++i) x = merge(x, prims[i].box);
        return x;
    }
    box3 key_bounds(size_t lo, size_t hi) const {
        if (workers > 1 && hi - lo >= kWideNode) {
            std::vector<box3> part(workers, empty_box());
            for_chunks(hi - lo, [&](unsigned c, size_t b, size_t e) {
                box3 x = empty_box();
                for (size_t i = lo + b; i < lo + e; ++i) x = grow(x, prims[i].key);
                part[c] = x;
            });
            box3 x = empty_box();
            for (const box3& p : part) x = merge(x, p);
            return x;
        }
        box3 x = empty_box();
        for (size_t i = lo; i < hi; ++i) x = grow(x, prims[i].key);
        return x;
    }

    size_t median_split(size_t lo, size_t hi, int axis) {  // split_equal_counts, bvh.rs:422-436
        const size_t mid = (lo + hi) / 2;
        std::nth_element(prims + lo, prims + mid, prims + hi,
                         [axis](const Prim& a, const Prim& b) { return a.key.get(axis) < b.key.get(axis); });
        return mid;
    }

    // What a SAH split already knows about its two sides: the unions of the bucket boxes on either side of the chosen plane are
    // the children's shape bounds and "centroid" bounds (min / max unions are exact and association-free), so the children
    // need not scan their primitives again for them — two of the four passes a node makes over its primitives.
    struct ChildBounds {
        bool valid = false;
        box3 box[2], kb[2];
    };
    // Returns the split position, `lo` when the method declines (caller falls back to the median), or
    // SIZE_MAX when SAH prefers a leaf (bvh.rs:510-521).
    size_t choose_split(const box3& box, const box3& kb, size_t lo, size_t hi, int axis, ChildBounds* cb) {
        const size_t n = hi - lo;
        if (method == YK_SPLIT_EQUAL_COUNTS) return median_split(lo, hi, axis);
        if (method == YK_SPLIT_MIDDLE) {  // split_middle, bvh.rs:438-450
            const float pivot = (kb.lo.get(axis) + kb.hi.get(axis)) / 2.0f;
            return lo + front_partition(prims + lo, n, [=](const Prim& p) { return p.key.get(axis) < pivot; });
        }
        if (n <= 2) return lo;  // bvh.rs:461-462
        size_t count[kBuckets] = {};
        box3 bbox[kBuckets], kbox[kBuckets];
        for (auto& b : bbox) b = empty_box();
        for (auto& b : kbox) b = empty_box();
        // The partition below asks for every primitive's bucket again: keep it, in the top four bits of the primitive's index
        // (scenes below 2^28 shapes; the bits are masked off when the leaf order is read out), instead of dividing twice.
        const bool tag = n_total < (1u << 28);
        if (workers > 1 && n >= kWideNode) {
            struct Part { size_t count[kBuckets]; box3 bbox[kBuckets], kbox[kBuckets]; };
            std::vector<Part> part(workers);
            for (Part& p : part)
                for (int b = 0; b < kBuckets; ++b) { p.count[b] = 0; p.bbox[b] = empty_box(); p.kbox[b] = empty_box(); }
            for_chunks(n, [&](unsigned c, size_t b0, size_t e0) {
                Part& p = part[c];
                for (size_t i = lo + b0; i < lo + e0; ++i) {
                    const int b = bucket_index(kb, prims[i], axis);
                    if (tag) prims[i].index = (prims[i].index & 0x0fffffffu) | ((uint32_t)b << 28);
                    p.count[b] += 1;
                    p.bbox[b] = merge(p.bbox[b], prims[i].box);
                    p.kbox[b] = grow(p.kbox[b], prims[i].key);
                }
            });
            for (const Part& p : part)
                for (int b = 0; b < kBuckets; ++b) {
                    count[b] += p.count[b];
                    bbox[b] = merge(bbox[b], p.bbox[b]);
                    kbox[b] = merge(kbox[b], p.kbox[b]);
                }
        } else {
            for (size_t i = lo; i < hi; ++i) {
                const int b = bucket_index(kb, prims[i], axis);
                if (tag) prims[i].index = (prims[i].index & 0x0fffffffu) | ((uint32_t)b << 28);
                count[b] += 1;
                bbox[b] = merge(bbox[b], prims[i].box);
                kbox[b] = grow(kbox[b], prims[i].key);
            }
        }
        // Suffix boxes/counts once instead of the reference's O(buckets^2) folds; min/max unions are exact
        // and association-free, so each side's box and count are identical.
        box3 right_box[kBuckets];
        size_t right_count[kBuckets];
        box3 acc = empty_box();
        size_t cacc = 0;
        for (int b = kBuckets - 1; b >= 1; --b) {
            acc = merge(bbox[b], acc);
            cacc += count[b];
            right_box[b] = acc;
            right_count[b] = cacc;
        }
        const float denom = fmaxf(half_area_x2(box), 1e-10f);
        box3 left = empty_box();
        size_t left_count = 0;
        int best = 0;
        float best_cost = 0.0f;
        for (int s = 0; s < kBuckets - 1; ++s) {
            left = merge(left, bbox[s]);
            left_count += count[s];
            const float cost =
                1.0f + ((float)left_count * half_area_x2(left) + (float)right_count[s + 1] * half_area_x2(right_box[s + 1])) / denom;
            if (s == 0 || cost < best_cost) {  // min_by keeps the first minimum (bvh.rs:504-508)
                best = s;
                best_cost = cost;
            }
        }
        if (!(best_cost < (float)n)) return SIZE_MAX;
        if (cb) {
            for (int side = 0; side < 2; ++side) { cb->box[side] = empty_box(); cb->kb[side] = empty_box(); }
            for (int b = 0; b < kBuckets; ++b) {
                if (!count[b]) continue;
                const int side = b <= best ? 0 : 1;
                cb->box[side] = merge(cb->box[side], bbox[b]);
                cb->kb[side] = merge(cb->kb[side], kbox[b]);
            }
            cb->valid = true;
        }
        if (tag) return lo + front_partition(prims + lo, n, [best](const Prim& p) { return (int)(p.index >> 28) <= best; });
        return lo + front_partition(prims + lo, n, [&](const Prim& p) { return bucket_index(kb, p, axis) <= best; });
    }

    static void emit_leaf(std::vector<yk_bvh_node>& out, const box3& box, size_t lo, size_t hi) {
        yk_bvh_node nd{};
        store3(box.lo, nd.p_min);
        store3(box.hi, nd.p_max);
        nd.offset = (uint32_t)lo;  // == ordered_shapes.len() at emission time in the reference's DFS
        nd.shape_count = (uint16_t)(hi - lo);
        nd.is_leaf = 1;
        out.push_back(nd);
    }

    // Appends the subtree over prims[lo, hi) to `out` in pre-order; returns its bounds. Node indices
    // written into `offset` are relative to out[0]; `depth` bounds how far down subtrees run as tasks.
    // (`known_box` / `known_kb`: the node's bounds when its parent's SAH split already produced them, else null.)
    box3 emit(std::vector<yk_bvh_node>& out, size_t lo, size_t hi, int task_depth, const box3* known_box = nullptr, const box3* known_kb = nullptr) {
        const box3 box = known_box ? *known_box : prim_bounds(lo, hi);
        const size_t n = hi - lo;
        if (n <= leaf_max) {
            emit_leaf(out, box, lo, hi);
            return box;
        }
        const box3 kb = known_kb ? *known_kb : key_bounds(lo, hi);
        const int axis = widest_axis(kb);
        if (kb.hi.get(axis) == kb.lo.get(axis)) {  // bvh.rs:343
            emit_leaf(out, box, lo, hi);
            return box;
        }
        ChildBounds cb;
        size_t mid = choose_split(box, kb, lo, hi, axis, &cb);
        if (method != YK_SPLIT_EQUAL_COUNTS && (mid == lo || mid == hi)) {
            mid = median_split(lo, hi, axis);
            cb.valid = false;  // another partition than the one the bucket unions describe
        }
        if (mid == lo) {  // assert_ne!(mid, start, "BVH: Split failed") — bvh.rs:368
            failed = true;
            emit_leaf(out, box, lo, hi);
            return box;
        }
        if (mid == SIZE_MAX) {
            emit_leaf(out, box, lo, hi);
            return box;
        }
        const size_t self = out.size();
        out.push_back(yk_bvh_node{});
        box3 lbox, rbox;
        uint32_t second;
        if (task_depth > 0 && n >= kParallelCutoff) {
            std::vector<yk_bvh_node> lsub, rsub;
            const box3 *lb = cb.valid ? &cb.box[0] : nullptr, *lk = cb.valid ? &cb.kb[0] : nullptr;
            const box3 *rb = cb.valid ? &cb.box[1] : nullptr, *rk = cb.valid ? &cb.kb[1] : nullptr;
            auto fut = std::async(std::launch::async, [&] { return emit(rsub, mid, hi, task_depth - 1, rb, rk); });
            lbox = emit(lsub, lo, mid, task_depth - 1, lb, lk);
            rbox = fut.get();
            const uint32_t lbase = (uint32_t)out.size();
            for (auto nd : lsub) {
                if (!nd.is_leaf) nd.offset += lbase;
                out.push_back(nd);
            }
            second = (uint32_t)out.size();
            for (auto nd : rsub) {
                if (!nd.is_leaf) nd.offset += second;
                out.push_back(nd);
            }
        } else {
            lbox = emit(out, lo, mid, 0, cb.valid ? &cb.box[0] : nullptr, cb.valid ? &cb.kb[0] : nullptr);
            second = (uint32_t)out.size();
            rbox = emit(out, mid, hi, 0, cb.valid ? &cb.box[1] : nullptr, cb.valid ? &cb.kb[1] : nullptr);
        }
        const box3 both = merge(lbox, rbox);  // BVHBuildNode::interior, bvh.rs:605-614
        yk_bvh_node& nd = out[self];
        store3(both.lo, nd.p_min);
        store3(both.hi, nd.p_max);
        nd.offset = second;
        nd.shape_count = 0;
        nd.split_axis = (uint8_t)axis;
        nd.is_leaf = 0;
        return both;
    }
};

// Deepest root-to-leaf path; the traversal stack holds 64 entries (bvh.rs:172-174).
static uint32_t tree_height(const yk_bvh_node* nodes, uint32_t n_nodes) {
    std::vector<std::pair<uint32_t, uint32_t>> stack{{0u, 1u}};
    uint32_t best = 0;
    while (!stack.empty()) {
        auto [i, d] = stack.back();
        stack.pop_back();
        if (i >= n_nodes) continue;
        best = std::max(best, d);
        if (!nodes[i].is_leaf) {
            stack.push_back({i + 1, d + 1});
            stack.push_back({nodes[i].offset, d + 1});
        }
    }
    return best;
}

int bvh_build_boxes(const float* boxes6, uint32_t n, uint32_t max_shapes_in_node, uint32_t split_method, std::vector<yk_bvh_node>* nodes,
                    std::vector<uint32_t>* order, const char** why);

int bvh_build(const float* tri_vertices, uint32_t n_tris, uint32_t max_shapes_in_node, uint32_t split_method,
              std::vector<yk_bvh_node>* nodes, std::vector<uint32_t>* order, const char** why) {
    if (!tri_vertices || n_tris == 0) {
        *why = "yk_bvh_build: empty triangle list";
        return YK_ERR_INVALID;
    }
    std::vector<float> boxes((size_t)n_tris * 6);
    for (uint32_t i = 0; i < n_tris; ++i) {
        const float* v = tri_vertices + (size_t)i * 9;
        // Triangle::world_bound (triangle.rs:229-235): Bounds3::new(p0, p1).union_p(p2)
        box3 b{min3(load3(v), load3(v + 3)), max3(load3(v), load3(v + 3))};
        b = grow(b, load3(v + 6));
        store3(b.lo, &boxes[(size_t)i * 6]);
        store3(b.hi, &boxes[(size_t)i * 6 + 3]);
    }
    return bvh_build_boxes(boxes.data(), n_tris, max_shapes_in_node, split_method, nodes, order, why);
}

// The build only sees world bounds (Shape::world_bound, shapes/mod.rs:33): `boxes6` = (p_min, p_max) per shape.
int bvh_build_boxes(const float* boxes6, uint32_t n_tris, uint32_t max_shapes_in_node, uint32_t split_method,
                    std::vector<yk_bvh_node>* nodes, std::vector<uint32_t>* order, const char** why) {
    if (!boxes6 || n_tris == 0) {
        *why = "yk_bvh_build: empty shape list";
        return YK_ERR_INVALID;
    }
    if (split_method > YK_SPLIT_EQUAL_COUNTS) {
        *why = "yk_bvh_build: unknown split method";
        return YK_ERR_INVALID;
    }
    std::vector<Prim> prims(n_tris);
    for (uint32_t i = 0; i < n_tris; ++i) {
        const box3 b{load3(boxes6 + (size_t)i * 6), load3(boxes6 + (size_t)i * 6 + 3)};
        prims[i] = {i, b, add(b.lo, divs(sub(b.hi, b.lo), 0.5f))};
    }
    Builder bld;
    bld.prims = prims.data();
    bld.leaf_max = max_shapes_in_node;
    bld.n_total = n_tris;
    bld.method = split_method;
    nodes->clear();
    nodes->reserve((size_t)2 * n_tris);
    unsigned hw = std::thread::hardware_concurrency();
    int task_depth = 0;
    while ((1u << task_depth) < hw && task_depth < 6) ++task_depth;
    bld.workers = std::max(1u, std::min(hw, 32u));
    bld.emit(*nodes, 0, n_tris, task_depth);
    if (bld.failed) {
        *why = "BVH: Split failed (bvh.rs:368)";
        return YK_ERR_BVH;
    }
    if (tree_height(nodes->data(), (uint32_t)nodes->size()) > 64) {
        *why = "BVH deeper than the 64-entry traversal stack (bvh.rs:172-174)";
        return YK_ERR_BVH;
    }
    order->resize(n_tris);
    for (uint32_t i = 0; i < n_tris; ++i) (*order)[i] = n_tris < (1u << 28) ? (prims[i].index & 0x0fffffffu) : prims[i].index;
    return YK_OK;
}

}  // namespace ykh
