// ORACLE — TEST INFRASTRUCTURE ONLY (see yko_math.h header).
//
// CPU restatement of yuki's scene-side hot-path types: Mesh/Triangle (shapes/{mesh,triangle}.rs),
// SurfaceInteraction (interaction.rs), BVH build + traversal (bvh.rs), textures (textures/*.rs),
// lights (lights/*.rs), camera (camera.rs). PARITY UNPINNED: the reference has no tests or golden
// vectors for any of these; each function follows the cited source line by line.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

#include "yko_math.h"
#include "yko_sampling.h"

namespace yko {

// ------------------------------------------------------------------------------------------------
// Textures (textures/constant.rs:23-30, textures/image_texture.rs:81-111)
enum TextureKind : uint32_t { TEX_CONSTANT = 0, TEX_IMAGE = 1 };
struct Texture {
    TextureKind kind;
    Spec value;            // constant; float textures use .r
    uint32_t width = 0, height = 0;
    std::vector<float> texels;  // RGB f32, row-major, first row = top of the image file
};

enum MaterialKind : uint32_t { MAT_MATTE = 0, MAT_GLASS = 1, MAT_METAL = 2, MAT_GLOSSY = 3 };
struct Material {
    MaterialKind kind;
    // matte: kd, sigma | glass: r, t | metal: eta, k, roughness | glossy: rs, roughness
    int32_t tex[3];
    float eta;             // glass
    bool remap_roughness;  // metal, glossy
};

enum LightKind : uint32_t { LIGHT_POINT = 0, LIGHT_SPOT = 1, LIGHT_RECT = 2, LIGHT_DISTANT = 3 };
struct Light {
    LightKind kind;
    V3 p;                  // point/spot position; distant: w
    Spec i;                // point/spot intensity; rect: L; distant: radiance
    float cos_total_width = 0, cos_falloff_start = 0;  // spot
    Transform world_to_light;   // spot
    Transform sample_to_world;  // rect
    float area = 0;             // rect
};

struct Mesh {
    std::vector<V3> points;   // world space (mesh.rs:27-29)
    std::vector<V3> normals;  // world space, NOT renormalised (mesh.rs:31-33)
    std::vector<V2> uvs;
    bool transform_swaps_handedness;
};

// One entry of the reference's `Vec<Arc<dyn Shape>>`: a mesh triangle (shapes/triangle.rs) or, when `sphere >= 0`, a
// sphere (shapes/sphere.rs).
struct Triangle {
    uint32_t mesh;
    uint32_t v[3];
    int32_t material;
    int32_t area_light;  // index into lights or -1
    uint32_t orig_id;    // index in the un-reordered shape list
    int32_t sphere = -1; // index into Scene::spheres
};

// shapes/sphere.rs:15-33
struct Sphere {
    Transform object_to_world, world_to_object;
    float radius;
    int32_t material;
    bool transform_swaps_handedness;
};

// interaction.rs:84-94
struct SurfaceInteraction {
    V3 p, n;
    V2 uv;
    V3 dpdu, dpdv;
    V3 sh_n, sh_dpdu, sh_dpdv;
    V3 wo;
    int32_t area_light;
};

struct Hit {
    float t;
    SurfaceInteraction si;
    uint32_t shape;  // index into ordered shapes
};

// ------------------------------------------------------------------------------------------------
// BVH (bvh.rs)
enum SplitMethod : uint32_t { SPLIT_SAH = 0, SPLIT_MIDDLE = 1, SPLIT_EQUAL_COUNTS = 2 };

// bvh.rs:536-588 — 32 bytes. `offset` is second_child_index (interior) or first_shape_index (leaf).
struct BVHNode {
    float p_min[3];
    float p_max[3];
    uint32_t offset;
    uint16_t shape_count;  // 0 for interior nodes
    uint8_t split_axis;
    uint8_t is_leaf;
};
static_assert(sizeof(BVHNode) == 32, "BVHNode must be 32 bytes (bvh.rs:556)");

struct TraversalStats {
    uint64_t closest_nodes = 0, closest_tris = 0, any_nodes = 0, any_tris = 0;
};

struct IntersectionResult {
    bool has_hit = false;
    Hit hit;
    uint64_t intersection_test_count = 0, intersection_count = 0;
};

struct Scene {
    std::vector<Mesh> meshes;
    std::vector<Triangle> shapes;  // BVH (leaf) order after build
    std::vector<Sphere> spheres;
    std::vector<BVHNode> nodes;
    std::vector<Texture> textures;
    std::vector<Material> materials;
    std::vector<Light> lights;
    Spec background{0, 0, 0};
    uint32_t max_shapes_in_node = 1;
    SplitMethod split_method = SPLIT_SAH;

    // shapes/triangle.rs:229-235, shapes/sphere.rs:121-123
    Bounds3 world_bound(const Triangle& t) const {
        if (t.sphere >= 0) {
            const Sphere& sp = spheres[t.sphere];
            const float r = sp.radius;
            return xf_bounds(sp.object_to_world, Bounds3{{-r, -r, -r}, {r, r, r}});
        }
        const Mesh& m = meshes[t.mesh];
        return union_p(bounds_new(m.points[t.v[0]], m.points[t.v[1]]), m.points[t.v[2]]);
    }

    bool triangle_intersect(const Triangle& tri, const Ray& ray, float* t_out, SurfaceInteraction* si,
                            bool want_si) const;
    bool sphere_intersect(const Sphere& sp, const Ray& ray, float* t_out, SurfaceInteraction* si, bool want_si) const;
    bool build_bvh();
    IntersectionResult intersect(Ray ray, TraversalStats* ts) const;
    bool any_intersect(const Ray& ray, int32_t area_light, TraversalStats* ts) const;
    Bounds3 bounds() const {
        const BVHNode& n = nodes[0];
        return {{n.p_min[0], n.p_min[1], n.p_min[2]}, {n.p_max[0], n.p_max[1], n.p_max[2]}};
    }
};

// shapes/sphere.rs:36-119
inline bool Scene::sphere_intersect(const Sphere& sp, const Ray& ray, float* t_out, SurfaceInteraction* si, bool want_si) const {
    const Ray r = xf_ray(sp.world_to_object, ray);
    const float a = r.d.x * r.d.x + r.d.y * r.d.y + r.d.z * r.d.z;
    const float b = 2.0f * (r.d.x * r.o.x + r.d.y * r.o.y + r.d.z * r.o.z);
    const float c = r.o.x * r.o.x + r.o.y * r.o.y + r.o.z * r.o.z - sp.radius * sp.radius;
    const float discrim = b * b - 4.0f * a * c;
    if (discrim < 0.0f) return false;
    const float rd = std::sqrt(discrim);
    const float q = b < 0.0f ? -0.5f * (b - rd) : -0.5f * (b + rd);
    float t0 = q / a, t1 = c / q;
    if (t0 > t1) std::swap(t0, t1);
    if (t0 > r.t_max || t1 <= 0.0f) return false;
    float t = t0;
    if (t <= 0.0f) {
        t = t1;
        if (t > r.t_max) return false;
    }
    *t_out = t;
    if (!want_si) return true;
    V3 p = r.o + r.d * t;
    p = p * (sp.radius / len(p - v3(0, 0, 0)));
    if (p.x == 0.0f && p.y == 0.0f) p.x = 1e-5f * sp.radius;
    const float kPiF = 3.14159274101257324f;
    float phi = std::atan2(p.y, p.x);
    if (phi < 0.0f) phi += 2.0f * kPiF;
    const float phi_max = 2.0f * kPiF, theta_min = kPiF, theta_max = 0.0f;
    const float u = phi / phi_max;
    const float cz = p.z / sp.radius;
    const float theta = std::acos(cz < -1.0f ? -1.0f : (cz > 1.0f ? 1.0f : cz));
    const float v = (theta - theta_min) / (theta_max - theta_min);
    const float z_radius = std::sqrt(p.x * p.x + p.y * p.y);
    const float inv_z_radius = 1.0f / z_radius;
    const float cos_phi = p.x * inv_z_radius, sin_phi = p.y * inv_z_radius;
    const V3 dpdu = v3(-phi_max * p.y, phi_max * p.x, 0.0f);
    const V3 dpdv = v3(p.z * cos_phi, p.z * sin_phi, -sp.radius * std::sin(theta)) * (theta_max - theta_min);
    // SurfaceInteraction::new (interaction.rs:95-124) in object space ...
    V3 n = normalized(cross(dpdu, dpdv));
    if (sp.transform_swaps_handedness) n = -n;
    // ... then &object_to_world * si (interaction.rs:141-164); note wo = -ray.d of the WORLD ray goes through the transform
    const Transform& o2w = sp.object_to_world;
    const V3 n_w = normalized(xf_normal(o2w, n));
    V3 sh_n = normalized(xf_normal(o2w, n));
    sh_n = faceforward_n(sh_n, n_w);
    si->p = xf_point(o2w, p);
    si->n = n_w;
    si->uv = {u, v};
    si->dpdu = xf_vec(o2w, dpdu);
    si->dpdv = xf_vec(o2w, dpdv);
    si->wo = normalized(xf_vec(o2w, -ray.d));
    si->sh_n = faceforward_n(sh_n, n_w);
    si->sh_dpdu = xf_vec(o2w, dpdu);
    si->sh_dpdv = xf_vec(o2w, dpdv);
    si->area_light = -1;
    return true;
}

// shapes/triangle.rs:49-227
inline bool Scene::triangle_intersect(const Triangle& tri, const Ray& ray, float* t_out, SurfaceInteraction* si,
                                      bool want_si) const {
    if (tri.sphere >= 0) return sphere_intersect(spheres[tri.sphere], ray, t_out, si, want_si);
    const Mesh& mesh = meshes[tri.mesh];
    V3 p0 = mesh.points[tri.v[0]], p1 = mesh.points[tri.v[1]], p2 = mesh.points[tri.v[2]];

    // :58-88 translate, permute, shear
    V3 p0t = p0 - ray.o, p1t = p1 - ray.o, p2t = p2 - ray.o;
    int kz = max_dimension(vabs(ray.d));
    int kx = kz < 2 ? kz + 1 : 0;
    int ky = kx < 2 ? kx + 1 : 0;
    p0t = permuted(p0t, kx, ky, kz);
    p1t = permuted(p1t, kx, ky, kz);
    p2t = permuted(p2t, kx, ky, kz);
    V3 d = permuted(ray.d, kx, ky, kz);
    float sx = -d.x / d.z, sy = -d.y / d.z, sz = 1.0f / d.z;
    p0t.x += sx * p0t.z; p0t.y += sy * p0t.z;
    p1t.x += sx * p1t.z; p1t.y += sy * p1t.z;
    p2t.x += sx * p2t.z; p2t.y += sy * p2t.z;

    // :91-106 edge functions with f64 fallback
    float e0 = p1t.x * p2t.y - p1t.y * p2t.x;
    float e1 = p2t.x * p0t.y - p2t.y * p0t.x;
    float e2 = p0t.x * p1t.y - p0t.y * p1t.x;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {
        e0 = (float)((double)p1t.x * (double)p2t.y - (double)p1t.y * (double)p2t.x);
        e1 = (float)((double)p2t.x * (double)p0t.y - (double)p2t.y * (double)p0t.x);
        e2 = (float)((double)p0t.x * (double)p1t.y - (double)p0t.y * (double)p1t.x);
    }
    // :109-117
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    // :120-130
    float p0z = p0t.z * sz, p1z = p1t.z * sz, p2z = p2t.z * sz;
    float t_scaled = e0 * p0z + e1 * p1z + e2 * p2z;
    if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < ray.t_max * det)) ||
        (det > 0.0f && (t_scaled <= 0.0f || t_scaled > ray.t_max * det)))
        return false;
    // :133-139
    float inv_det = 1.0f / det;
    float b0 = e0 * inv_det, b1 = e1 * inv_det, b2 = e2 * inv_det;
    *t_out = t_scaled * inv_det;
    if (!want_si) return true;

    // :143-172 partial derivatives
    V2 uv0{0.0f, 0.0f}, uv1{1.0f, 0.0f}, uv2{1.0f, 1.0f};
    if (!mesh.uvs.empty()) {
        uv0 = mesh.uvs[tri.v[0]]; uv1 = mesh.uvs[tri.v[1]]; uv2 = mesh.uvs[tri.v[2]];
    }
    V2 duv02 = uv0 - uv2, duv12 = uv1 - uv2;
    V3 dp02 = p0 - p2, dp12 = p1 - p2;
    float uv_det = duv02.x * duv12.y - duv02.y * duv12.x;
    V3 dpdu, dpdv;
    if (uv_det == 0.0f) {
        V3 n = normalized(cross(p2 - p0, p1 - p0));
        coordinate_system(n, &dpdu, &dpdv);
    } else {
        float inv_uv_det = 1.0f / uv_det;
        dpdu = (dp02 * duv12.y - dp12 * duv02.y) * inv_uv_det;
        dpdv = ((-dp02) * duv12.x + dp12 * duv02.x) * inv_uv_det;
    }
    // :174-185
    V3 p_hit = p0 * b0 + p1 * b1 + p2 * b2;
    V2 uv_hit = uv0 * b0 + uv1 * b1 + uv2 * b2;
    si->p = p_hit;
    si->wo = -ray.d;
    si->uv = uv_hit;
    si->dpdu = dpdu; si->dpdv = dpdv;
    si->sh_dpdu = dpdu; si->sh_dpdv = dpdv;
    si->area_light = tri.area_light;
    // :187-194 (overrides the cross(dpdu,dpdv) normal of SurfaceInteraction::new)
    V3 n = normalized(cross(dp02, dp12));
    if (mesh.transform_swaps_handedness) n = -n;
    si->n = n;
    si->sh_n = n;
    // :197-224 shading normals
    if (!mesh.normals.empty()) {
        V3 n0 = mesh.normals[tri.v[0]], n1 = mesh.normals[tri.v[1]], n2 = mesh.normals[tri.v[2]];
        V3 ns = normalized(n0 * b0 + n1 * b1 + n2 * b2);
        if (len_sqr(ns) > 0.0f) ns = normalized(ns);
        else ns = si->n;
        V3 ss = normalized(si->dpdu);
        V3 ts = cross(ss, ns);
        if (len_sqr(ts) > 0.0f) {
            ts = normalized(ts);
            ss = cross(ts, ns);
        } else {
            coordinate_system(ns, &ss, &ts);
        }
        // SurfaceInteraction::set_shading_geometry, interaction.rs:126-132
        si->sh_n = normalized(cross(ss, ts));
        si->n = faceforward_n(si->n, si->sh_n);
        si->sh_dpdu = ss;
        si->sh_dpdv = ts;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------
// BVH build, bvh.rs:39-115, 305-523
namespace bvh_detail {
struct PrimInfo {
    uint32_t shape_index;
    Bounds3 bounds;
    V3 centroid;
};
struct BuildNode {
    Bounds3 bounds;
    int32_t child0 = -1, child1 = -1;  // indices into the build-node pool
    int split_axis = 0;
    uint32_t first_shape = 0, shape_count = 0;
};
// itertools 0.10 `partition` (third-party, yuki/Cargo.toml:21): front scan; on a failing front
// element scan from the back for a passing one and swap. Returns the split index.
template <class Pred>
inline size_t itertools_partition(PrimInfo* a, size_t n, Pred pred) {
    size_t split = 0, front = 0, back = n;  // the double-ended iterator is [front, back)
    while (front < back) {                   // iter.next()
        PrimInfo* f = &a[front++];
        if (!pred(*f)) {
            for (;;) {
                if (front >= back) return split;  // iter.next_back() == None => break 'main
                PrimInfo* b = &a[--back];
                if (pred(*b)) {
                    std::swap(*f, *b);
                    break;
                }
            }
        }
        split += 1;
    }
    return split;
}
}  // namespace bvh_detail

inline bool Scene::build_bvh() {
    using namespace bvh_detail;
    const size_t n = shapes.size();
    if (n == 0) return false;
    std::vector<PrimInfo> info(n);
    for (size_t i = 0; i < n; ++i) {
        Bounds3 b = world_bound(shapes[i]);
        // bvh.rs:56 — "centroid" = p_min + diagonal / 0.5 (reference quirk, reproduced)
        info[i] = {(uint32_t)i, b, b.p_min + (diagonal(b) / 0.5f)};
    }
    std::vector<BuildNode> pool;
    pool.reserve(2 * n);
    const uint32_t max_in_node = max_shapes_in_node;
    const SplitMethod method = split_method;
    bool ok = true;

    // bvh.rs:305-390. Returns the pool index of the subtree root.
    struct Rec {
        std::vector<PrimInfo>& info;
        std::vector<BuildNode>& pool;
        uint32_t max_in_node;
        SplitMethod method;
        bool& ok;

        size_t split_equal_counts(size_t start, size_t end, int axis) {  // bvh.rs:422-436
            size_t mid = (start + end) / 2;
            // select_nth_unstable_by: resulting order inside each half is unspecified in the reference
            // (rustc-version dependent); the two SETS are what parity can rely on (distinct keys).
            std::nth_element(info.begin() + start, info.begin() + mid, info.begin() + end,
                             [axis](const PrimInfo& a, const PrimInfo& b) { return a.centroid[axis] < b.centroid[axis]; });
            return mid;
        }
        size_t split_middle(const Bounds3& cb, size_t start, size_t end, int axis) {  // bvh.rs:438-450
            float mid_value = (cb.p_min[axis] + cb.p_max[axis]) / 2.0f;
            return itertools_partition(&info[start], end - start,
                                       [=](const PrimInfo& s) { return s.centroid[axis] < mid_value; }) + start;
        }
        size_t split_sah(const Bounds3& bounds, const Bounds3& cb, size_t start, size_t end, int axis) {  // :452-523
            size_t count = end - start;
            if (count <= 2) return start;
            constexpr int NB = 12;
            struct Bucket { size_t count = 0; Bounds3 bounds = bounds_default(); } buckets[NB];
            auto bucket_of = [&](const PrimInfo& s) {
                float bf = (float)NB * offset(cb, s.centroid)[axis];
                float cl = fmax_(bf, 0.0f);
                // `as usize` saturates; NaN -> 0
                size_t b = cl != cl ? 0 : (cl >= 1.8446744e19f ? SIZE_MAX : (size_t)cl);
                return std::min<size_t>(b, NB - 1);
            };
            for (size_t i = start; i < end; ++i) {
                size_t b = bucket_of(info[i]);
                buckets[b].count += 1;
                buckets[b].bounds = union_b(buckets[b].bounds, info[i].bounds);
            }
            float costs[NB - 1];
            for (int i = 0; i < NB - 1; ++i) {
                Bounds3 b0 = bounds_default(), b1 = bounds_default();
                size_t c0 = 0, c1 = 0;
                for (int j = 0; j <= i; ++j) { b0 = union_b(b0, buckets[j].bounds); c0 += buckets[j].count; }
                for (int j = i + 1; j < NB; ++j) { b1 = union_b(b1, buckets[j].bounds); c1 += buckets[j].count; }
                costs[i] = 1.0f + ((float)c0 * surface_area(b0) + (float)c1 * surface_area(b1)) /
                                      fmax_(surface_area(bounds), 1e-10f);
            }
            int best = 0;  // Iterator::min_by keeps the first minimum
            for (int i = 1; i < NB - 1; ++i)
                if (costs[i] < costs[best]) best = i;
            float min_cost = costs[best];
            float leaf_cost = (float)count;
            if (min_cost < leaf_cost) {
                return itertools_partition(&info[start], end - start,
                                           [&](const PrimInfo& s) { return bucket_of(s) <= (size_t)best; }) + start;
            }
            return SIZE_MAX;
        }
        int32_t leaf(size_t start, size_t end, const Bounds3& bounds) {
            BuildNode nd;
            nd.bounds = bounds;
            nd.first_shape = (uint32_t)start;  // == ordered_shapes.len() at this point of the DFS
            nd.shape_count = (uint32_t)(end - start);
            pool.push_back(nd);
            return (int32_t)pool.size() - 1;
        }
        int32_t build(size_t start, size_t end) {
            Bounds3 bounds = bounds_default();
            for (size_t i = start; i < end; ++i) bounds = union_b(bounds, info[i].bounds);
            size_t count = end - start;
            if (count <= max_in_node) return leaf(start, end, bounds);
            Bounds3 cb = bounds_default();
            for (size_t i = start; i < end; ++i) cb = union_p(cb, info[i].centroid);
            int axis = maximum_extent(cb);
            if (cb.p_max[axis] == cb.p_min[axis]) return leaf(start, end, bounds);
            size_t mid;
            switch (method) {
                case SPLIT_SAH:
                    mid = split_sah(bounds, cb, start, end, axis);
                    if (!(mid != start && mid != end)) mid = split_equal_counts(start, end, axis);
                    break;
                case SPLIT_MIDDLE:
                    mid = split_middle(cb, start, end, axis);
                    if (!(mid != start && mid != end)) mid = split_equal_counts(start, end, axis);
                    break;
                default:
                    mid = split_equal_counts(start, end, axis);
            }
            if (mid == start) { ok = false; return leaf(start, end, bounds); }  // assert_ne!(mid, start)
            if (mid == SIZE_MAX) return leaf(start, end, bounds);
            int32_t c0 = build(start, mid);
            int32_t c1 = build(mid, end);
            BuildNode nd;
            nd.bounds = union_b(pool[c0].bounds, pool[c1].bounds);
            nd.child0 = c0; nd.child1 = c1; nd.split_axis = axis;
            pool.push_back(nd);
            return (int32_t)pool.size() - 1;
        }
    } rec{info, pool, max_in_node, method, ok};

    int32_t root = rec.build(0, n);
    if (!ok) return false;

    // ordered shapes == final PrimInfo order (leaves are emitted left to right over [start, end))
    std::vector<Triangle> ordered(n);
    for (size_t i = 0; i < n; ++i) ordered[i] = shapes[info[i].shape_index];
    shapes.swap(ordered);

    // flatten_tree, bvh.rs:396-419 (pre-order, first child at self+1) — iterative to spare the stack
    nodes.assign(pool.size(), BVHNode{});
    struct Item { int32_t bn; uint32_t parent; };  // parent: index of the node whose `offset` awaits this child, or ~0u
    std::vector<Item> stack{{root, ~0u}};
    uint32_t next = 0;
    while (!stack.empty()) {
        Item it = stack.back();
        stack.pop_back();
        const BuildNode& bn = pool[it.bn];
        uint32_t self = next++;
        if (it.parent != ~0u) nodes[it.parent].offset = self;
        BVHNode& out = nodes[self];
        for (int k = 0; k < 3; ++k) { out.p_min[k] = bn.bounds.p_min[k]; out.p_max[k] = bn.bounds.p_max[k]; }
        if (bn.child0 >= 0) {
            out.is_leaf = 0; out.split_axis = (uint8_t)bn.split_axis; out.shape_count = 0;
            stack.push_back({bn.child1, self});  // second child: its index is patched when it is emitted
            stack.push_back({bn.child0, ~0u});   // first child comes right after self
        } else {
            out.is_leaf = 1; out.split_axis = 0;
            out.offset = bn.first_shape;
            out.shape_count = (uint16_t)bn.shape_count;  // bvh.rs:546 stores u16
        }
    }
    return true;
}

// bvh.rs:160-232
inline IntersectionResult Scene::intersect(Ray ray, TraversalStats* ts) const {
    IntersectionResult res;
    V3 inv_dir = {1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
    bool dir_is_neg[3] = {inv_dir.x < 0.0f, inv_dir.y < 0.0f, inv_dir.z < 0.0f};
    size_t current = 0, to_visit = 0;
    size_t stack[64];
    uint64_t tri_tests = 0;
    for (;;) {
        const BVHNode& node = nodes[current];
        res.intersection_test_count += 1;
        Bounds3 b{{node.p_min[0], node.p_min[1], node.p_min[2]}, {node.p_max[0], node.p_max[1], node.p_max[2]}};
        if (bounds_intersect(b, ray, inv_dir)) {
            res.intersection_count += 1;
            if (!node.is_leaf) {
                if (dir_is_neg[node.split_axis]) {
                    stack[to_visit++] = current + 1;
                    current = node.offset;
                } else {
                    stack[to_visit++] = node.offset;
                    current += 1;
                }
            } else {
                for (uint32_t s = node.offset; s < node.offset + node.shape_count; ++s) {
                    float t;
                    SurfaceInteraction si;
                    tri_tests += 1;
                    if (triangle_intersect(shapes[s], ray, &t, &si, true)) {
                        res.has_hit = true;
                        res.hit = {t, si, s};
                        ray.t_max = t;
                    }
                }
                if (to_visit == 0) break;
                current = stack[--to_visit];
            }
        } else {
            if (to_visit == 0) break;
            current = stack[--to_visit];
        }
    }
    if (ts) { ts->closest_nodes += res.intersection_test_count; ts->closest_tris += tri_tests; }
    return res;
}

// bvh.rs:235-302
inline bool Scene::any_intersect(const Ray& ray, int32_t area_light, TraversalStats* ts) const {
    V3 inv_dir = {1.0f / ray.d.x, 1.0f / ray.d.y, 1.0f / ray.d.z};
    size_t current = 0, to_visit = 0;
    size_t stack[64];
    uint64_t node_tests = 0, tri_tests = 0;
    bool result = false;
    for (;;) {
        const BVHNode& node = nodes[current];
        node_tests += 1;
        Bounds3 b{{node.p_min[0], node.p_min[1], node.p_min[2]}, {node.p_max[0], node.p_max[1], node.p_max[2]}};
        if (bounds_intersect(b, ray, inv_dir)) {
            if (!node.is_leaf) {
                if (inv_dir[node.split_axis] < 0.0f) {
                    stack[to_visit++] = current + 1;
                    current = node.offset;
                } else {
                    stack[to_visit++] = node.offset;
                    current += 1;
                }
            } else {
                for (uint32_t s = node.offset; s < node.offset + node.shape_count; ++s) {
                    float t;
                    tri_tests += 1;
                    if (triangle_intersect(shapes[s], ray, &t, nullptr, false)) {
                        // :269-280 — a hit on the target light's own emissive geometry does not occlude
                        if (area_light >= 0 && shapes[s].area_light >= 0) {
                            if (shapes[s].area_light != area_light) { result = true; goto done; }
                        } else {
                            result = true;
                            goto done;
                        }
                    }
                }
                if (to_visit == 0) break;
                current = stack[--to_visit];
            }
        } else {
            if (to_visit == 0) break;
            current = stack[--to_visit];
        }
    }
done:
    if (ts) { ts->any_nodes += node_tests; ts->any_tris += tri_tests; }
    return result;
}

// ------------------------------------------------------------------------------------------------
// interaction.rs:27-59
inline Ray spawn_ray(V3 p, V3 n, V3 d) {
    V3 off = n * 0.001f;
    V3 o = dot(d, n) > 0.0f ? p + off : p - off;
    return {o, d, INFINITY};
}
inline Ray spawn_ray_to(V3 p, V3 n, V3 other_p) {
    V3 off = n * 0.001f;
    V3 o = dot(other_p - p, n) > 0.0f ? p + off : p - off;
    return {o, other_p - o, 0.9999f};  // direction NOT normalised
}

// ------------------------------------------------------------------------------------------------
// Textures
inline Spec texture_eval(const Texture& tex, const SurfaceInteraction& si) {
    if (tex.kind == TEX_CONSTANT) return tex.value;
    // image_texture.rs:81-111 — repeat, flip y, nearest texel, no filtering, no gamma
    float sx = si.uv.x, sy = si.uv.y;
    sx = sx - std::trunc(sx);  // f32::fract
    if (sx < 0.0f) sx = 1.0f + sx;
    sy = sy - std::trunc(sy);
    if (sy < 0.0f) sy = 1.0f + sy;
    sy = 1.0f - sy;
    sx = sx * (float)tex.width - 0.5f;
    sy = sy * (float)tex.height - 0.5f;
    // `as usize` saturates at 0 for negatives / NaN
    size_t ix = sx > 0.0f ? (size_t)sx : 0, iy = sy > 0.0f ? (size_t)sy : 0;
    size_t idx = iy * tex.width + ix;  // the reference would panic out of bounds; uv==exact edge cases stay in range
    const float* t = &tex.texels[idx * 3];
    return {t[0], t[1], t[2]};
}

// ------------------------------------------------------------------------------------------------
// Lights
struct LightSample {
    V3 l;
    Spec li;
    bool has_vis;
    Ray vis_ray;
    int32_t vis_area_light;  // light index whose own geometry must not occlude, or -1
    float pdf;
};

inline LightSample sample_li(const Scene& scene, int32_t light_index, const SurfaceInteraction& si, V2 u) {
    const Light& L = scene.lights[light_index];
    LightSample s{};
    s.vis_area_light = -1;
    switch (L.kind) {
        case LIGHT_POINT: {  // point_light.rs:27-49
            V3 to_light = L.p - si.p;
            float dist_sqr = len_sqr(to_light);
            s.li = L.i / dist_sqr;
            float dist = std::sqrt(dist_sqr);
            s.l = to_light / dist;
            s.has_vis = true;
            s.vis_ray = spawn_ray_to(si.p, si.n, L.p);
            s.pdf = 1.0f;
        } break;
        case LIGHT_SPOT: {  // spot_light.rs:38-80
            V3 to_light = L.p - si.p;
            float dist_sqr = len_sqr(to_light);
            float dist = std::sqrt(dist_sqr);
            s.l = to_light / dist;
            // falloff(), :38-50
            V3 dir_local = normalized(xf_vec(L.world_to_light, -s.l));
            float cos_theta = dir_local.z, fall;
            if (cos_theta < L.cos_total_width) fall = 0.0f;
            else if (cos_theta > L.cos_falloff_start) fall = 1.0f;
            else {
                float delta = (cos_theta - L.cos_total_width) / (L.cos_falloff_start - L.cos_total_width);
                fall = (delta * delta) * (delta * delta);
            }
            s.li = L.i * fall / dist_sqr;
            s.has_vis = !is_black(s.li);
            if (s.has_vis) s.vis_ray = spawn_ray_to(si.p, si.n, L.p);
            s.pdf = 1.0f;
        } break;
        case LIGHT_RECT: {  // rectangular_light.rs:46-72
            V3 p = xf_point(L.sample_to_world, v3(u.x, 0.0f, u.y));
            V3 n = xf_normal(L.sample_to_world, v3(0.0f, -1.0f, 0.0f));  // not normalised
            V3 wi = normalized(p - si.p);
            s.li = dot_nv(n, -wi) > 0.0f ? L.i : spec1(0.0f);
            s.l = wi;
            s.has_vis = true;
            s.vis_ray = spawn_ray_to(si.p, si.n, p);
            s.vis_area_light = light_index;
            s.pdf = len_sqr(si.p - p) / (std::fabs(dot_nv(n, -wi)) * L.area);
        } break;
        case LIGHT_DISTANT: {  // distant_light.rs:24-43
            s.li = L.i;
            s.l = L.p;
            s.has_vis = true;
            s.vis_ray = spawn_ray_to(si.p, si.n, si.p + L.p * 10000.0f);
            s.pdf = 1.0f;
        } break;
    }
    return s;
}

// interaction.rs:134-138 + rectangular_light.rs:74-81
inline Spec emitted_radiance(const Scene& scene, const SurfaceInteraction& si, V3 w) {
    if (si.area_light < 0) return spec1(0.0f);
    return dot_nv(si.n, w) > 0.0f ? scene.lights[si.area_light].i : spec1(0.0f);
}

// ------------------------------------------------------------------------------------------------
// Camera (camera.rs)
enum FovAxis : uint32_t { FOV_X = 0, FOV_Y = 1 };
struct Camera {
    Transform camera_to_world, raster_to_camera;
};
// camera.rs:52-102
inline bool camera_new(V3 pos, V3 target, V3 up, FovAxis axis, float fov_deg, uint32_t res_x, uint32_t res_y,
                       Camera* cam) {
    Transform w2c;
    if (!xf_look_at(pos, target, up, &w2c)) return false;
    cam->camera_to_world = xf_inverted(w2c);
    const float near_ = 1e-2f, far_ = 1000.0f;
    float inv_tan = 1.0f / std::tan(to_radians(fov_deg) / 2.0f);
    M44 persp{};
    persp.m[0][0] = 1.0f; persp.m[1][1] = 1.0f;
    persp.m[2][2] = far_ / (far_ - near_);
    persp.m[2][3] = -(far_ * near_) / (far_ - near_);
    persp.m[3][2] = 1.0f;
    Transform persp_t;
    if (!xf_new(persp, &persp_t)) return false;
    Transform camera_to_screen = xf_mul(xf_scale(inv_tan, inv_tan, 1.0f), persp_t);
    float film_x = (float)res_x, film_y = (float)res_y;
    V2 smin, smax;
    if (axis == FOV_X) {
        float ar = film_x / film_y;
        smin = {-1.0f, -1.0f / ar}; smax = {1.0f, 1.0f / ar};
    } else {
        float ar = film_y / film_x;
        smin = {-1.0f / ar, -1.0f}; smax = {1.0f / ar, 1.0f};
    }
    Transform screen_to_raster =
        xf_mul(xf_scale(film_x, film_y, 1.0f), xf_mul(xf_scale(1.0f / (smax.x - smin.x), 1.0f / (smin.y - smax.y), 1.0f),
                                                        xf_translation(v3(-smin.x, -smax.y, 0.0f))));
    Transform raster_to_screen = xf_inverted(screen_to_raster);
    cam->raster_to_camera = xf_mul(xf_inverted(camera_to_screen), raster_to_screen);
    return true;
}
// camera.rs:105-114
inline Ray camera_ray(const Camera& cam, V2 p_film) {
    V3 p_camera = xf_point(cam.raster_to_camera, v3(p_film.x, p_film.y, 0.0f));
    Ray r{v3(0.0f, 0.0f, 0.0f), normalized(p_camera), INFINITY};
    return xf_ray(cam.camera_to_world, r);
}

}  // namespace yko
