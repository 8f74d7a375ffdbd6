// Host-side scene assembly for the B200 backend and the Level-2 entry points of include/yuki_gpu.h.
// Mirrors what yuki's loaders do before the hot path starts: Mesh::new (shapes/mesh.rs:21-43),
// Triangle::new (shapes/triangle.rs:27-46), the light constructors (lights/*.rs), Camera::new
// (camera.rs:52-102), film_tiles (film.rs:299-376) — then flattens everything into the leaf-ordered SoA
// arrays yk_scene_create uploads.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "host_math.h"
#include "yk_fastdiv.h"
#include "yuki_gpu.h"
#include "yk_guard.h"

namespace ykh {
int bvh_build(const float* tri_vertices, uint32_t n_tris, uint32_t max_shapes_in_node, uint32_t split_method,
              std::vector<yk_bvh_node>* nodes, std::vector<uint32_t>* order, const char** why);
int bvh_build_boxes(const float* boxes6, uint32_t n, uint32_t max_shapes_in_node, uint32_t split_method, std::vector<yk_bvh_node>* nodes,
                    std::vector<uint32_t>* order, const char** why);
}

// ---- error channel (shared with the device side) ------------------------------------------------
static thread_local std::string g_last_error;
extern "C" const char* yk_last_error(void) { return g_last_error.c_str(); }
int yk_set_error(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}

using namespace ykh;

static xform to_xform(const yk_transform& t) {
    xform x;
    std::memcpy(x.m.e, t.m, 64);
    std::memcpy(x.inv.e, t.m_inv, 64);
    return x;
}
static void from_xform(const xform& x, yk_transform* t) {
    std::memcpy(t->m, x.m.e, 64);
    std::memcpy(t->m_inv, x.inv.e, 64);
}

struct yk_host_scene {
    std::vector<yk_bvh_node> nodes;
    std::vector<float> tri_vertices, tri_normals, tri_uvs;
    std::vector<uint32_t> tri_orig_id, tri_material;
    std::vector<int32_t> tri_area_light;
    std::vector<uint8_t> tri_flags;
    std::vector<yk_texture_desc> textures;
    std::vector<std::vector<float>> texel_storage;
    std::vector<yk_material_desc> materials;
    std::vector<yk_light> lights;
    float background[3];
    std::vector<yk_sphere> spheres;
    std::vector<int32_t> tri_sphere;  // empty when the scene has no sphere
};

extern "C" {

int yk_light_make(const yk_light_desc* d, yk_light* out) {
    return yk_guard("yk_light_make", [&]() -> int {
    if (!d || !out) return yk_set_error(YK_ERR_INVALID, "yk_light_make: null argument");
    std::memset(out, 0, sizeof(*out));
    out->kind = d->kind;
    std::memcpy(out->i, d->intensity, 12);
    const xform l2w = to_xform(d->light_to_world);
    switch (d->kind) {
        case YK_LIGHT_POINT:  // PointLight::new, point_light.rs:18-24
            store3(apply_point(l2w.m, mk3(0, 0, 0)), out->p);
            break;
        case YK_LIGHT_SPOT: {  // SpotLight::new, spot_light.rs:23-36
            store3(apply_point(l2w.m, mk3(0, 0, 0)), out->p);
            std::memcpy(out->world_to_light, l2w.inv.e, 64);
            out->cos_total_width = cosf(deg2rad(d->total_width_deg));
            out->cos_falloff_start = cosf(deg2rad(d->falloff_start_deg));
        } break;
        case YK_LIGHT_RECT: {  // RectangularLight::new, rectangular_light.rs:27-43
            const xform s2l = xf_compose(xf_scaling(d->size[0], 1.0f, d->size[1]), xf_translate(mk3(-0.5f, 0.0f, -0.5f)));
            const xform s2w = xf_compose(l2w, s2l);
            std::memcpy(out->sample_to_world, s2w.m.e, 64);
            std::memcpy(out->sample_to_world_inv, s2w.inv.e, 64);
            out->area = d->size[0] * d->size[1];
        } break;
        case YK_LIGHT_DISTANT:  // DistantLight::new, distant_light.rs:17-21
            std::memcpy(out->p, d->direction, 12);
            break;
        default:
            return yk_set_error(YK_ERR_INVALID, "yk_light_make: unknown light kind");
    }
    return YK_OK;
    });
}

int yk_bvh_build(const float* tri_vertices, uint32_t n_tris, uint32_t max_shapes_in_node, uint32_t split_method,
                 yk_bvh_node* nodes, uint32_t* n_nodes, uint32_t* order) {
    return yk_guard("yk_bvh_build", [&]() -> int {
    if (!nodes || !n_nodes || !order) return yk_set_error(YK_ERR_INVALID, "yk_bvh_build: null output");
    std::vector<yk_bvh_node> nv;
    std::vector<uint32_t> ov;
    const char* why = "";
    int rc = bvh_build(tri_vertices, n_tris, max_shapes_in_node, split_method, &nv, &ov, &why);
    if (rc != YK_OK) return yk_set_error(rc, why);
    std::memcpy(nodes, nv.data(), nv.size() * sizeof(yk_bvh_node));
    std::memcpy(order, ov.data(), ov.size() * sizeof(uint32_t));
    *n_nodes = (uint32_t)nv.size();
    return YK_OK;
    });
}

int yk_host_scene_build(const yk_host_scene_desc* d, yk_host_scene** out) {
    return yk_guard("yk_host_scene_build", [&]() -> int {
    if (!d || !out) return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: null argument");
    auto hs = std::make_unique<yk_host_scene>();
    const bool timing = getenv("YK_SCENE_TIMING") != nullptr;  // development: phase times to stderr
    const auto t_start = std::chrono::steady_clock::now();
    auto lap = [&](const char* what) {
        if (timing) fprintf(stderr, "yk_host_scene_build: %-24s at %8.1f ms\n", what,
                            1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count());
    };

    // Shapes in declaration order (the reference's `shapes` Vec before the BVH reorders it). The list is sized first (one shape
    // per index triplet, one per sphere), then filled by several threads — a 10 M-triangle mesh spent 3.4 s here appending
    // vertex by vertex.
    bool any_normals = false, any_uvs = false;
    for (uint32_t mi = 0; mi < d->n_meshes; ++mi) {
        any_normals |= d->meshes[mi].normals != nullptr;
        any_uvs |= d->meshes[mi].uvs != nullptr;
    }
    if (d->n_spheres && !d->spheres) return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: null sphere list");
    if (d->n_meshes && !d->meshes) return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: null mesh list");
    struct Object { int32_t id; size_t first; };  // id >= 0: mesh, < 0: sphere -1 - id; first = its first shape slot
    std::vector<Object> objects;
    size_t n_shapes = 0;
    auto add_object = [&](int32_t o) -> int {
        if (o >= 0 ? (uint32_t)o >= d->n_meshes : (uint32_t)(-1 - o) >= d->n_spheres)
            return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: object index out of range");
        objects.push_back(Object{o, n_shapes});
        n_shapes += o >= 0 ? d->meshes[o].n_indices / 3 : 1;
        return YK_OK;
    };
    int rc = YK_OK;
    if (d->objects) {  // the loader's declaration order (pbrt/mod.rs:797-809)
        for (uint32_t i = 0; i < d->n_objects; ++i)
            if ((rc = add_object(d->objects[i])) != YK_OK) return rc;
    } else {  // meshes, then spheres (scene/mod.rs:497)
        for (uint32_t mi = 0; mi < d->n_meshes; ++mi) add_object((int32_t)mi);
        for (uint32_t k = 0; k < d->n_spheres; ++k) add_object(-1 - (int32_t)k);
    }
    if (n_shapes > 0xffffffffull) return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: more than 2^32 shapes");
    // (uninitialised storage: every slot is written below, by the thread that first touches its pages)
    auto raw = [](size_t n, auto tag) { return std::unique_ptr<decltype(tag)[]>(new decltype(tag)[std::max<size_t>(n, 1)]); };
    const auto verts_p = raw(n_shapes * 9, 0.0f), norms_p = raw(any_normals ? n_shapes * 9 : 0, 0.0f), uvs_p = raw(any_uvs ? n_shapes * 6 : 0, 0.0f);
    const auto boxes_p = raw(n_shapes * 6, 0.0f);      // world bounds per shape, in shape-list order
    const auto mats_p = raw(n_shapes, uint32_t(0));
    const auto alights_p = raw(n_shapes, int32_t(0));
    const auto flags_p = raw(n_shapes, uint8_t(0));
    float *verts = verts_p.get(), *norms = norms_p.get(), *uvs = uvs_p.get(), *boxes = boxes_p.get();
    uint32_t* mats = mats_p.get();
    int32_t* alights = alights_p.get();
    uint8_t* flags = flags_p.get();
    std::vector<int32_t> sphere_of(d->n_spheres ? n_shapes : 0, -1);  // sphere index per shape, -1 for triangles (scenes with spheres only)
    const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    auto parallel_for = [hw](size_t n, size_t grain, auto&& f) {  // f(begin, end) over [0, n) on up to hw threads
        const size_t w = std::max<size_t>(1, std::min<size_t>(hw, n / std::max<size_t>(grain, 1)));
        std::vector<std::thread> th;
        for (size_t c = 1; c < w; ++c) th.emplace_back([&f, c, w, n] { f(n * c / w, n * (c + 1) / w); });
        f((size_t)0, n / w);
        for (auto& t : th) t.join();
    };
    // Mesh::new + Triangle::new: one shape per index triplet (mesh.rs:21-43, triangle.rs:229-235 for the bounds)
    auto add_mesh = [&](uint32_t mi, size_t first) -> int {
        const yk_mesh_desc& m = d->meshes[mi];
        if (m.material < 0 || (uint32_t)m.material >= d->n_materials)
            return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: mesh material index out of range");
        if (m.area_light >= (int32_t)d->n_lights || (m.area_light >= 0 && d->lights[m.area_light].kind != YK_LIGHT_RECT))
            return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: mesh area_light must index a rectangular light");
        const xform o2w = to_xform(m.object_to_world);
        std::vector<f3> wp(m.n_points), wn(m.normals ? m.n_points : 0);
        parallel_for(m.n_points, 1u << 16, [&](size_t b, size_t e) {
            for (size_t k = b; k < e; ++k) wp[k] = apply_point(o2w.m, load3(m.points + 3 * k));  // mesh.rs:27-29
            if (m.normals)
                for (size_t k = b; k < e; ++k) wn[k] = apply_normal(o2w.inv, load3(m.normals + 3 * k));  // :31-33
        });
        const uint8_t fl = (flips_handedness(o2w.m) ? YK_TRI_SWAPS_HANDEDNESS : 0u) | (m.normals ? YK_TRI_HAS_NORMALS : 0u) |
                           (m.uvs ? YK_TRI_HAS_UVS : 0u);
        std::atomic<bool> bad_index{false};
        parallel_for(m.n_indices / 3, 1u << 15, [&](size_t b, size_t e) {
            for (size_t t = b; t < e; ++t) {
                const size_t slot = first + t;
                float* v = &verts[slot * 9];
                for (int c = 0; c < 3; ++c) {
                    const uint32_t vi = m.indices[3 * t + c];
                    if (vi >= m.n_points) { bad_index = true; return; }
                    store3(wp[vi], v + 3 * c);
                    if (any_normals) {
                        float* nn = &norms[slot * 9 + 3 * c];
                        if (m.normals) store3(wn[vi], nn);
                        else nn[0] = nn[1] = nn[2] = 0.0f;
                    }
                    if (any_uvs) {
                        uvs[slot * 6 + 2 * c] = m.uvs ? m.uvs[2 * vi] : 0.0f;
                        uvs[slot * 6 + 2 * c + 1] = m.uvs ? m.uvs[2 * vi + 1] : 0.0f;
                    }
                }
                box3 bx{min3(load3(v), load3(v + 3)), max3(load3(v), load3(v + 3))};
                bx = grow(bx, load3(v + 6));
                store3(bx.lo, &boxes[slot * 6]);
                store3(bx.hi, &boxes[slot * 6 + 3]);
                mats[slot] = (uint32_t)m.material;
                alights[slot] = m.area_light;
                flags[slot] = fl;
            }
        });
        if (bad_index) return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: vertex index out of range");
        return YK_OK;
    };
    // Sphere::new (sphere.rs:23-33); bounds = object_to_world * [-r, r]^3 as the union of the eight transformed corners
    // (sphere.rs:121-123, math/transform.rs:186-201)
    auto add_sphere = [&](uint32_t k, size_t slot) -> int {
        const yk_sphere_desc& sd = d->spheres[k];
        if (sd.material < 0 || (uint32_t)sd.material >= d->n_materials)
            return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: sphere material index out of range");
        const xform o2w = to_xform(sd.object_to_world);
        yk_sphere sp{};
        std::memcpy(sp.object_to_world, o2w.m.e, 64);
        std::memcpy(sp.world_to_object, o2w.inv.e, 64);
        sp.radius = sd.radius;
        sp.swaps_handedness = flips_handedness(o2w.m) ? 1u : 0u;
        hs->spheres.push_back(sp);
        const float r = sd.radius;
        const f3 mi = mk3(-r, -r, -r), ma = mk3(r, r, r);
        const f3 corners[8] = {mi, mk3(ma.x, mi.y, mi.z), mk3(mi.x, ma.y, mi.z), mk3(mi.x, mi.y, ma.z),
                               mk3(ma.x, ma.y, mi.z), mk3(ma.x, mi.y, ma.z), mk3(mi.x, ma.y, ma.z), ma};
        box3 b{mk3(3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f), mk3(-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f)};
        for (const f3& cnr : corners) b = grow(b, apply_point(o2w.m, cnr));
        store3(b.lo, &boxes[slot * 6]);
        store3(b.hi, &boxes[slot * 6 + 3]);
        // the shape slot: no vertices (zeros), the sphere's material, no area light (Sphere::new takes none)
        std::fill(verts + slot * 9, verts + slot * 9 + 9, 0.0f);
        if (any_normals) std::fill(norms + slot * 9, norms + slot * 9 + 9, 0.0f);
        if (any_uvs) std::fill(uvs + slot * 6, uvs + slot * 6 + 6, 0.0f);
        mats[slot] = (uint32_t)sd.material;
        alights[slot] = -1;
        flags[slot] = (uint8_t)(YK_TRI_IS_SPHERE | (sp.swaps_handedness ? YK_TRI_SWAPS_HANDEDNESS : 0u));
        sphere_of[slot] = (int32_t)hs->spheres.size() - 1;
        return YK_OK;
    };
    for (const Object& ob : objects)
        if ((rc = ob.id >= 0 ? add_mesh((uint32_t)ob.id, ob.first) : add_sphere((uint32_t)(-1 - ob.id), ob.first)) != YK_OK) return rc;
    const uint32_t n_tris = (uint32_t)n_shapes;
    lap("shapes flattened");
    std::vector<uint32_t> order;
    const char* why = "";
    rc = bvh_build_boxes(boxes, n_tris, d->max_shapes_in_node ? d->max_shapes_in_node : 1u, d->split_method, &hs->nodes,
                             &order, &why);
    if (rc != YK_OK) return yk_set_error(rc, why);
    lap("BVH built");

    // Gather into leaf order.
    hs->tri_vertices.resize((size_t)n_tris * 9);
    if (any_normals) hs->tri_normals.resize((size_t)n_tris * 9);
    if (any_uvs) hs->tri_uvs.resize((size_t)n_tris * 6);
    hs->tri_orig_id = order;
    hs->tri_material.resize(n_tris);
    hs->tri_area_light.resize(n_tris);
    hs->tri_flags.resize(n_tris);
    parallel_for(n_tris, 1u << 16, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            const uint32_t s = order[i];
            std::memcpy(&hs->tri_vertices[i * 9], &verts[(size_t)s * 9], 36);
            if (any_normals) std::memcpy(&hs->tri_normals[i * 9], &norms[(size_t)s * 9], 36);
            if (any_uvs) std::memcpy(&hs->tri_uvs[i * 6], &uvs[(size_t)s * 6], 24);
            hs->tri_material[i] = mats[s];
            hs->tri_area_light[i] = alights[s];
            hs->tri_flags[i] = flags[s];
        }
    });
    if (!hs->spheres.empty()) {
        hs->tri_sphere.resize(n_tris);
        for (uint32_t i = 0; i < n_tris; ++i) hs->tri_sphere[i] = sphere_of[order[i]];
    }

    lap("gathered into leaf order");
    hs->texel_storage.resize(d->n_textures);
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        yk_texture_desc t = d->textures[i];
        if (t.kind == YK_TEX_IMAGE) {
            if (!t.texels || !t.width || !t.height) return yk_set_error(YK_ERR_INVALID, "yk_host_scene_build: empty image texture");
            hs->texel_storage[i].assign(t.texels, t.texels + (size_t)t.width * t.height * 3);
            t.texels = hs->texel_storage[i].data();
        } else {
            t.texels = nullptr;
        }
        hs->textures.push_back(t);
    }
    hs->materials.assign(d->materials, d->materials + d->n_materials);
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        yk_light l;
        rc = yk_light_make(&d->lights[i], &l);
        if (rc != YK_OK) return rc;
        hs->lights.push_back(l);
    }
    std::memcpy(hs->background, d->background, 12);
    *out = hs.release();
    return YK_OK;
    });
}

void yk_host_scene_destroy(yk_host_scene* hs) { delete hs; }

void yk_host_scene_flat(const yk_host_scene* hs, yk_scene_desc* o) {
    std::memset(o, 0, sizeof(*o));
    o->n_nodes = (uint32_t)hs->nodes.size();
    o->nodes = hs->nodes.data();
    o->n_tris = (uint32_t)hs->tri_material.size();
    o->tri_vertices = hs->tri_vertices.data();
    o->tri_normals = hs->tri_normals.empty() ? nullptr : hs->tri_normals.data();
    o->tri_uvs = hs->tri_uvs.empty() ? nullptr : hs->tri_uvs.data();
    o->tri_orig_id = hs->tri_orig_id.data();
    o->tri_material = hs->tri_material.data();
    o->tri_area_light = hs->tri_area_light.data();
    o->tri_flags = hs->tri_flags.data();
    o->n_spheres = (uint32_t)hs->spheres.size();
    o->spheres = hs->spheres.empty() ? nullptr : hs->spheres.data();
    o->tri_sphere = hs->tri_sphere.empty() ? nullptr : hs->tri_sphere.data();
    o->n_textures = (uint32_t)hs->textures.size();
    o->n_materials = (uint32_t)hs->materials.size();
    o->n_lights = (uint32_t)hs->lights.size();
    o->textures = hs->textures.data();
    o->materials = hs->materials.data();
    o->lights = hs->lights.data();
    std::memcpy(o->background, hs->background, 12);
}

// Camera::new, camera.rs:52-102
int yk_camera_make(const yk_camera_params* p, uint32_t res_x, uint32_t res_y, yk_camera* out) {
    return yk_guard("yk_camera_make", [&]() -> int {
    if (!p || !out || !res_x || !res_y) return yk_set_error(YK_ERR_INVALID, "yk_camera_make: bad argument");
    xform w2c;
    if (!xf_look_at(load3(p->position), load3(p->target), load3(p->up), &w2c))
        return yk_set_error(YK_ERR_SINGULAR, "Can't invert, singular matrix (look_at)");
    const xform c2w = xf_flip(w2c);
    const float z_near = 1e-2f, z_far = 1000.0f;
    const float inv_tan = 1.0f / tanf(deg2rad(p->fov_deg) / 2.0f);
    mat4 proj{};
    proj.at(0, 0) = 1.0f;
    proj.at(1, 1) = 1.0f;
    proj.at(2, 2) = z_far / (z_far - z_near);
    proj.at(2, 3) = -(z_far * z_near) / (z_far - z_near);
    proj.at(3, 2) = 1.0f;
    xform proj_t;
    if (!xf_from_matrix(proj, &proj_t)) return yk_set_error(YK_ERR_SINGULAR, "Can't invert, singular matrix (projection)");
    const xform cam_to_screen = xf_compose(xf_scaling(inv_tan, inv_tan, 1.0f), proj_t);
    const float fx = (float)res_x, fy = (float)res_y;
    float lo_x, lo_y, hi_x, hi_y;  // screen window, camera.rs:78-87
    if (p->fov_axis == YK_FOV_X) {
        const float ar = fx / fy;
        lo_x = -1.0f; lo_y = -1.0f / ar; hi_x = 1.0f; hi_y = 1.0f / ar;
    } else {
        const float ar = fy / fx;
        lo_x = -1.0f / ar; lo_y = -1.0f; hi_x = 1.0f / ar; hi_y = 1.0f;
    }
    const xform screen_to_raster =
        xf_compose(xf_scaling(fx, fy, 1.0f),
                   xf_compose(xf_scaling(1.0f / (hi_x - lo_x), 1.0f / (lo_y - hi_y), 1.0f), xf_translate(mk3(-lo_x, -hi_y, 0.0f))));
    const xform raster_to_cam = xf_compose(xf_flip(cam_to_screen), xf_flip(screen_to_raster));
    for (int i = 0; i < 16; ++i)  // Matrix4x4::new debug-asserts !has_nans (matrix.rs:23-27), e.g. target == position
        if (c2w.m.e[i] != c2w.m.e[i] || raster_to_cam.m.e[i] != raster_to_cam.m.e[i])
            return yk_set_error(YK_ERR_SINGULAR, "yk_camera_make: camera matrices contain NaN (degenerate look_at / fov)");
    std::memcpy(out->camera_to_world, c2w.m.e, 64);
    std::memcpy(out->raster_to_camera, raster_to_cam.m.e, 64);
    return YK_OK;
    });
}

// generate_tiles + outward_spiral, film.rs:299-376. Walks the square spiral around the centre tile and
// emits the in-range tiles; tile rectangles are computed on the fly instead of through a hash map.
uint32_t yk_film_tiles(uint32_t res_x, uint32_t res_y, uint32_t tile_dim, yk_tile* out, uint32_t cap) {
    if (!res_x || !res_y || !tile_dim) return 0;
    const int cols = (int)ceilf((float)res_x / (float)tile_dim);
    const int rows = (int)ceilf((float)res_y / (float)tile_dim);
    const int cx = cols / 2 - (1 - cols % 2), cy = rows / 2 - (1 - rows % 2);
    const int side = cols > rows ? cols : rows;
    int x = 0, y = 0, dx = 0, dy = -1;
    uint32_t n = 0;
    for (int step = 0; step < side * side; ++step) {
        const int tx = cx + x, ty = cy + y;
        if (tx >= 0 && tx < cols && ty >= 0 && ty < rows) {
            if (out && n < cap) {
                const uint32_t px = (uint32_t)tx * tile_dim, py = (uint32_t)ty * tile_dim;
                yk_tile t;
                t.x0 = (uint16_t)px;
                t.y0 = (uint16_t)py;
                t.x1 = (uint16_t)(px + tile_dim < res_x ? px + tile_dim : res_x);
                t.y1 = (uint16_t)(py + tile_dim < res_y ? py + tile_dim : res_y);
                t.sample = 0;
                t._pad = 0;
                t.index = (uint32_t)(ty * cols + tx);  // flat_index, film.rs:308-326
                out[n] = t;
            }
            ++n;
        }
        if (x == y || (x < 0 && x == -y) || (x > 0 && x == 1 - y)) {
            const int t = dx;
            dx = -dy;
            dy = t;
        }
        x += dx;
        y += dy;
    }
    return n;
}

void yk_xf_identity(yk_transform* o) { from_xform(xf_id(), o); }
void yk_xf_translation(const float* d, yk_transform* o) { from_xform(xf_translate(load3(d)), o); }
void yk_xf_scale(float x, float y, float z, yk_transform* o) { from_xform(xf_scaling(x, y, z), o); }
void yk_xf_rotation(float theta, const float* axis, yk_transform* o) { from_xform(xf_rotate(theta, load3(axis)), o); }
int yk_xf_new(const float* m16, yk_transform* o) {
    return yk_guard("yk_xf_new", [&]() -> int {
    mat4 m;
    std::memcpy(m.e, m16, 64);
    xform t;
    if (!xf_from_matrix(m, &t)) return yk_set_error(YK_ERR_SINGULAR, "Can't invert, singular matrix");
    from_xform(t, o);
    return YK_OK;
    });
}
int yk_xf_look_at(const float* pos, const float* target, const float* up, yk_transform* o) {
    return yk_guard("yk_xf_look_at", [&]() -> int {
    xform t;
    if (!xf_look_at(load3(pos), load3(target), load3(up), &t)) return yk_set_error(YK_ERR_SINGULAR, "Can't invert, singular matrix");
    from_xform(t, o);
    return YK_OK;
    });
}
void yk_xf_mul(const yk_transform* a, const yk_transform* b, yk_transform* o) { from_xform(xf_compose(to_xform(*a), to_xform(*b)), o); }
void yk_xf_inverted(const yk_transform* a, yk_transform* o) { from_xform(xf_flip(to_xform(*a)), o); }
void yk_xf_point(const yk_transform* t, const float* p, float* o) { store3(apply_point(to_xform(*t).m, load3(p)), o); }
void yk_xf_vec(const yk_transform* t, const float* v, float* o) { store3(apply_vec(to_xform(*t).m, load3(v)), o); }
void yk_xf_normal(const yk_transform* t, const float* n, float* o) { store3(apply_normal(to_xform(*t).inv, load3(n)), o); }

uint64_t yk_selftest_fastdiv(uint32_t d, const uint32_t* numerators, uint64_t count) {
    const FastDiv f = FastDiv::make(d);
    uint64_t bad = 0;
    for (uint64_t i = 0; i < count; ++i) bad += f.div_host(numerators[i]) != numerators[i] / f.d;
    return bad;
}

}  // extern "C"
