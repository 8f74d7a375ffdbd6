// ORACLE — TEST INFRASTRUCTURE ONLY (see yk_oracle.h). C entry points over yko_*.h.
#include "yk_oracle.h"

#include <chrono>
#include <cstring>
#include <memory>

#include "yko_render.h"

using namespace yko;

struct yko_scene {
    Scene s;
};

static Transform to_xf(const yko_transform& t) {
    Transform r;
    std::memcpy(r.m.m, t.m, sizeof(float) * 16);
    std::memcpy(r.m_inv.m, t.m_inv, sizeof(float) * 16);
    return r;
}
static void from_xf(const Transform& t, yko_transform* o) {
    std::memcpy(o->m, t.m.m, sizeof(float) * 16);
    std::memcpy(o->m_inv, t.m_inv.m, sizeof(float) * 16);
}
static V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }

extern "C" {

yko_scene* yko_scene_create(const yko_host_scene_desc* d) {
    auto sc = std::make_unique<yko_scene>();
    Scene& s = sc->s;
    for (uint32_t i = 0; i < d->n_textures; ++i) {
        const yko_texture_desc& t = d->textures[i];
        Texture tex;
        tex.kind = (TextureKind)t.kind;
        tex.value = {t.value[0], t.value[1], t.value[2]};
        tex.width = t.width;
        tex.height = t.height;
        if (t.kind == TEX_IMAGE) tex.texels.assign(t.texels, t.texels + (size_t)t.width * t.height * 3);
        s.textures.push_back(std::move(tex));
    }
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const yko_material_desc& m = d->materials[i];
        s.materials.push_back({(MaterialKind)m.kind, {m.tex[0], m.tex[1], m.tex[2]}, m.eta, m.remap_roughness != 0});
    }
    for (uint32_t i = 0; i < d->n_lights; ++i) {
        const yko_light_desc& l = d->lights[i];
        Light L{};
        L.kind = (LightKind)l.kind;
        L.i = {l.intensity[0], l.intensity[1], l.intensity[2]};
        Transform l2w = to_xf(l.light_to_world);
        switch (L.kind) {
            case LIGHT_POINT:  // point_light.rs:18-24
                L.p = xf_point(l2w, v3(0, 0, 0));
                break;
            case LIGHT_SPOT:  // spot_light.rs:23-36
                L.world_to_light = xf_inverted(l2w);
                L.p = xf_point(l2w, v3(0, 0, 0));
                L.cos_total_width = std::cos(to_radians(l.total_width_deg));
                L.cos_falloff_start = std::cos(to_radians(l.falloff_start_deg));
                break;
            case LIGHT_RECT: {  // rectangular_light.rs:27-43
                Transform sample_to_light = xf_mul(xf_scale(l.size[0], 1.0f, l.size[1]), xf_translation(v3(-0.5f, 0.0f, -0.5f)));
                L.sample_to_world = xf_mul(l2w, sample_to_light);
                L.area = l.size[0] * l.size[1];
            } break;
            case LIGHT_DISTANT:  // distant_light.rs:17-21
                L.p = ld3(l.direction);
                break;
        }
        s.lights.push_back(L);
    }
    uint32_t orig = 0;
    auto add_mesh = [&](uint32_t i) {
        const yko_mesh_desc& md = d->meshes[i];
        Transform o2w = to_xf(md.object_to_world);
        Mesh mesh;
        mesh.points.resize(md.n_points);
        for (uint32_t k = 0; k < md.n_points; ++k) mesh.points[k] = xf_point(o2w, ld3(md.points + 3 * k));  // mesh.rs:27-29
        if (md.normals) {
            mesh.normals.resize(md.n_points);
            for (uint32_t k = 0; k < md.n_points; ++k) mesh.normals[k] = xf_normal(o2w, ld3(md.normals + 3 * k));
        }
        if (md.uvs) {
            mesh.uvs.resize(md.n_points);
            for (uint32_t k = 0; k < md.n_points; ++k) mesh.uvs[k] = {md.uvs[2 * k], md.uvs[2 * k + 1]};
        }
        mesh.transform_swaps_handedness = xf_swaps_handedness(o2w);
        s.meshes.push_back(std::move(mesh));
        const uint32_t mesh_index = (uint32_t)s.meshes.size() - 1;
        for (uint32_t v0 = 0; v0 + 2 < md.n_indices; v0 += 3)
            s.shapes.push_back({mesh_index, {md.indices[v0], md.indices[v0 + 1], md.indices[v0 + 2]}, md.material, md.area_light, orig++});
    };
    auto add_sphere = [&](uint32_t k) {  // Sphere::new, shapes/sphere.rs:23-33
        const yko_sphere_desc& sd = d->spheres[k];
        Sphere sp;
        sp.object_to_world = to_xf(sd.object_to_world);
        sp.world_to_object = xf_inverted(sp.object_to_world);
        sp.radius = sd.radius;
        sp.material = sd.material;
        sp.transform_swaps_handedness = xf_swaps_handedness(sp.object_to_world);
        s.spheres.push_back(sp);
        Triangle shape{0, {0, 0, 0}, sd.material, -1, orig++};
        shape.sphere = (int32_t)s.spheres.size() - 1;
        s.shapes.push_back(shape);
    };
    if (d->objects) {  // declaration order of a loaded file (pbrt/mod.rs:797-809)
        for (uint32_t i = 0; i < d->n_objects; ++i) {
            if (d->objects[i] >= 0) add_mesh((uint32_t)d->objects[i]);
            else add_sphere((uint32_t)(-1 - d->objects[i]));
        }
    } else {  // meshes, then spheres (scene/mod.rs:497)
        for (uint32_t i = 0; i < d->n_meshes; ++i) add_mesh(i);
        for (uint32_t k = 0; k < d->n_spheres; ++k) add_sphere(k);
    }
    s.background = {d->background[0], d->background[1], d->background[2]};
    s.max_shapes_in_node = d->max_shapes_in_node;
    s.split_method = (SplitMethod)d->split_method;
    if (!s.build_bvh()) return nullptr;
    return sc.release();
}
void yko_scene_destroy(yko_scene* s) { delete s; }
uint32_t yko_scene_node_count(const yko_scene* s) { return (uint32_t)s->s.nodes.size(); }
uint32_t yko_scene_shape_count(const yko_scene* s) { return (uint32_t)s->s.shapes.size(); }
void yko_scene_copy_nodes(const yko_scene* s, void* out) { std::memcpy(out, s->s.nodes.data(), s->s.nodes.size() * sizeof(BVHNode)); }
void yko_scene_copy_order(const yko_scene* s, uint32_t* out) {
    for (size_t i = 0; i < s->s.shapes.size(); ++i) out[i] = s->s.shapes[i].orig_id;
}

static bool make_camera(const yko_camera_params* p, uint32_t rx, uint32_t ry, Camera* cam) {
    return camera_new(ld3(p->position), ld3(p->target), ld3(p->up), (FovAxis)p->fov_axis, p->fov_deg, rx, ry, cam);
}

int yko_render(const yko_scene* sc, const yko_camera_params* cp, const yko_film_settings* fs, const yko_sampler_desc* sd,
               const yko_integrator_desc* id, const yko_tile* tiles, uint32_t n_tiles, uint32_t n_threads, float* film,
               int32_t* hit_ids, uint32_t aux_sample, yko_stats* stats) {
    Camera cam;
    if (!make_camera(cp, fs->res_x, fs->res_y, &cam)) return -1;
    Sampler sampler{};
    sampler.kind = (SamplerKind)sd->kind;
    sampler.nx = sd->nx;
    sampler.ny = sd->kind == SAMPLER_UNIFORM ? 1 : sd->ny;
    sampler.jitter = sd->jitter != 0;
    sampler.seed = sd->seed;
    Integrator integ{(IntegratorKind)id->kind, id->max_depth, id->has_clamp != 0, id->indirect_clamp};
    std::vector<FilmTile> tl;
    if (tiles) {
        for (uint32_t i = 0; i < n_tiles; ++i)
            tl.push_back({tiles[i].x0, tiles[i].y0, tiles[i].x1, tiles[i].y1, tiles[i].sample, tiles[i].index});
    } else {
        tl = film_tiles(fs->res_x, fs->res_y, fs->tile_dim);
    }
    RenderTotals t = render(sc->s, cam, sampler, integ, fs->res_x, fs->res_y, fs->accumulate != 0, tiles == nullptr, tl, n_threads,
                            {film, hit_ids, aux_sample});
    if (stats) {
        stats->ray_count = t.ray_count; stats->shadow_rays = t.shadow_rays; stats->samples = t.samples;
        stats->closest_nodes = t.ts.closest_nodes; stats->closest_tris = t.ts.closest_tris;
        stats->any_nodes = t.ts.any_nodes; stats->any_tris = t.ts.any_tris;
        stats->primary_hit_hash = t.primary_hit_hash;
        stats->seconds = t.seconds; stats->threads = t.threads; stats->_pad = 0;
    }
    return 0;
}

int yko_debug_ray_path(const yko_scene* sc, const yko_camera_params* cp, const yko_film_settings* fs, const yko_sampler_desc* sd,
                       const yko_integrator_desc* id, uint32_t px, uint32_t py, yko_debug_ray* out, uint32_t cap, float* li_rgb,
                       uint64_t* ray_count) {
    Camera cam;
    if (!make_camera(cp, fs->res_x, fs->res_y, &cam)) return -1;
    // `sampler.instantiate(false).as_ref().clone()` (window.rs:884): never started on a pixel sample
    Sampler sampler{};
    sampler.kind = (SamplerKind)sd->kind;
    sampler.nx = sd->nx;
    sampler.ny = sd->kind == SAMPLER_UNIFORM ? 1 : sd->ny;
    sampler.jitter = sd->jitter != 0;
    sampler.seed = sd->seed;
    sampler.rng = Pcg32::make(sd->seed, 0);
    Integrator integ{(IntegratorKind)id->kind, id->max_depth, id->has_clamp != 0, id->indirect_clamp};
    ThreadStats st;
    RenderCtx ctx{sc->s, integ, &st};
    const V2 jitter = sampler.get_2d();
    const Ray ray = camera_ray(cam, V2{(float)px + jitter.x, (float)py + jitter.y});  // window.rs:886-888
    RayLog rays;
    const RadianceResult r = integrator_li_debug(ctx, ray, sampler, &rays);
    for (size_t i = 0; i < rays.size() && i < cap; ++i) {
        out[i].o[0] = rays[i].ray.o.x; out[i].o[1] = rays[i].ray.o.y; out[i].o[2] = rays[i].ray.o.z;
        out[i].d[0] = rays[i].ray.d.x; out[i].d[1] = rays[i].ray.d.y; out[i].d[2] = rays[i].ray.d.z;
        out[i].t_max = rays[i].ray.t_max;
        out[i].ray_type = rays[i].ray_type;
    }
    if (li_rgb) { li_rgb[0] = r.li.r; li_rgb[1] = r.li.g; li_rgb[2] = r.li.b; }
    if (ray_count) *ray_count = r.rays;
    return (int)rays.size();
}

uint32_t yko_film_tiles(uint32_t rx, uint32_t ry, uint32_t dim, yko_tile* out, uint32_t cap) {
    std::vector<FilmTile> t = film_tiles(rx, ry, dim);
    for (size_t i = 0; i < t.size() && i < cap; ++i) out[i] = {t[i].x0, t[i].y0, t[i].x1, t[i].y1, t[i].sample, 0, t[i].index};
    return (uint32_t)t.size();
}
int yko_camera_make(const yko_camera_params* p, uint32_t rx, uint32_t ry, float* c2w, float* r2c) {
    Camera cam;
    if (!make_camera(p, rx, ry, &cam)) return -1;
    std::memcpy(c2w, cam.camera_to_world.m.m, 64);
    std::memcpy(r2c, cam.raster_to_camera.m.m, 64);
    return 0;
}
void yko_camera_rays(const yko_camera_params* p, uint32_t rx, uint32_t ry, const float* pf, uint32_t n, float* o, float* d) {
    Camera cam;
    if (!make_camera(p, rx, ry, &cam)) return;
    for (uint32_t i = 0; i < n; ++i) {
        Ray r = camera_ray(cam, V2{pf[2 * i], pf[2 * i + 1]});
        o[3 * i] = r.o.x; o[3 * i + 1] = r.o.y; o[3 * i + 2] = r.o.z;
        d[3 * i] = r.d.x; d[3 * i + 1] = r.d.y; d[3 * i + 2] = r.d.z;
    }
}
void yko_xf_identity(yko_transform* o) { from_xf(xf_identity(), o); }
void yko_xf_translation(const float* d, yko_transform* o) { from_xf(xf_translation(ld3(d)), o); }
void yko_xf_scale(float x, float y, float z, yko_transform* o) { from_xf(xf_scale(x, y, z), o); }
void yko_xf_rotation(float theta, const float* axis, yko_transform* o) { from_xf(xf_rotation(theta, ld3(axis)), o); }
int yko_xf_new(const float* m16, yko_transform* o) {
    M44 m;
    std::memcpy(m.m, m16, 64);
    Transform t;
    if (!xf_new(m, &t)) return -1;
    from_xf(t, o);
    return 0;
}
int yko_xf_look_at(const float* pos, const float* target, const float* up, yko_transform* o) {
    Transform t;
    if (!xf_look_at(ld3(pos), ld3(target), ld3(up), &t)) return -1;
    from_xf(t, o);
    return 0;
}
void yko_xf_mul(const yko_transform* a, const yko_transform* b, yko_transform* o) { from_xf(xf_mul(to_xf(*a), to_xf(*b)), o); }
void yko_xf_inverted(const yko_transform* a, yko_transform* o) { from_xf(xf_inverted(to_xf(*a)), o); }
static void st3(V3 v, float* o) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
void yko_xf_point(const yko_transform* t, const float* p, float* o) { st3(xf_point(to_xf(*t), ld3(p)), o); }
void yko_xf_vec(const yko_transform* t, const float* p, float* o) { st3(xf_vec(to_xf(*t), ld3(p)), o); }
void yko_xf_normal(const yko_transform* t, const float* p, float* o) { st3(xf_normal(to_xf(*t), ld3(p)), o); }
void yko_cross(const float* a, const float* b, float* o) { st3(cross(ld3(a), ld3(b)), o); }

// One entry point over the vector / bounds / transform helpers of yko_math.h the hot path is built from, so that the
// reference's own unit tests (tests/src/{vector,normal,point,ray,bounds,transform}.rs) can be replayed against them
// (tests/test_oracle_math.py). `in` / `out` are flat float arrays whose meaning depends on `op`.
int yko_math_kat(uint32_t op, const float* in, float* out) {
    const V3 a = ld3(in), b = ld3(in + 3);
    switch (op) {
        case 0: out[0] = dot(a, b); return 0;                 // Vec3::dot
        case 1: out[0] = dot_nv(a, b); return 0;              // Vec3::dot_n / Normal::dot_v
        case 2: out[0] = len_sqr(a); return 0;
        case 3: out[0] = len(a); return 0;
        case 4: st3(normalized(a), out); return 0;
        case 5: st3(vmin(a, b), out); return 0;
        case 6: st3(vmax(a, b), out); return 0;
        case 7: out[0] = min_comp(a); return 0;
        case 8: out[0] = max_comp(a); return 0;
        case 9: out[0] = (float)max_dimension(a); return 0;
        case 10: st3(permuted(a, (int)in[3], (int)in[4], (int)in[5]), out); return 0;
        case 11: { Bounds3 r = union_p(bounds_new(a, b), ld3(in + 6)); st3(r.p_min, out); st3(r.p_max, out + 3); return 0; }
        case 12: { Bounds3 r = union_b(bounds_new(a, b), bounds_new(ld3(in + 6), ld3(in + 9))); st3(r.p_min, out); st3(r.p_max, out + 3); return 0; }
        case 13: st3(diagonal(bounds_new(a, b)), out); return 0;
        case 14: st3(offset(bounds_new(a, b), ld3(in + 6)), out); return 0;
        case 15: out[0] = surface_area(bounds_new(a, b)); return 0;
        case 16: out[0] = (float)maximum_extent(bounds_new(a, b)); return 0;
        case 17: { Bounds3 r = bounds_default(); st3(r.p_min, out); st3(r.p_max, out + 3); return 0; }
        case 18: {  // Transform::new(m).swaps_handedness()
            M44 m;
            for (int i = 0; i < 16; ++i) m.m[i / 4][i % 4] = in[i];
            Transform t;
            if (!xf_new(m, &t)) return -1;
            out[0] = xf_swaps_handedness(t) ? 1.0f : 0.0f;
            return 0;
        }
        case 19: { Ray r{a, b, in[6]}; st3(r.o + r.d * in[7], out); return 0; }  // Ray::point(t)
        case 20: out[0] = len(b - a); return 0;               // Point3::dist
        case 21: out[0] = len_sqr(b - a); return 0;           // Point3::dist_sqr
        case 22: st3(faceforward(a, b), out); return 0;
        default: return -2;
    }
}

// The BxDFs of materials/bsdfs/*.rs in their local frame (z = normal), for analytic property tests of the restatement
// (tests/test_oracle_bsdf.py: pdf normalisation, sample / eval consistency, energy bounds, Fresnel closed forms).
// kind: 0 Lambertian(r), 1 Oren-Nayar(r, sigma rad), 2 specular reflection(r, eta), 3 specular transmission(r, eta),
// 4 Torrance-Sparrow / GGX with conductor Fresnel(eta_t rgb, k rgb, alpha), 5 the same with Schlick Fresnel(rs rgb, alpha).
// mode 0: in = n x (wo, wi) -> out = n x (f rgb, pdf); mode 1: in = n x (wo, u0, u1, 0) -> out = n x (wi, f rgb, pdf, type).
int yko_lobe_eval(uint32_t kind, const float* params, uint32_t mode, const float* in, uint32_t n, float* out) {
    Lobe l{};
    l.r = spec(params[0], params[1], params[2]);
    switch (kind) {
        case 0: l.kind = LOBE_LAMBERT; break;
        case 1: {  // oren_nayar.rs:18-25
            l.kind = LOBE_OREN_NAYAR;
            const float s2 = params[3] * params[3];
            l.a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
            l.b = 0.45f * s2 / (s2 + 0.09f);
        } break;
        case 2: l.kind = LOBE_SPEC_REFL; l.fresnel = FRESNEL_DIELECTRIC; l.eta_i = 1.0f; l.eta_t = params[3]; break;
        case 3: l.kind = LOBE_SPEC_TRANS; l.eta_i = 1.0f; l.eta_t = params[3]; break;
        case 4:
            l.kind = LOBE_MICROFACET; l.fresnel = FRESNEL_CONDUCTOR; l.c_eta_i = spec1(1.0f);
            l.c_eta_t = spec(params[3], params[4], params[5]); l.c_k = spec(params[6], params[7], params[8]);
            l.alpha = fmax_(params[9], 0.001f);
            break;
        case 5:
            l.kind = LOBE_MICROFACET; l.fresnel = FRESNEL_SCHLICK; l.rs = spec(params[3], params[4], params[5]);
            l.alpha = fmax_(params[6], 0.001f);
            break;
        default: return -1;
    }
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 6 * (size_t)i;
        const V3 wo = ld3(a);
        if (mode == 0) {
            const V3 wi = ld3(a + 3);
            const Spec f = l.f(wo, wi);
            float* o = out + 4 * (size_t)i;
            o[0] = f.r; o[1] = f.g; o[2] = f.b; o[3] = l.pdf(wo, wi);
        } else {
            const BxdfSample sm = l.sample_f(wo, V2{a[3], a[4]});
            float* o = out + 8 * (size_t)i;
            st3(sm.wi, o);
            o[3] = sm.f.r; o[4] = sm.f.g; o[5] = sm.f.b; o[6] = sm.pdf; o[7] = (float)sm.sample_type;
        }
    }
    return 0;
}

// Light::sample_li (lights/*.rs) of light `light` for n shading points: in = n x (p xyz, n xyz, u0, u1),
// out = n x (l xyz, li rgb, pdf, has_vis, vis ray o xyz, vis ray d xyz) = 14 floats. For tests/test_oracle_lights.py.
int yko_light_sample(const yko_scene* sc, uint32_t light, const float* in, uint32_t n, float* out) {
    const Scene& s = sc->s;
    if (light >= s.lights.size()) return -1;
    for (uint32_t i = 0; i < n; ++i) {
        const float* a = in + 8 * (size_t)i;
        SurfaceInteraction si{};
        si.p = ld3(a);
        si.n = ld3(a + 3);
        si.sh_n = si.n;
        si.area_light = -1;
        const LightSample ls = sample_li(s, (int32_t)light, si, V2{a[6], a[7]});
        float* o = out + 14 * (size_t)i;
        st3(ls.l, o);
        o[3] = ls.li.r; o[4] = ls.li.g; o[5] = ls.li.b;
        o[6] = ls.pdf;
        o[7] = ls.has_vis ? 1.0f : 0.0f;
        st3(ls.has_vis ? ls.vis_ray.o : v3(0, 0, 0), o + 8);
        st3(ls.has_vis ? ls.vis_ray.d : v3(0, 0, 0), o + 11);
    }
    return 0;
}

uint64_t yko_siphash13(const uint8_t* msg, uint64_t n) { return siphash13(msg, (size_t)n); }
void yko_pcg32_sequence(uint64_t state, uint64_t stream, uint64_t adv, uint32_t n, uint32_t* out) {
    Pcg32 p = Pcg32::make(state, stream);
    p.advance(adv);
    for (uint32_t i = 0; i < n; ++i) out[i] = p.next_u32();
}
uint32_t yko_permutation_element(uint32_t i, uint32_t l, uint32_t p) { return permutation_element(i, l, p); }
void yko_sampler_draws(const yko_sampler_desc* sd, uint32_t px, uint32_t py, uint32_t index, uint32_t start_dim,
                       const uint8_t* pattern, uint32_t n, float* out) {
    Sampler s{};
    s.kind = (SamplerKind)sd->kind; s.nx = sd->nx; s.ny = sd->kind == SAMPLER_UNIFORM ? 1 : sd->ny;
    s.jitter = sd->jitter != 0; s.seed = sd->seed;
    s.start_pixel_sample((uint16_t)px, (uint16_t)py, index, start_dim);
    for (uint32_t k = 0; k < n; ++k) {
        if (pattern[k] == 1) *out++ = s.get_1d();
        else { V2 v = s.get_2d(); *out++ = v.x; *out++ = v.y; }
    }
}

void yko_trace(const yko_scene* sc, const float* o, const float* d, const float* tmax, uint32_t n, int brute, float* t_out,
               int32_t* id_out, uint32_t* counts) {
    const Scene& s = sc->s;
    for (uint32_t i = 0; i < n; ++i) {
        Ray ray{ld3(o + 3 * i), ld3(d + 3 * i), tmax ? tmax[i] : INFINITY};
        float best_t = INFINITY;
        int32_t best = -1;
        if (brute) {
            for (size_t k = 0; k < s.shapes.size(); ++k) {
                float t;
                if (s.triangle_intersect(s.shapes[k], ray, &t, nullptr, false)) { ray.t_max = t; best_t = t; best = (int32_t)s.shapes[k].orig_id; }
            }
        } else {
            IntersectionResult r = s.intersect(ray, nullptr);
            if (r.has_hit) { best_t = r.hit.t; best = (int32_t)s.shapes[r.hit.shape].orig_id; }
            if (counts) { counts[2 * i] = (uint32_t)r.intersection_test_count; counts[2 * i + 1] = (uint32_t)r.intersection_count; }
        }
        t_out[i] = best_t;
        id_out[i] = best;
    }
}
void yko_occluded(const yko_scene* sc, const float* o, const float* d, const float* tmax, uint32_t n, int brute, uint8_t* out) {
    const Scene& s = sc->s;
    for (uint32_t i = 0; i < n; ++i) {
        Ray ray{ld3(o + 3 * i), ld3(d + 3 * i), tmax ? tmax[i] : INFINITY};
        bool occ = false;
        if (brute) {
            for (size_t k = 0; k < s.shapes.size() && !occ; ++k) {
                float t;
                occ = s.triangle_intersect(s.shapes[k], ray, &t, nullptr, false);
            }
        } else {
            occ = s.any_intersect(ray, -1, nullptr);
        }
        out[i] = occ ? 1 : 0;
    }
}

}  // extern "C"
