/* ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
 *
 * C entry points of the CPU restatement of yuki's per-pixel rendering hot path (liboracle.so).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library. The reference (Rust) cannot be compiled in this image (no cargo/rustc), so
 * there is no oracle/_ref; PARITY IS UNPINNED for everything but the math subset (see yko_*.h).
 *
 * The descriptor structs below deliberately have the same memory layout as the product's
 * include/yuki_gpu.h host-side descriptors so one Python scene description feeds both.
 */
#ifndef YK_ORACLE_H
#define YK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float m[16]; float m_inv[16]; } yko_transform;

typedef struct {
    uint32_t kind;            /* 0 constant, 1 image */
    float value[3];
    uint32_t width, height;
    const float* texels;      /* image: width*height*3 f32 RGB, first row = top row of the file */
} yko_texture_desc;

typedef struct {
    uint32_t kind;            /* 0 matte, 1 glass, 2 metal, 3 glossy */
    int32_t tex[3];           /* matte: kd, sigma | glass: r, t | metal: eta, k, roughness | glossy: rs, roughness */
    float eta;
    uint32_t remap_roughness;
} yko_material_desc;

typedef struct {
    uint32_t kind;            /* 0 point, 1 spot, 2 rect, 3 distant */
    yko_transform light_to_world;
    float intensity[3];       /* I (point/spot), L (rect), radiance (distant) */
    float total_width_deg, falloff_start_deg;
    float size[2];
    float direction[3];
} yko_light_desc;

typedef struct {
    yko_transform object_to_world;
    uint32_t n_points;
    uint32_t n_indices;
    const float* points;      /* object space, n_points*3 */
    const float* normals;     /* NULL or n_points*3 */
    const float* uvs;         /* NULL or n_points*2 */
    const uint32_t* indices;  /* n_indices, triplets */
    int32_t material;
    int32_t area_light;       /* index of a rect light, or -1 */
} yko_mesh_desc;

typedef struct {
    yko_transform object_to_world;
    float radius;
    int32_t material;
} yko_sphere_desc;

typedef struct {
    uint32_t n_meshes, n_textures, n_materials, n_lights;
    const yko_mesh_desc* meshes;
    const yko_texture_desc* textures;
    const yko_material_desc* materials;
    const yko_light_desc* lights;
    float background[3];
    uint32_t max_shapes_in_node;
    uint32_t split_method;    /* 0 SAH, 1 Middle, 2 EqualCounts */
    uint32_t n_spheres;       /* shapes = the meshes' triangles in order, then the spheres (scene/mod.rs:497) ... */
    const yko_sphere_desc* spheres;
    uint32_t n_objects;       /* ... unless `objects` lists the declaration order: mesh index, or -1 - sphere index */
    const int32_t* objects;
} yko_host_scene_desc;

typedef struct { float position[3]; float target[3]; float up[3]; uint32_t fov_axis; float fov_deg; } yko_camera_params;
typedef struct { uint32_t res_x, res_y, tile_dim, accumulate; } yko_film_settings;
typedef struct { uint32_t kind; uint32_t nx, ny; uint32_t jitter; uint64_t seed; } yko_sampler_desc;
typedef struct { uint32_t kind; uint32_t max_depth; uint32_t has_clamp; float indirect_clamp; } yko_integrator_desc;
typedef struct { uint16_t x0, y0, x1, y1; uint16_t sample; uint16_t _pad; uint32_t index; } yko_tile;

typedef struct {
    uint64_t ray_count, shadow_rays, samples;
    uint64_t closest_nodes, closest_tris, any_nodes, any_tris;
    uint64_t primary_hit_hash;
    double seconds;
    uint32_t threads;
    uint32_t _pad;
} yko_stats;

typedef struct yko_scene yko_scene;

yko_scene* yko_scene_create(const yko_host_scene_desc* desc);
void yko_scene_destroy(yko_scene*);
uint32_t yko_scene_node_count(const yko_scene*);
uint32_t yko_scene_shape_count(const yko_scene*);
void yko_scene_copy_nodes(const yko_scene*, void* out32B);        /* node_count * 32 bytes */
void yko_scene_copy_order(const yko_scene*, uint32_t* orig_ids);  /* shape_count */

int yko_render(const yko_scene*, const yko_camera_params*, const yko_film_settings*, const yko_sampler_desc*,
               const yko_integrator_desc*, const yko_tile* tiles, uint32_t n_tiles, uint32_t n_threads,
               float* film_rgb, int32_t* hit_ids, uint32_t aux_sample, yko_stats* stats);

/* launch_debug_ray (app/window.rs:812-905) + Integrator::li_debug (integrators/mod.rs:103-118): the rays of one path through
   film pixel (px, py), traced with a freshly cloned sampler (pixel (0,0), sample 0, PCG stream 0, no seek). ray_type:
   0 direct, 1 reflection, 2 refraction, 3 normal, 4 shadow (integrators/mod.rs:82-89). Returns the number of rays the
   integrator collected (at most `cap` are written), or -1 on a degenerate camera. */
typedef struct { float o[3]; float d[3]; float t_max; uint32_t ray_type; } yko_debug_ray;
int yko_debug_ray_path(const yko_scene*, const yko_camera_params*, const yko_film_settings*, const yko_sampler_desc*,
                       const yko_integrator_desc*, uint32_t px, uint32_t py, yko_debug_ray* out, uint32_t cap,
                       float* li_rgb, uint64_t* ray_count);

/* host helpers restated from film.rs / camera.rs / math/transforms.rs */
uint32_t yko_film_tiles(uint32_t res_x, uint32_t res_y, uint32_t tile_dim, yko_tile* out, uint32_t cap);
int yko_camera_make(const yko_camera_params*, uint32_t res_x, uint32_t res_y, float* camera_to_world16,
                    float* raster_to_camera16);
void yko_camera_rays(const yko_camera_params*, uint32_t res_x, uint32_t res_y, const float* p_film_xy, uint32_t n,
                     float* o_xyz, float* d_xyz);
void yko_xf_identity(yko_transform*);
void yko_xf_translation(const float* d3, yko_transform*);
void yko_xf_scale(float x, float y, float z, yko_transform*);
void yko_xf_rotation(float theta, const float* axis3, yko_transform*);
int yko_xf_new(const float* m16, yko_transform*);
int yko_xf_look_at(const float* pos3, const float* target3, const float* up3, yko_transform*);
void yko_xf_mul(const yko_transform* a, const yko_transform* b, yko_transform* out);
void yko_xf_inverted(const yko_transform* a, yko_transform* out);
void yko_xf_point(const yko_transform*, const float* p3, float* out3);
void yko_xf_vec(const yko_transform*, const float* v3, float* out3);
void yko_xf_normal(const yko_transform*, const float* n3, float* out3);
void yko_cross(const float* a3, const float* b3, float* out3);
int yko_math_kat(uint32_t op, const float* in, float* out);
/* one BxDF of materials/bsdfs in its local frame: f / pdf (mode 0) or sample_f (mode 1) for n inputs, see yk_oracle.cpp */
int yko_lobe_eval(uint32_t kind, const float* params, uint32_t mode, const float* in, uint32_t n, float* out);
/* Light::sample_li of one light of the scene for n shading points, see yk_oracle.cpp */
int yko_light_sample(const yko_scene*, uint32_t light, const float* in, uint32_t n, float* out);       /* yko_math.h helpers by opcode, for the reference's unit-test KATs */

/* KAT / self-check helpers */
uint64_t yko_siphash13(const uint8_t* msg, uint64_t n);
void yko_pcg32_sequence(uint64_t state, uint64_t stream, uint64_t advance, uint32_t n, uint32_t* out);
uint32_t yko_permutation_element(uint32_t i, uint32_t l, uint32_t p);
/* draws: pattern[k] == 1 -> get_1d (1 float out), 2 -> get_2d (2 floats out) */
void yko_sampler_draws(const yko_sampler_desc*, uint32_t px, uint32_t py, uint32_t index, uint32_t start_dim,
                       const uint8_t* pattern, uint32_t n, float* out);
/* closest hit per ray through the BVH and by brute force over all triangles (same triangle test) */
void yko_trace(const yko_scene*, const float* o_xyz, const float* d_xyz, const float* t_max, uint32_t n,
               int brute_force, float* t_out, int32_t* orig_id_out, uint32_t* counts_out /* n*2 or NULL */);
void yko_occluded(const yko_scene*, const float* o_xyz, const float* d_xyz, const float* t_max, uint32_t n,
                  int brute_force, uint8_t* out);

#ifdef __cplusplus
}
#endif
#endif
