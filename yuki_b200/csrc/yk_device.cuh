// Device-side building blocks of the B200 backend: f32 vector math, yuki's seekable samplers, surface
// interaction set-up, BSDFs and lights. Everything here is evaluated with the reference's operation
// order; the translation unit is compiled with --fmad=false (Rust never contracts to FMA) and IEEE
// division / square root, so primary-hit ids and BVH counters are bit-exact against the CPU path.
// sin/cos restate glibc's sinf/cosf bit for bit (yk_libm.h): those are the libm results the reference gets from
// f32::sin/cos, and a 1-ulp difference there is enough to flip a shadow-ray decision a few bounces later.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "yk_fastdiv.h"
#include "yk_libm.h"
#include "yuki_gpu.h"

#define YK_DEV __device__ __forceinline__

namespace ykd {

constexpr float kPi = 3.14159274101257324f;
constexpr float kInvPi = 0.318309886183790672f;
constexpr float kPiOver2 = 1.57079632679489662f;
constexpr float kPiOver4 = 0.785398163397448310f;

struct V3 {
    float x, y, z;
};
struct V2 {
    float x, y;
};
struct RGB {
    float r, g, b;
};

// IEEE division. CUDA's correctly rounded `/` leaves its inline fast path for a subroutine whenever the numerator is
// zero, which axis-aligned normals and single-channel colours make the common case here (profiles/r01: 11 % of the
// shading kernel's instructions). 0 / b for finite non-zero b is a signed zero; everything else takes `/`.
YK_DEV float fdiv(float a, float b) {
    const uint32_t ub = __float_as_uint(b) & 0x7fffffffu;
    if (a == 0.0f && ub - 1u < 0x7f7fffffu) return __uint_as_float((__float_as_uint(a) ^ __float_as_uint(b)) & 0x80000000u);
    return a / b;
}
YK_DEV V3 mk(float x, float y, float z) { return V3{x, y, z}; }
YK_DEV V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
YK_DEV V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
YK_DEV V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
YK_DEV V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
YK_DEV V3 operator/(V3 a, float s) { return {fdiv(a.x, s), fdiv(a.y, s), fdiv(a.z, s)}; }
// Vec3::dot has a leading zero term, dot_n/dot_v do not (yuki_derive/src/impl_vec_like.rs:193-197,
// math/vector.rs:228-230, math/normal.rs:57-59).
YK_DEV float dot0(V3 a, V3 b) { return 0.0f + a.x * b.x + a.y * b.y + a.z * b.z; }
YK_DEV float dotn(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
YK_DEV float length(V3 a) { return sqrtf(dot0(a, a)); }  // == (f64 sqrt) as f32 for f32 input
YK_DEV V3 unit(V3 a) { return a / length(a); }
YK_DEV V3 cross64(V3 a, V3 b) {  // math/vector.rs:236-255
    const double ax = a.x, ay = a.y, az = a.z, bx = b.x, by = b.y, bz = b.z;
    return {(float)(ay * bz - az * by), (float)(az * bx - ax * bz), (float)(ax * by - ay * bx)};
}
YK_DEV float comp(V3 a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
YK_DEV V3 flip_toward(V3 n, V3 v) { return dotn(n, v) < 0.0f ? -n : n; }    // Normal::faceforward_v
YK_DEV V3 flip_toward_n(V3 n, V3 m) { return dot0(n, m) < 0.0f ? -n : n; }  // Normal::faceforward_n
YK_DEV float clamp01ish(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }  // f32::clamp
YK_DEV float sin_f32(float x) { return yklibm::sinf_glibc(x); }
YK_DEV float cos_f32(float x) { return yklibm::cosf_glibc(x); }

// math/mod.rs:26-34, including the reference's un-rooted divisor in the else branch.
YK_DEV void frame_from(V3 v, V3* a, V3* b) {
    if (fabsf(v.x) > fabsf(v.y)) *a = mk(-v.z, 0.0f, v.x) / sqrtf(v.x * v.x + v.z * v.z);
    else *a = mk(0.0f, v.z, -v.y) / (v.y * v.y + v.z + v.z);
    *b = cross64(v, *a);
}

YK_DEV RGB rgb(float r, float g, float b) { return RGB{r, g, b}; }
YK_DEV RGB gray(float v) { return RGB{v, v, v}; }
YK_DEV RGB operator+(RGB a, RGB b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
YK_DEV RGB operator-(RGB a, RGB b) { return {a.r - b.r, a.g - b.g, a.b - b.b}; }
YK_DEV RGB operator*(RGB a, RGB b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
YK_DEV RGB operator/(RGB a, RGB b) { return {fdiv(a.r, b.r), fdiv(a.g, b.g), fdiv(a.b, b.b)}; }
YK_DEV RGB operator*(RGB a, float s) { return {a.r * s, a.g * s, a.b * s}; }
YK_DEV RGB operator/(RGB a, float s) { return {fdiv(a.r, s), fdiv(a.g, s), fdiv(a.b, s)}; }
YK_DEV bool black(RGB a) { return a.r == 0.0f && a.g == 0.0f && a.b == 0.0f; }
YK_DEV RGB rsqrt3(RGB a) { return {sqrtf(a.r), sqrtf(a.g), sqrtf(a.b)}; }

// Transform application on a row-major f32[16] (math/transform.rs:103-162)
YK_DEV V3 xf_vec(const float* m, V3 v) {
    return {m[0] * v.x + m[1] * v.y + m[2] * v.z, m[4] * v.x + m[5] * v.y + m[6] * v.z, m[8] * v.x + m[9] * v.y + m[10] * v.z};
}
YK_DEV V3 xf_point(const float* m, V3 p) {
    const float x = m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3];
    const float y = m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7];
    const float z = m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11];
    const float w = m[12] * p.x + m[13] * p.y + m[14] * p.z + m[15];
    if (w == 1.0f) return {x, y, z};
    return {x / w, y / w, z / w};
}
YK_DEV V3 xf_normal(const float* inv, V3 n) {
    return {inv[0] * n.x + inv[4] * n.y + inv[8] * n.z, inv[1] * n.x + inv[5] * n.y + inv[9] * n.z,
            inv[2] * n.x + inv[6] * n.y + inv[10] * n.z};
}

// ---- samplers -----------------------------------------------------------------------------------
// SipHash-1-3 with zero keys = Rust's DefaultHasher::default() (sampling/mod.rs:89-103); PCG32 =
// rand_pcg 0.3 Lcg64Xsh32; f32 draws = rand 0.8 Standard.
YK_DEV uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
struct Sip {
    uint64_t a, b, c, d;
    YK_DEV void init() {
        a = 0x736f6d6570736575ULL; b = 0x646f72616e646f6dULL; c = 0x6c7967656e657261ULL; d = 0x7465646279746573ULL;
    }
    YK_DEV void round() {
        a += b; b = rotl(b, 13); b ^= a; a = rotl(a, 32);
        c += d; d = rotl(d, 16); d ^= c;
        a += d; d = rotl(d, 21); d ^= a;
        c += b; b = rotl(b, 17); b ^= c; c = rotl(c, 32);
    }
    YK_DEV void absorb(uint64_t m) { d ^= m; round(); a ^= m; }
    YK_DEV uint64_t finish() {
        c ^= 0xff;
        round(); round(); round();
        return a ^ b ^ c ^ d;
    }
};
// message = x:u16 y:u16 (4 bytes) -> only the length/tail block
YK_DEV uint64_t hash_pixel(uint32_t px, uint32_t py) {
    Sip s; s.init();
    s.absorb((4ULL << 56) | (uint64_t)(px & 0xffffu) | ((uint64_t)(py & 0xffffu) << 16));
    return s.finish();
}
// message = x:u16 y:u16 dim:u32 seed:u64 (16 bytes) -> two full words + the length block
static __device__ __noinline__ uint64_t hash_pixel_dim_seed(uint32_t px, uint32_t py, uint32_t dim, uint64_t seed) {  // tabulated on the hot path
    Sip s; s.init();
    s.absorb((uint64_t)(px & 0xffffu) | ((uint64_t)(py & 0xffffu) << 16) | ((uint64_t)dim << 32));
    s.absorb(seed);
    s.absorb(16ULL << 56);
    return s.finish();
}
constexpr uint64_t kPcgMult = 6364136223846793005ULL;
struct Pcg {
    uint64_t state, inc;
    YK_DEV void seed(uint64_t st, uint64_t stream) {
        inc = (stream << 1) | 1ULL;
        state = (st + inc) * kPcgMult + inc;
    }
    YK_DEV void advance(uint64_t delta) {
        uint64_t am = 1, ap = 0, cm = kPcgMult, cp = inc;
        while (delta) {
            if (delta & 1ULL) { am *= cm; ap = ap * cm + cp; }
            cp = (cm + 1ULL) * cp;
            cm *= cm;
            delta >>= 1;
        }
        state = am * state + ap;
    }
    YK_DEV uint32_t next() {
        const uint64_t old = state;
        state = old * kPcgMult + inc;
        const uint32_t xsh = (uint32_t)(((old >> 18) ^ old) >> 27);
        const uint32_t rot = (uint32_t)(old >> 59);
        return __funnelshift_r(xsh, xsh, rot);
    }
    YK_DEV float next_f32() { return (float)(next() >> 8) * (1.0f / 16777216.0f); }
};
YK_DEV uint32_t permutation_element(uint32_t i, const FastDiv& ld, uint32_t p) {  // stratified.rs:147-178
    const uint32_t l = ld.d;
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p; i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8; i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1; i *= 1u | p >> 27;
        i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2; i *= 0x9e501cc3u;
        i ^= (i & w) >> 2; i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return ld.mod(i + p);
}

struct SamplerCfg {
    uint32_t kind, nx, ny, jitter;
    uint64_t seed;
    FastDiv div_n, div_nx, div_ny;  // n = nx * ny
    // hash_values!(pixel, dimension, seed) (stratified.rs:105,122) depends on neither the sample index nor the path, so
    // the 32 bits the permutation uses are tabulated once per pixel group: hash_table[dim * hash_stride + job], dim <
    // n_hash_dims. One SipHash per (pixel, dimension) instead of one per (pixel, dimension, sample).
    const uint32_t* hash_table;
    uint32_t n_hash_dims, hash_stride;
};
// Per-path sampler registers. `dim` is only meaningful for the stratified sampler (it keys the hash).
struct SamplerState {
    Pcg rng;
    uint32_t px, py, index, dim, job;
    // uniform.rs:72-84, stratified.rs:90-102 (always called with dimension 0 by Integrator::render).
    // `inc` = (hash_values!(x, y) << 1) | 1, the pixel's PCG stream, hashed once per pixel (Job::rng_inc).
    YK_DEV void start(const SamplerCfg& c, uint32_t x, uint32_t y, uint32_t sample_index, uint64_t inc, uint32_t job_index) {
        px = x; py = y; index = sample_index; dim = 0; job = job_index;
        rng.inc = inc;
        rng.state = (c.seed + inc) * kPcgMult + inc;  // Lcg64Xsh32::new
        rng.advance((uint64_t)sample_index * 65536ULL);
    }
    YK_DEV uint32_t dim_hash(const SamplerCfg& c) const {
        if (dim < c.n_hash_dims) return __ldg(&c.hash_table[(size_t)dim * c.hash_stride + job]);
        return (uint32_t)hash_pixel_dim_seed(px, py, dim, c.seed);
    }
    YK_DEV float get_1d(const SamplerCfg& c) {
        if (c.kind == YK_SAMPLER_UNIFORM) { dim += 1; return rng.next_f32(); }
        const uint32_t stratum = permutation_element(index, c.div_n, dim_hash(c));
        dim += 1;
        const float delta = c.jitter ? rng.next_f32() : 0.5f;
        return ((float)stratum + delta) / (float)c.div_n.d;
    }
    YK_DEV V2 get_2d(const SamplerCfg& c) {
        if (c.kind == YK_SAMPLER_UNIFORM) {
            dim += 2;
            const float x = rng.next_f32();
            const float y = rng.next_f32();
            return {x, y};
        }
        const uint32_t stratum = permutation_element(index, c.div_n, dim_hash(c));
        dim += 2;
        const uint32_t sx = c.div_nx.mod(stratum);
        const uint32_t sy = c.div_ny.div(stratum);  // reference divides by pixel_samples.y (stratified.rs:128)
        const float dx = c.jitter ? rng.next_f32() : 0.5f;
        const float dy = c.jitter ? rng.next_f32() : 0.5f;
        return {((float)sx + dx) / (float)c.nx, ((float)sy + dy) / (float)c.ny};
    }
};

// sampling/mod.rs:62-87
YK_DEV V3 cosine_hemisphere(V2 u) {
    const float ox = u.x * 2.0f - 1.0f, oy = u.y * 2.0f - 1.0f;
    float dx = 0.0f, dy = 0.0f;
    if (!(ox == 0.0f && oy == 0.0f)) {
        float theta, r;
        if (fabsf(ox) > fabsf(oy)) { theta = kPiOver4 * (oy / ox); r = ox; }
        else { theta = kPiOver2 - kPiOver4 * (ox / oy); r = oy; }
        dx = cos_f32(theta) * r;
        dy = sin_f32(theta) * r;
    }
    return {dx, dy, sqrtf(fmaxf(1.0f - dx * dx - dy * dy, 0.0f))};
}

// ---- geometry -------------------------------------------------------------------------------------
struct Ray {
    V3 o, d;
    float t_max;
};
// Per-ray constants of the watertight triangle test (shapes/triangle.rs:58-80); they depend on the ray
// only, so they are hoisted out of the per-triangle code.
struct TriRay {
    int kx, ky, kz;
    float sx, sy, sz;
    YK_DEV void setup(V3 d) {
        const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
        kz = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);  // Vec3::max_dimension, math/vector.rs:188-202
        kx = kz < 2 ? kz + 1 : 0;
        ky = kx < 2 ? kx + 1 : 0;
        const float dx = comp(d, kx), dy = comp(d, ky), dz = comp(d, kz);
        sx = -dx / dz; sy = -dy / dz; sz = 1.0f / dz;
    }
};
struct TriHit {
    float t, b0, b1, b2;
};
// shapes/triangle.rs:49-139. Accepts t == t_max (strict comparisons at :126-130).
YK_DEV bool tri_test(const TriRay& tr, V3 o, float t_max, V3 p0, V3 p1, V3 p2, TriHit* h) {
    const V3 a = p0 - o, b = p1 - o, c = p2 - o;
    float ax = comp(a, tr.kx), ay = comp(a, tr.ky);
    float bx = comp(b, tr.kx), by = comp(b, tr.ky);
    float cx = comp(c, tr.kx), cy = comp(c, tr.ky);
    const float az = comp(a, tr.kz), bz = comp(b, tr.kz), cz = comp(c, tr.kz);
    ax += tr.sx * az; ay += tr.sy * az;
    bx += tr.sx * bz; by += tr.sy * bz;
    cx += tr.sx * cz; cy += tr.sy * cz;
    float e0 = bx * cy - by * cx;
    float e1 = cx * ay - cy * ax;
    float e2 = ax * by - ay * bx;
    if (e0 == 0.0f || e1 == 0.0f || e2 == 0.0f) {  // f64 fallback, :98-105
        e0 = (float)((double)bx * (double)cy - (double)by * (double)cx);
        e1 = (float)((double)cx * (double)ay - (double)cy * (double)ax);
        e2 = (float)((double)ax * (double)by - (double)ay * (double)bx);
    }
    if ((e0 < 0.0f || e1 < 0.0f || e2 < 0.0f) && (e0 > 0.0f || e1 > 0.0f || e2 > 0.0f)) return false;
    const float det = e0 + e1 + e2;
    if (det == 0.0f) return false;
    const float t_scaled = e0 * (az * tr.sz) + e1 * (bz * tr.sz) + e2 * (cz * tr.sz);
    if ((det < 0.0f && (t_scaled >= 0.0f || t_scaled < t_max * det)) || (det > 0.0f && (t_scaled <= 0.0f || t_scaled > t_max * det)))
        return false;
    const float inv_det = 1.0f / det;
    h->b0 = e0 * inv_det; h->b1 = e1 * inv_det; h->b2 = e2 * inv_det;
    h->t = t_scaled * inv_det;
    return true;
}

// interaction.rs:27-59
YK_DEV Ray spawn_ray(V3 p, V3 n, V3 d) {
    const V3 off = n * 0.001f;
    return {dot0(d, n) > 0.0f ? p + off : p - off, d, __int_as_float(0x7f800000)};
}
YK_DEV Ray spawn_ray_to(V3 p, V3 n, V3 target) {
    const V3 off = n * 0.001f;
    const V3 o = dot0(target - p, n) > 0.0f ? p + off : p - off;
    return {o, target - o, 0.9999f};  // direction deliberately not normalised
}

// SurfaceInteraction after Triangle::intersect's SI part (triangle.rs:141-226)
struct Surface {
    V3 p, n, wo;
    V2 uv;
    V3 sh_n, sh_dpdu;
    int area_light;
};

// ---- sphere (shapes/sphere.rs:36-119) -----------------------------------------------------------------
// The quadratic of :40-77 in the sphere's object space. Returns the hit distance and the object-space ray.
YK_DEV bool sphere_test(const yk_sphere& sp, V3 o_w, V3 d_w, float t_max, float* t_out, V3* o_obj, V3* d_obj) {
    const V3 o = xf_point(sp.world_to_object, o_w), d = xf_vec(sp.world_to_object, d_w);  // &world_to_object * ray, transform.rs:171-177
    const float a = d.x * d.x + d.y * d.y + d.z * d.z;
    const float b = 2.0f * (d.x * o.x + d.y * o.y + d.z * o.z);
    const float c = o.x * o.x + o.y * o.y + o.z * o.z - sp.radius * sp.radius;
    const float discrim = b * b - 4.0f * a * c;
    if (discrim < 0.0f) return false;
    const float rd = sqrtf(discrim);
    const float q = b < 0.0f ? -0.5f * (b - rd) : -0.5f * (b + rd);
    float t0 = q / a, t1 = c / q;
    if (t0 > t1) { const float tmp = t0; t0 = t1; t1 = tmp; }
    if (t0 > t_max || t1 <= 0.0f) return false;
    float t = t0;
    if (t <= 0.0f) {
        t = t1;
        if (t > t_max) return false;
    }
    *t_out = t;
    *o_obj = o;
    *d_obj = d;
    return true;
}
// Surface interaction of a sphere hit (:79-117) moved to world space by `&object_to_world * SurfaceInteraction`
// (interaction.rs:141-164). atan2 / acos restate glibc's routines (yk_libm.h) like sin / cos, so uv and the shading
// frame carry the bits the CPU path computes; the hit itself uses only + - * / sqrt.
static __device__ __noinline__ void sphere_surface(const yk_sphere& sp, V3 o_w, V3 d_w, Surface* si) {  // rare: out of line
    float t = 0.0f;
    V3 o = mk(0, 0, 0), d = mk(0, 0, 1);
    sphere_test(sp, o_w, d_w, __int_as_float(0x7f800000), &t, &o, &d);
    V3 p = o + d * t;
    p = p * (sp.radius / length(p - mk(0.0f, 0.0f, 0.0f)));
    if (p.x == 0.0f && p.y == 0.0f) p.x = 1e-5f * sp.radius;
    float phi = yklibm::atan2f_glibc(p.y, p.x);
    if (phi < 0.0f) phi += 2.0f * kPi;
    const float phi_max = 2.0f * kPi, theta_min = kPi, theta_max = 0.0f;
    const float u = phi / phi_max;
    const float theta = yklibm::acosf_glibc(clamp01ish(p.z / sp.radius, -1.0f, 1.0f));
    const float v = (theta - theta_min) / (theta_max - theta_min);
    const float z_radius = sqrtf(p.x * p.x + p.y * p.y);
    const float inv_z_radius = 1.0f / z_radius;
    const float cos_phi = p.x * inv_z_radius, sin_phi = p.y * inv_z_radius;
    const V3 dpdu = mk(-phi_max * p.y, phi_max * p.x, 0.0f);
    const V3 dpdv = mk(p.z * cos_phi, p.z * sin_phi, -sp.radius * sin_f32(theta)) * (theta_max - theta_min);
    V3 n = unit(cross64(dpdu, dpdv));  // SurfaceInteraction::new, interaction.rs:105-113
    if (sp.swaps_handedness) n = -n;
    // &object_to_world * si
    const V3 n_w = unit(xf_normal(sp.world_to_object, n));
    V3 sh_n = unit(xf_normal(sp.world_to_object, n));
    sh_n = flip_toward_n(sh_n, n_w);
    si->p = xf_point(sp.object_to_world, p);
    si->n = n_w;
    si->uv = {u, v};
    si->wo = unit(xf_vec(sp.object_to_world, -d_w));  // the world-space -ray.d goes through the transform again (sphere.rs:116)
    si->sh_n = flip_toward_n(sh_n, n_w);
    si->sh_dpdu = xf_vec(sp.object_to_world, dpdu);
    si->area_light = -1;
}

// ---- BSDF -----------------------------------------------------------------------------------------
enum : uint32_t { BX_REFLECTION = 1, BX_TRANSMISSION = 2, BX_DIFFUSE = 4, BX_GLOSSY = 8, BX_SPECULAR = 16, BX_ALL = 31 };

YK_DEV float cos2_theta(V3 w) { return w.z * w.z; }
YK_DEV float sin2_theta(V3 w) { return fmaxf(1.0f - cos2_theta(w), 0.0f); }
YK_DEV float sin_theta(V3 w) { return sqrtf(sin2_theta(w)); }
YK_DEV float tan_theta(V3 w) { return sin_theta(w) / w.z; }
YK_DEV float tan2_theta(V3 w) { return sin2_theta(w) / cos2_theta(w); }
YK_DEV float sin_phi(V3 w) { const float s = sin_theta(w); return s == 0.0f ? 1.0f : clamp01ish(w.y / s, -1.0f, 1.0f); }
YK_DEV float cos_phi(V3 w) { const float s = sin_theta(w); return s == 0.0f ? 1.0f : clamp01ish(w.x / s, -1.0f, 1.0f); }
YK_DEV bool same_hemi(V3 a, V3 b) { return a.z * b.z > 0.0f; }

// fresnel.rs:21-51
YK_DEV RGB fr_dielectric(float eta_i0, float eta_t0, float ci) {
    ci = clamp01ish(ci, -1.0f, 1.0f);
    const bool entering = ci > 0.0f;
    const float ei = entering ? eta_i0 : eta_t0, et = entering ? eta_t0 : eta_i0;
    if (!entering) ci = fabsf(ci);
    const float si = sqrtf(fmaxf(1.0f - ci * ci, 0.0f));
    const float st = ei / et * si;
    if (st >= 1.0f) return gray(1.0f);
    const float ct = sqrtf(fmaxf(1.0f - st * st, 0.0f));
    const float rpar = ((et * ci) - (ei * ct)) / ((et * ci) + (ei * ct));
    const float rper = ((ei * ci) - (et * ct)) / ((ei * ci) + (et * ct));
    return gray(1.0f) * (rpar * rpar + rper * rper) / 2.0f;
}
// fresnel.rs:68-96 with eta_i = 1
YK_DEV RGB fr_conductor(RGB eta_t, RGB k, float ci) {
    ci = fminf(fabsf(ci), 1.0f);
    const RGB one = gray(1.0f);
    const RGB eta = eta_t / one, eta_k = k / one;
    const float c2 = ci * ci, s2 = 1.0f - c2;
    const RGB eta2 = eta * eta, etak2 = eta_k * eta_k;
    const RGB t0 = eta2 - etak2 - gray(s2);
    const RGB a2b2 = rsqrt3(t0 * t0 + eta2 * etak2 * 4.0f);
    const RGB t1 = a2b2 + gray(c2);
    const RGB a = rsqrt3((a2b2 + t0) * 0.5f);
    const RGB t2 = a * ci * 2.0f;
    const RGB rs = (t1 - t2) / (t1 + t2);
    const RGB t3 = a2b2 * c2 + gray(s2 * s2);
    const RGB t4 = t2 * s2;
    const RGB rp = rs * (t3 - t4) / (t3 + t4);
    return (rp + rs) * 0.5f;
}
// fresnel.rs:108-117
YK_DEV RGB fr_schlick(RGB rs, float ci) {
    ci = clamp01ish(ci, -1.0f, 1.0f);
    const float v = 1.0f - ci;
    return rs + (gray(1.0f) - rs) * ((v * v) * (v * v) * v);
}
// trowbridge_reitz.rs:34-78
YK_DEV float ggx_d(float alpha, V3 wh) {
    const float t2 = tan2_theta(wh);
    if (isinf(t2)) return 0.0f;
    const float a2 = alpha * alpha;
    const float c4 = cos2_theta(wh) * cos2_theta(wh);
    const float cp = cos_phi(wh), sp = sin_phi(wh);
    const float e = ((cp * cp) / a2 + (sp * sp) / a2) * t2;
    return 1.0f / (kPi * a2 * c4 * (1.0f + e) * (1.0f + e));
}
YK_DEV float ggx_lambda(float alpha, V3 w) {
    const float at = fabsf(tan_theta(w));
    if (isinf(at)) return 0.0f;
    const float cp = cos_phi(w), sp = sin_phi(w);
    const float a = sqrtf((cp * cp) * alpha * alpha + (sp * sp) * alpha * alpha);
    const float q = (a * at) * (a * at);
    return (-1.0f + sqrtf(1.0f + q)) / 2.0f;
}
YK_DEV float ggx_g(float alpha, V3 wo, V3 wi) { return 1.0f / (1.0f + ggx_lambda(alpha, wo) + ggx_lambda(alpha, wi)); }
YK_DEV float ggx_pdf(float alpha, V3 wh) { return ggx_d(alpha, wh) * wh.z; }
YK_DEV V3 ggx_sample_wh(float alpha, V3 wo, V2 u) {
    const float tan2 = alpha * alpha * u.x / (1.0f - u.x);
    const float ct = 1.0f / sqrtf(1.0f + tan2);
    const float phi = 2.0f * kPi * u.y;
    const float st = sqrtf(fmaxf(1.0f - ct * ct, 0.0f));
    const V3 wh = mk(st * cos_f32(phi), st * sin_f32(phi), ct);
    return same_hemi(wo, wh) ? wh : -wh;
}

// One material's scattering functions, evaluated at a surface point (Material::compute_scattering_functions,
// materials/*.rs). `kind` is a compile-time constant in the per-material shading kernels.
struct Bsdf {
    uint32_t kind;      // yk_material_kind
    bool empty;         // Matte with black kd adds no lobe (matte.rs:31)
    RGB c0, c1;         // matte: kd | glass: R, T | metal: eta, k | glossy: rs
    float p0, p1;       // matte: Oren-Nayar A, B (p1 < 0 => Lambertian) | glass: eta | metal/glossy: alpha
    V3 ng, ns, ss, ts;  // Bsdf::new, bsdfs/mod.rs:87-99

    YK_DEV V3 to_local(V3 v) const { return {dot0(v, ss), dot0(v, ts), dotn(v, ns)}; }
    YK_DEV V3 to_world(V3 v) const {
        return {ss.x * v.x + ts.x * v.y + ns.x * v.z, ss.y * v.x + ts.y * v.y + ns.y * v.z, ss.z * v.x + ts.z * v.y + ns.z * v.z};
    }
    YK_DEV RGB fresnel(float c) const { return kind == YK_MAT_METAL ? fr_conductor(c0, c1, c) : fr_schlick(c0, c); }

    // lambertian.rs:21-23 / oren_nayar.rs:29-53 (first argument is what the trait passes as wo)
    YK_DEV RGB diffuse_f(V3 first, V3 second) const {
        if (p1 < 0.0f) return c0 * kInvPi;
        const float s_i = sin_theta(first), s_o = sin_theta(second);
        float max_cos = 0.0f;
        if (s_i > 1e-4f && s_o > 1e-4f) {
            const float d = cos_phi(first) * cos_phi(second) + sin_phi(first) * sin_phi(second);
            max_cos = fmaxf(d, 0.0f);
        }
        float sin_alpha, tan_beta;
        if (fabsf(first.z) > fabsf(second.z)) { sin_alpha = s_o; tan_beta = s_i / fabsf(first.z); }
        else { sin_alpha = s_i; tan_beta = s_o / fabsf(second.z); }
        return c0 * kInvPi * (p0 + p1 * max_cos * sin_alpha * tan_beta);
    }
    // microfacet.rs:53-74
    YK_DEV RGB microfacet_f(V3 wo, V3 wi) const {
        const float co = fabsf(wo.z), ci = fabsf(wi.z);
        if (ci == 0.0f || co == 0.0f) return gray(0.0f);
        V3 wh = wi + wo;
        if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return gray(0.0f);
        wh = unit(wh);
        const RGB fr = fresnel(dot0(wi, flip_toward(wh, mk(0.0f, 0.0f, 1.0f))));
        return gray(1.0f) * ggx_d(p0, wh) * ggx_g(p0, wo, wi) * fr / (4.0f * ci * co);
    }

    // Bsdf::f with BxdfType::all(), bsdfs/mod.rs:125-147
    YK_DEV RGB f(V3 wo_w, V3 wi_w) const {
        if (empty || kind == YK_MAT_GLASS) return gray(0.0f);  // specular lobes evaluate to zero (specular.rs:22,65)
        const bool reflect = dotn(wi_w, ng) * dotn(wo_w, ng) > 0.0f;
        if (!reflect) return gray(0.0f);  // only reflection lobes exist outside glass
        const V3 wo = to_local(wo_w), wi = to_local(wi_w);
        return gray(0.0f) + (kind == YK_MAT_MATTE ? diffuse_f(wo, wi) : microfacet_f(wo, wi));
    }

    struct Sample {
        V3 wi;
        RGB f;
        float pdf;
        uint32_t type;  // 0 = BxdfType::NONE
    };
    // specular.rs:26-37
    YK_DEV void sample_spec_reflect(V3 wo, Sample* s) const {
        const V3 wi = mk(-wo.x, -wo.y, wo.z);
        s->wi = wi;
        s->f = c0 * fr_dielectric(1.0f, p0, wi.z) / fabsf(wi.z);
        s->pdf = 1.0f;
        s->type = BX_SPECULAR | BX_REFLECTION;
    }
    // specular.rs:69-92 + refract, bsdfs/mod.rs:284-296
    YK_DEV void sample_spec_transmit(V3 wo, Sample* s) const {
        const bool entering = wo.z > 0.0f;
        const float ei = entering ? 1.0f : p0, et = entering ? p0 : 1.0f;
        const V3 n = flip_toward(mk(0.0f, 0.0f, 1.0f), wo);
        const float eta = ei / et;
        const float ci = dotn(n, wo);
        const float s2i = fmaxf(1.0f - ci * ci, 0.0f);
        const float s2t = eta * eta * s2i;
        if (s2t >= 1.0f) return;  // total internal reflection: default (NONE) sample
        const float ct = sqrtf(1.0f - s2t);
        const V3 wi = (-wo) * eta + n * (eta * ci - ct);
        s->wi = wi;
        s->f = c1 * (gray(1.0f) - fr_dielectric(1.0f, p0, wi.z)) / fabsf(wi.z);
        s->pdf = 1.0f;
        s->type = BX_SPECULAR | BX_TRANSMISSION;
    }
    // Bsdf::sample_f, bsdfs/mod.rs:150-222. `want` is BX_ALL (path) or BX_SPECULAR|{REFLECTION,TRANSMISSION} (whitted).
    YK_DEV Sample sample_f(V3 wo_w, V2 u, uint32_t want) const {
        Sample s{mk(0, 0, 0), gray(0.0f), 0.0f, 0u};
        if (empty) return s;
        const V3 wo = to_local(wo_w);
        if (kind == YK_MAT_GLASS) {
            const bool want_r = (want & (BX_SPECULAR | BX_REFLECTION)) == (BX_SPECULAR | BX_REFLECTION);
            const bool want_t = (want & (BX_SPECULAR | BX_TRANSMISSION)) == (BX_SPECULAR | BX_TRANSMISSION);
            const int matching = (want_r ? 1 : 0) + (want_t ? 1 : 0);
            if (matching == 0) return s;
            const float fl = floorf(u.x * (float)matching);
            int pick = fl > 0.0f ? (int)fl : 0;
            if (pick > matching - 1) pick = matching - 1;
            const bool do_reflect = want_r && pick == 0;
            if (do_reflect) sample_spec_reflect(wo, &s);
            else sample_spec_transmit(wo, &s);
            if (s.pdf == 0.0f) return Sample{mk(0, 0, 0), gray(0.0f), 0.0f, 0u};
            if (matching > 1) s.pdf /= (float)matching;
            s.wi = to_world(s.wi);
            return s;
        }
        const uint32_t mine = kind == YK_MAT_MATTE ? (BX_DIFFUSE | BX_REFLECTION) : (BX_REFLECTION | BX_GLOSSY);
        if ((want & mine) != mine) return s;
        // single lobe: comp = 0, u_remapped.x = u.x * (1 - 0) (bsdfs/mod.rs:176)
        const V2 ur{u.x * 1.0f, u.y};
        if (kind == YK_MAT_MATTE) {  // lambertian.rs:25-47 / oren_nayar.rs:55-77
            V3 wi = cosine_hemisphere(ur);
            if (wo.z < 0.0f) wi.z *= -1.0f;
            s.pdf = same_hemi(wo, wi) ? fabsf(wi.z) * kInvPi : 0.0f;
            s.f = diffuse_f(wo, wi);
            s.wi = wi;
        } else {  // microfacet.rs:76-99
            if (wo.z == 0.0f) return s;
            const V3 wh = ggx_sample_wh(p0, wo, ur);
            if (dot0(wo, wh) < 0.0f) return s;
            const V3 wi = (-wo) + wh * 2.0f * dot0(wo, wh);
            if (!same_hemi(wo, wi)) return s;
            s.pdf = ggx_pdf(p0, wh) / (4.0f * dot0(wo, wh));
            s.f = microfacet_f(wo, wi);
            s.wi = wi;
        }
        if (s.pdf == 0.0f) return Sample{mk(0, 0, 0), gray(0.0f), 0.0f, 0u};
        s.type = mine;
        s.wi = to_world(s.wi);
        return s;
    }
};

}  // namespace ykd
