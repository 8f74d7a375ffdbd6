// ORACLE — TEST INFRASTRUCTURE ONLY (see yko_math.h header).
//
// CPU restatement of yuki's seekable samplers (yuki/src/sampling/{mod,uniform,stratified}.rs) and
// of the third-party arithmetic they call, none of which is vendored under /root/reference:
//   * rand_pcg "0.3" (yuki/Cargo.toml:26): Pcg32 = Lcg64Xsh32 — new / next_u32 / advance,
//     restated from the published PCG algorithm (O'Neill 2014; pcg32 "XSH RR 64/32").
//   * rand "0.8" (yuki/Cargo.toml:25): Standard f32 = (next_u32() >> 8) * 2^-24.
//   * Rust std DefaultHasher::default() = SipHash-1-3 with k0 = k1 = 0 over the native-endian
//     bytes of the hashed integers (sampling/mod.rs:89-103).
// PARITY UNPINNED against the reference (it has no sampler tests); pinned instead against the
// published KATs: pcg32 (seed 42, stream 54) demo output and SipHash-1-3(empty, k=0) =
// Rust's DefaultHasher::new().finish() (tests/test_oracle_sampling.py).
#pragma once
#include <cstdint>

#include "yko_math.h"

namespace yko {

// --- SipHash-1-3, keys (0,0) --------------------------------------------------------------------
inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
struct SipState {
    uint64_t v0, v1, v2, v3;
    void round() {
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32);
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32);
    }
};
inline uint64_t siphash13(const uint8_t* msg, size_t n) {
    const uint64_t k0 = 0, k1 = 0;
    SipState s{k0 ^ 0x736f6d6570736575ULL, k1 ^ 0x646f72616e646f6dULL, k0 ^ 0x6c7967656e657261ULL,
               k1 ^ 0x7465646279746573ULL};
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        uint64_t m = 0;
        for (int b = 0; b < 8; ++b) m |= (uint64_t)msg[i + b] << (8 * b);
        s.v3 ^= m;
        s.round();  // c = 1
        s.v0 ^= m;
    }
    uint64_t b = (uint64_t)n << 56;
    for (int k = 0; i < n; ++i, ++k) b |= (uint64_t)msg[i] << (8 * k);
    s.v3 ^= b;
    s.round();
    s.v0 ^= b;
    s.v2 ^= 0xff;
    s.round(); s.round(); s.round();  // d = 3
    return s.v0 ^ s.v1 ^ s.v2 ^ s.v3;
}
// hash_values!(pixel) — Point2<u16> derives Hash: x then y (math/point.rs:40-63)
inline uint64_t hash_pixel(uint16_t px, uint16_t py) {
    uint8_t m[4] = {(uint8_t)(px & 0xff), (uint8_t)(px >> 8), (uint8_t)(py & 0xff), (uint8_t)(py >> 8)};
    return siphash13(m, 4);
}
// hash_values!(pixel, dimension: u32, rng_seed: u64) (stratified.rs:105,122)
inline uint64_t hash_pixel_dim_seed(uint16_t px, uint16_t py, uint32_t dim, uint64_t seed) {
    uint8_t m[16];
    m[0] = px & 0xff; m[1] = px >> 8; m[2] = py & 0xff; m[3] = py >> 8;
    for (int i = 0; i < 4; ++i) m[4 + i] = (uint8_t)(dim >> (8 * i));
    for (int i = 0; i < 8; ++i) m[8 + i] = (uint8_t)(seed >> (8 * i));
    return siphash13(m, 16);
}

// --- PCG32 (rand_pcg::Lcg64Xsh32) ---------------------------------------------------------------
struct Pcg32 {
    uint64_t state, inc;
    static constexpr uint64_t MULT = 6364136223846793005ULL;
    static Pcg32 make(uint64_t state, uint64_t stream) {
        Pcg32 p{state, (stream << 1) | 1};
        p.state = p.state + p.inc;
        p.step();
        return p;
    }
    void step() { state = state * MULT + inc; }
    uint32_t next_u32() {
        uint64_t old = state;
        step();
        uint32_t rot = (uint32_t)(old >> 59);
        uint32_t xsh = (uint32_t)(((old >> 18) ^ old) >> 27);
        return (xsh >> rot) | (xsh << ((32 - rot) & 31));
    }
    void advance(uint64_t delta) {
        uint64_t acc_mult = 1, acc_plus = 0, cur_mult = MULT, cur_plus = inc;
        while (delta > 0) {
            if (delta & 1) {
                acc_mult *= cur_mult;
                acc_plus = acc_plus * cur_mult + cur_plus;
            }
            cur_plus = (cur_mult + 1) * cur_plus;
            cur_mult *= cur_mult;
            delta /= 2;
        }
        state = acc_mult * state + acc_plus;
    }
    // rand 0.8 Standard for f32: 24 high bits scaled by 2^-24, in [0, 1)
    float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};

// stratified.rs:147-178
inline uint32_t permutation_element(uint32_t i, uint32_t l, uint32_t p) {
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p;             i *= 0xe170893du;
        i ^= p >> 16;
        i ^= (i & w) >> 4;
        i ^= p >> 8;        i *= 0x0929eb3fu;
        i ^= p >> 23;
        i ^= (i & w) >> 1;  i *= 1u | p >> 27;
                            i *= 0x6935fa69u;
        i ^= (i & w) >> 11; i *= 0x74dcb303u;
        i ^= (i & w) >> 2;  i *= 0x9e501cc3u;
        i ^= (i & w) >> 2;  i *= 0xc860a3dfu;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (i + p) % l;
}

// --- Sampler (sampling/mod.rs:46-57) -------------------------------------------------------------
enum SamplerKind : uint32_t { SAMPLER_UNIFORM = 0, SAMPLER_STRATIFIED = 1 };

struct Sampler {
    SamplerKind kind;
    uint32_t nx, ny;      // stratified pixel_samples; uniform: nx = pixel_samples, ny = 1
    bool jitter;
    uint64_t seed;        // explicit here; the reference draws it from thread_rng (uniform.rs:37)
    uint16_t px = 0, py = 0;
    uint32_t sample_index = 0, dimension = 0;
    Pcg32 rng{0, 1};

    uint32_t samples_per_pixel() const { return kind == SAMPLER_UNIFORM ? nx : nx * ny; }

    // uniform.rs:72-84 / stratified.rs:90-102 (the stratified one zeroes `dimension`)
    void start_pixel_sample(uint16_t x, uint16_t y, uint32_t index, uint32_t dim) {
        px = x; py = y; sample_index = index;
        dimension = kind == SAMPLER_UNIFORM ? dim : 0;
        rng = Pcg32::make(seed, hash_pixel(px, py));
        rng.advance((uint64_t)index * 65536ULL + (uint64_t)dim);
    }
    float get_1d() {
        if (kind == SAMPLER_UNIFORM) {  // uniform.rs:86-89
            dimension += 1;
            return rng.next_f32();
        }
        // stratified.rs:104-119
        uint64_t h = hash_pixel_dim_seed(px, py, dimension, seed);
        uint32_t stratum = permutation_element(sample_index, samples_per_pixel(), (uint32_t)h);
        dimension += 1;
        float delta = jitter ? rng.next_f32() : 0.5f;
        return ((float)stratum + delta) / (float)samples_per_pixel();
    }
    V2 get_2d() {
        if (kind == SAMPLER_UNIFORM) {  // uniform.rs:91-94, x drawn first
            dimension += 2;
            float x = rng.next_f32();
            float y = rng.next_f32();
            return {x, y};
        }
        // stratified.rs:121-143 — note y = stratum / pixel_samples.y (reference quirk)
        uint64_t h = hash_pixel_dim_seed(px, py, dimension, seed);
        uint32_t stratum = permutation_element(sample_index, samples_per_pixel(), (uint32_t)h);
        dimension += 2;
        uint32_t x = stratum % nx;
        uint32_t y = stratum / ny;
        float dx = jitter ? rng.next_f32() : 0.5f;
        float dy = jitter ? rng.next_f32() : 0.5f;
        return {((float)x + dx) / (float)nx, ((float)y + dy) / (float)ny};
    }
};

// sampling/mod.rs:68-87
inline V2 concentric_sample_disk(V2 u) {
    V2 off = u * 2.0f - V2{1.0f, 1.0f};
    if (off.x == 0.0f && off.y == 0.0f) return {0.0f, 0.0f};
    float theta, r;
    if (std::fabs(off.x) > std::fabs(off.y)) {
        theta = FRAC_PI_4_F * (off.y / off.x);
        r = off.x;
    } else {
        theta = FRAC_PI_2_F - FRAC_PI_4_F * (off.x / off.y);
        r = off.y;
    }
    return V2{std::cos(theta), std::sin(theta)} * r;
}
// sampling/mod.rs:62-66
inline V3 cosine_sample_hemisphere(V2 u) {
    V2 d = concentric_sample_disk(u);
    float z = std::sqrt(fmax_(1.0f - d.x * d.x - d.y * d.y, 0.0f));
    return {d.x, d.y, z};
}

}  // namespace yko
