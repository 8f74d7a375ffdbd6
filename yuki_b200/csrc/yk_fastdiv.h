// Division by a launch-invariant u32 (Granlund-Montgomery round-up method; exact for every 32-bit numerator).
// Used for the per-path index arithmetic the reference does with `/` and `%` on values that are constant for a whole
// render: samples per pixel (stratified.rs:127-128, 177), pixels per batch.
#pragma once
#include <stdint.h>

struct FastDiv {
    uint32_t d, m, s1, s2;
#if defined(__CUDACC__)
    __host__ __device__
#endif
    static FastDiv make(uint32_t d) {
        FastDiv f;
        f.d = d ? d : 1u;
        uint32_t l = 0;
        while ((1ull << l) < f.d) ++l;  // ceil(log2 d)
        f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - f.d)) / f.d + 1ull);
        f.s1 = l < 1u ? l : 1u;
        f.s2 = l > 0u ? l - 1u : 0u;
        return f;
    }
#if defined(__CUDACC__)
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        const uint32_t t = __umulhi(m, n);
        return (t + ((n - t) >> s1)) >> s2;
    }
    __device__ __forceinline__ uint32_t mod(uint32_t n) const { return n - div(n) * d; }
#endif
    uint32_t div_host(uint32_t n) const {
        const uint32_t t = (uint32_t)(((uint64_t)m * n) >> 32);
        return (t + ((n - t) >> s1)) >> s2;
    }
};

