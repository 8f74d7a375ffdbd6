// The two display passes that follow the hot path in yuki (the step after Film::update_tile): the filmic tone map with
// per-tile sample-count normalisation and the B->G->R heat map of the BVH-intersection film, written as CUDA kernels
// instead of the reference's GLSL fragment shaders (yuki/src/app/renderpasses/tonemap.rs:318-418) plus the min/max
// search of `find_min_max` (tonemap.rs:447-472). Pixel coordinates are film coordinates (row-major, y down); the tile
// of a pixel is (x / tile_dim, y / tile_dim) with the shader's x_tile_count = res.x / tile_dim (tonemap.rs:386).
// Both shader quirks are kept: the heat map reads luminance for channel 0 (`channel > 0 && channel < 3`, :412), while
// the min/max search reads the red channel for it (:455).
#include <cuda_runtime.h>

#include <cfloat>
#include <cstring>
#include <string>

#include "yuki_gpu.h"
#include "yk_guard.h"

int yk_set_error(int code, const std::string& msg);  // host_scene.cpp
int yk_context_activate(yk_context* c);                 // render.cu: cudaSetDevice(the context's device)

namespace {

#define POST_TRY(expr)                                                                                \
    do {                                                                                              \
        cudaError_t e_ = (expr);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            cudaFree(d_in); cudaFree(d_out); cudaFree(d_aux);                                         \
            return yk_set_error(YK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
        }                                                                                             \
    } while (0)

__device__ __forceinline__ float sat(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
__device__ __forceinline__ float rrt_odt(float v) {  // tonemap.rs:353-358
    const float a = v * (v + 0.0245786f) - 0.000090537f;
    const float b = v * (0.983729f * v + 0.4329510f) + 0.238081f;
    return a / b;
}

__global__ void k_tonemap_filmic(const float* film, uint32_t res_x, uint32_t res_y, const float* tile_samples, uint32_t n_tiles,
                                 uint32_t tile_dim, float exposure, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= res_x * res_y) return;
    const uint32_t x = i % res_x, y = i / res_x;
    float r = film[3 * (size_t)i], g = film[3 * (size_t)i + 1], b = film[3 * (size_t)i + 2];
    if (tile_samples) {  // tonemap.rs:383-392
        const uint32_t x_tile_count = res_x / tile_dim;
        const uint32_t flat = (y / tile_dim) * x_tile_count + x / tile_dim;
        const float n = flat < n_tiles ? tile_samples[flat] : 0.0f;
        if (n > 0.0f) { r /= n; g /= n; b /= n; }
    }
    r *= exposure; g *= exposure; b *= exposure;
    // ACESFitted, tonemap.rs:338-375 (row-major matrices as written there)
    const float ir = 0.59719f * r + 0.35458f * g + 0.04823f * b;
    const float ig = 0.07600f * r + 0.90834f * g + 0.01566f * b;
    const float ib = 0.02840f * r + 0.13383f * g + 0.83777f * b;
    const float fr = rrt_odt(ir), fg = rrt_odt(ig), fb = rrt_odt(ib);
    out[3 * (size_t)i] = sat(1.60475f * fr + -0.53108f * fg + -0.07367f * fb);
    out[3 * (size_t)i + 1] = sat(-0.10208f * fr + 1.10813f * fg + -0.00605f * fb);
    out[3 * (size_t)i + 2] = sat(-0.00327f * fr + -0.07276f * fg + 1.07602f * fb);
}

__device__ __forceinline__ float luminance(float r, float g, float b) { return 0.2126f * r + 0.7152f * g + 0.0722f * b; }

// find_min_max, tonemap.rs:447-472: fold over pixels with f32::min / f32::max from (f32::MAX, f32::MIN)
__global__ void k_min_max(const float* film, uint32_t n_pixels, uint32_t channel, float* min_max) {
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pixels; i += gridDim.x * blockDim.x) {
        const float r = film[3 * (size_t)i], g = film[3 * (size_t)i + 1], b = film[3 * (size_t)i + 2];
        const float v = channel == 0 ? r : (channel == 1 ? g : (channel == 2 ? b : luminance(r, g, b)));
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_down_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_down_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {  // order-independent: min/max of floats (no NaN results: fminf/fmaxf ignore NaN like f32::min/max)
        atomicMin((int*)&min_max[0], __float_as_int(lo) >= 0 ? __float_as_int(lo) : (int)(0x80000000u - (uint32_t)__float_as_int(lo)));
        atomicMax((int*)&min_max[1], __float_as_int(hi) >= 0 ? __float_as_int(hi) : (int)(0x80000000u - (uint32_t)__float_as_int(hi)));
    }
}
__device__ __forceinline__ float ordered_to_float(int k) { return __int_as_float(k >= 0 ? k : (int)(0x80000000u - (uint32_t)k)); }

__global__ void k_heatmap(const float* film, uint32_t n_pixels, uint32_t channel, const float* range_ordered, float min_val, float max_val,
                          float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    if (range_ordered) {
        min_val = ordered_to_float(((const int*)range_ordered)[0]);
        max_val = ordered_to_float(((const int*)range_ordered)[1]);
    }
    const float r = film[3 * (size_t)i], g = film[3 * (size_t)i + 1], b = film[3 * (size_t)i + 2];
    float value;
    if (channel > 0 && channel < 3) value = channel == 1 ? g : b;  // tonemap.rs:412-417 (channel 0 falls through to luminance)
    else value = luminance(r, g, b);
    const float s = (value - min_val) / (max_val - min_val);
    // mix(mix(LOW, MID, saturate(2s)), HIGH, saturate(2s - 1)) with LOW = blue, MID = green, HIGH = red
    const float t0 = sat(s * 2.0f), t1 = sat(s * 2.0f - 1.0f);
    const float m_r = 0.0f, m_g = t0, m_b = 1.0f - t0;  // mix(x, y, a) = x * (1 - a) + y * a
    out[3 * (size_t)i] = m_r * (1.0f - t1) + t1;
    out[3 * (size_t)i + 1] = m_g * (1.0f - t1);
    out[3 * (size_t)i + 2] = m_b * (1.0f - t1);
}

}  // namespace

extern "C" {

int yk_tonemap_filmic(yk_context* c, const float* film_rgb, uint32_t res_x, uint32_t res_y, const float* tile_samples, uint32_t n_tiles,
                      uint32_t tile_dim, float exposure, float* out_rgb) {
    return yk_guard("yk_tonemap_filmic", [&]() -> int {
    if (!c || !film_rgb || !out_rgb || !res_x || !res_y) return yk_set_error(YK_ERR_INVALID, "yk_tonemap_filmic: null / empty argument");
    if (tile_samples && !tile_dim) return yk_set_error(YK_ERR_INVALID, "yk_tonemap_filmic: tile_dim is zero");
    if (int rc = yk_context_activate(c)) return rc;
    cudaStream_t s = (cudaStream_t)yk_context_stream(c);
    const size_t n = (size_t)res_x * res_y, bytes = n * 3 * sizeof(float);
    float *d_in = nullptr, *d_out = nullptr, *d_aux = nullptr;
    POST_TRY(cudaMalloc((void**)&d_in, bytes));
    POST_TRY(cudaMalloc((void**)&d_out, bytes));
    POST_TRY(cudaMemcpyAsync(d_in, film_rgb, bytes, cudaMemcpyHostToDevice, s));
    if (tile_samples) {
        POST_TRY(cudaMalloc((void**)&d_aux, (size_t)n_tiles * sizeof(float)));
        POST_TRY(cudaMemcpyAsync(d_aux, tile_samples, (size_t)n_tiles * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    k_tonemap_filmic<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_in, res_x, res_y, d_aux, n_tiles, tile_dim, exposure, d_out);
    POST_TRY(cudaGetLastError());
    POST_TRY(cudaMemcpyAsync(out_rgb, d_out, bytes, cudaMemcpyDeviceToHost, s));
    POST_TRY(cudaStreamSynchronize(s));
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_aux);
    return YK_OK;
    });
}

int yk_heatmap(yk_context* c, const float* film_rgb, uint32_t res_x, uint32_t res_y, uint32_t channel, int auto_range, float* min_val,
               float* max_val, float* out_rgb) {
    return yk_guard("yk_heatmap", [&]() -> int {
    if (!c || !film_rgb || !out_rgb || !res_x || !res_y || !min_val || !max_val)
        return yk_set_error(YK_ERR_INVALID, "yk_heatmap: null / empty argument");
    if (channel > 3) return yk_set_error(YK_ERR_INVALID, "yk_heatmap: channel must be 0..3 (R, G, B, luminance)");
    if (int rc = yk_context_activate(c)) return rc;
    cudaStream_t s = (cudaStream_t)yk_context_stream(c);
    const size_t n = (size_t)res_x * res_y, bytes = n * 3 * sizeof(float);
    float *d_in = nullptr, *d_out = nullptr, *d_aux = nullptr;
    POST_TRY(cudaMalloc((void**)&d_in, bytes));
    POST_TRY(cudaMalloc((void**)&d_out, bytes));
    POST_TRY(cudaMemcpyAsync(d_in, film_rgb, bytes, cudaMemcpyHostToDevice, s));
    if (auto_range) {
        POST_TRY(cudaMalloc((void**)&d_aux, 2 * sizeof(float)));
        const int init[2] = {0x7f7fffff, (int)(0x80000000u - 0xff7fffffu)};  // ordered keys of f32::MAX, f32::MIN
        POST_TRY(cudaMemcpyAsync(d_aux, init, sizeof(init), cudaMemcpyHostToDevice, s));
        k_min_max<<<296, 256, 0, s>>>(d_in, (uint32_t)n, channel, d_aux);
    }
    k_heatmap<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_in, (uint32_t)n, channel, d_aux, *min_val, *max_val, d_out);
    POST_TRY(cudaGetLastError());
    POST_TRY(cudaMemcpyAsync(out_rgb, d_out, bytes, cudaMemcpyDeviceToHost, s));
    int keys[2] = {0, 0};
    if (auto_range) POST_TRY(cudaMemcpyAsync(keys, d_aux, sizeof(keys), cudaMemcpyDeviceToHost, s));
    POST_TRY(cudaStreamSynchronize(s));
    if (auto_range) {
        auto back = [](int k) { int v = k >= 0 ? k : (int)(0x80000000u - (uint32_t)k); float f; memcpy(&f, &v, 4); return f; };
        *min_val = back(keys[0]);
        *max_val = back(keys[1]);
    }
    cudaFree(d_in); cudaFree(d_out); cudaFree(d_aux);
    return YK_OK;
    });
}

}  // extern "C"
