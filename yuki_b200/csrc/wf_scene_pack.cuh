// Scene packing kernels of yk_scene_create: reference-layout arrays -> two-child node records and transposed triangles.
// Part of the single translation unit render.cu (compiled --fmad=false: every float op is the reference's un-fused IEEE op).
#pragma once
#include "wf_common.cuh"

namespace {

// ---- scene packing (yk_scene_create): the reference-layout arrays are repacked on the device ---------------------------
__global__ void k_scene_interior_flags(const yk_bvh_node* nodes, uint32_t n, uint32_t* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = nodes[i].is_leaf ? 0u : 1u;
}
// rec[i] = number of interior nodes before node i: an interior node's record index; i - rec[i] = a leaf's table index.
__global__ void k_scene_records(const yk_bvh_node* nodes, const uint32_t* rec, uint32_t n, int packed_leaves, uint2* leaf_table, float4* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const yk_bvh_node nd = nodes[i];
    if (nd.is_leaf) return;
    const uint32_t kids[2] = {i + 1, nd.offset};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const yk_bvh_node ch = nodes[kids[k]];
        uint32_t ref;
        if (!ch.is_leaf) ref = kRefInterior | ((uint32_t)ch.split_axis << 29) | rec[kids[k]];
        else if (packed_leaves) ref = ((uint32_t)(ch.shape_count - 1) << kLeafFirstBits) | ch.offset;
        else {
            ref = kids[k] - rec[kids[k]];
            leaf_table[ref] = make_uint2(ch.offset, ch.shape_count);
        }
        out[(size_t)rec[i] * 4 + 2 * k] = make_float4(ch.p_min[0], ch.p_min[1], ch.p_min[2], __uint_as_float(ref));
        out[(size_t)rec[i] * 4 + 2 * k + 1] = make_float4(ch.p_max[0], ch.p_max[1], ch.p_max[2], 0.0f);
    }
}
__global__ void k_scene_tris(const float* verts, const uint32_t* orig, const uint32_t* mat, const int32_t* alight, const uint8_t* flags,
                             const int32_t* sphere, const uint8_t* mat_kind, uint32_t n, float4* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t m = mat[i], f = flags[i];
    // material index | YK_TRI_* flags << 24 | material kind << 28 (so the material sort needs no material table look-up)
    const uint32_t packed = m | ((f & 0xfu) << 24) | (((uint32_t)mat_kind[m] & 3u) << 28);
    const float fp = __uint_as_float(packed), fi = __uint_as_float(orig[i]);
    if (f & YK_TRI_IS_SPHERE) {  // NaN vertex lanes + (-2 - sphere index) where triangles keep their area light
        const float qnan = __int_as_float(0x7fc00000);
        out[3 * (size_t)i] = make_float4(qnan, qnan, qnan, __int_as_float(-2 - sphere[i]));
        out[3 * (size_t)i + 1] = make_float4(qnan, qnan, qnan, fp);
        out[3 * (size_t)i + 2] = make_float4(qnan, qnan, qnan, fi);
        return;
    }
    const float* v = verts + (size_t)i * 9;
    out[3 * (size_t)i] = make_float4(v[0], v[3], v[6], __int_as_float(alight[i]));
    out[3 * (size_t)i + 1] = make_float4(v[1], v[4], v[7], fp);
    out[3 * (size_t)i + 2] = make_float4(v[2], v[5], v[8], fi);
}

}  // namespace
