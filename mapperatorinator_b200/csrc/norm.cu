// LayerNorm over the last dimension, one warp per row, row cached in registers (two-pass mean / variance in fp32).
//   affine:   y = (x - mean) * rstd * w + b          (HF Whisper nn.LayerNorm, eps 1e-5)
//   modulate: y = (x - mean) * rstd * (1 + scale[b]) + shift[b]   (DiT adaLN, no affine, eps 1e-6; models.py:11-12,140)
#include "common.cuh"
#include "kernels.h"

namespace mb200 {
namespace {

template <int NV>  // dim == NV * 128
__global__ void __launch_bounds__(256) layernorm_kernel(LayerNormParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= p.rows) return;
    const float* x = p.x + (long long)warp * p.ldx;
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    const float inv_dim = 1.0f / (float)p.dim;
    const float mean = warp_sum(s) * inv_dim;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_dim + p.eps);
    float* y = p.y + (long long)warp * p.ldy;
    const float* shift = nullptr; const float* scale = nullptr;
    if (p.shift) {
        long long b = warp / p.rows_per_batch;
        shift = p.shift + b * p.mod_ld;
        scale = p.scale + b * p.mod_ld;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c = (i * 32 + lane) * 4;
        float4 o;
        o.x = (v[i].x - mean) * rstd; o.y = (v[i].y - mean) * rstd; o.z = (v[i].z - mean) * rstd; o.w = (v[i].w - mean) * rstd;
        if (p.weight) {
            float4 w = *reinterpret_cast<const float4*>(p.weight + c);
            o.x *= w.x; o.y *= w.y; o.z *= w.z; o.w *= w.w;
        }
        if (p.bias) {
            float4 b = *reinterpret_cast<const float4*>(p.bias + c);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
        }
        if (shift) {
            float4 sc = *reinterpret_cast<const float4*>(scale + c);
            float4 sh = *reinterpret_cast<const float4*>(shift + c);
            o.x = o.x * (1.f + sc.x) + sh.x; o.y = o.y * (1.f + sc.y) + sh.y;
            o.z = o.z * (1.f + sc.z) + sh.z; o.w = o.w * (1.f + sc.w) + sh.w;
        }
        *reinterpret_cast<float4*>(y + c) = o;
    }
}

}  // namespace

int launch_layernorm(const LayerNormParams& p, cudaStream_t stream) {
    MB_REQUIRE(p.dim % 128 == 0 && p.dim <= 1024, "layernorm dim must be a multiple of 128, <= 1024");
    MB_REQUIRE(p.ldx % 4 == 0 && p.ldy % 4 == 0 && p.mod_ld % 4 == 0, "layernorm strides must be multiples of 4");
    if (p.rows <= 0) return 0;
    const int threads = 256, rows_per_block = threads / 32;
    const int blocks = (p.rows + rows_per_block - 1) / rows_per_block;
    switch (p.dim / 128) {
#define MB_LN_CASE(n) case n: layernorm_kernel<n><<<blocks, threads, 0, stream>>>(p); break;
        MB_LN_CASE(1) MB_LN_CASE(2) MB_LN_CASE(3) MB_LN_CASE(4) MB_LN_CASE(5) MB_LN_CASE(6) MB_LN_CASE(7) MB_LN_CASE(8)
#undef MB_LN_CASE
    }
    MB_LAUNCH_CHECK();
    ++g_launch_count;
    return 0;
}

}  // namespace mb200
