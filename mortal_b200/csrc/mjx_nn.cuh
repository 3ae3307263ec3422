// mortal_b200 — fused elementwise kernels for the policy-net inference path (mortal/model.py ResBlock / ChannelAttention,
// restated in mortal_b200/model.py). The convolutions stay with cuDNN (tcgen05 implicit-GEMM kernels); what PyTorch runs
// between them as 5-6 separate bandwidth-bound passes per block (BatchNorm affine, Mish, two pooling reductions, gate
// multiply, residual add) is done here in three: one 16-byte vector of 8 bf16 channels per thread, NHWC
// (channels-last) activations [B, L, C], fp32 math, one rounding to bf16 at the end.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mjx_nn {

struct alignas(16) Vec8 { __nv_bfloat162 v[4]; };

__device__ __forceinline__ float mish_f(float x) {
    // x * tanh(softplus(x)); softplus saturates to x for large x exactly as log1p(exp(x)) does in fp32
    const float sp = x > 20.f ? x : log1pf(expf(x));
    return x * tanhf(sp);
}

// out = mish(x * scale[c] + bias[c])
__global__ void __launch_bounds__(256) k_affine_mish(const Vec8* __restrict__ x, const float* __restrict__ scale,
                                                     const float* __restrict__ bias, Vec8* __restrict__ out, size_t n_vec, int c8) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (size_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(v % (size_t)c8) * 8;
        const Vec8 in = x[v];
        Vec8 o;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 f = __bfloat1622float2(in.v[k]);
            const float a = mish_f(fmaf(f.x, __ldg(scale + c0 + 2 * k), __ldg(bias + c0 + 2 * k)));
            const float b = mish_f(fmaf(f.y, __ldg(scale + c0 + 2 * k + 1), __ldg(bias + c0 + 2 * k + 1)));
            o.v[k] = __floats2bfloat162_rn(a, b);
        }
        out[v] = o;
    }
}

// avg[b, c] = mean_l x[b, l, c], mx[b, c] = max_l x[b, l, c]   (ChannelAttention pooling)
__global__ void __launch_bounds__(256) k_pool(const Vec8* __restrict__ x, Vec8* __restrict__ avg, Vec8* __restrict__ mx,
                                              int batch, int length, int c8) {
    const size_t n = (size_t)batch * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / c8;
        const int cv = (int)(i - b * c8);
        float s[8], m[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { s[k] = 0.f; m[k] = -3.402823466e+38f; }
        const Vec8* p = x + b * (size_t)length * c8 + cv;
        for (int l = 0; l < length; l++) {
            const Vec8 in = p[(size_t)l * c8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 f = __bfloat1622float2(in.v[k]);
                s[2 * k] += f.x; s[2 * k + 1] += f.y;
                m[2 * k] = fmaxf(m[2 * k], f.x); m[2 * k + 1] = fmaxf(m[2 * k + 1], f.y);
            }
        }
        Vec8 oa, om;
        const float inv = 1.f / (float)length;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            oa.v[k] = __floats2bfloat162_rn(s[2 * k] * inv, s[2 * k + 1] * inv);
            om.v[k] = __floats2bfloat162_rn(m[2 * k], m[2 * k + 1]);
        }
        avg[i] = oa;
        mx[i] = om;
    }
}

// out = y * gate[b, c] + x   (channel gate + residual)
__global__ void __launch_bounds__(256) k_gate_residual(const Vec8* __restrict__ y, const Vec8* __restrict__ gate,
                                                       const Vec8* __restrict__ x, Vec8* __restrict__ out, size_t n_vec,
                                                       int length, int c8) {
    const size_t per_b = (size_t)length * c8;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < n_vec; v += (size_t)gridDim.x * blockDim.x) {
        const size_t b = v / per_b;
        const int cv = (int)(v % (size_t)c8);
        const Vec8 yy = y[v], xx = x[v], g = gate[b * c8 + cv];
        Vec8 o;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float2 fy = __bfloat1622float2(yy.v[k]), fx = __bfloat1622float2(xx.v[k]), fg = __bfloat1622float2(g.v[k]);
            o.v[k] = __floats2bfloat162_rn(fmaf(fy.x, fg.x, fx.x), fmaf(fy.y, fg.y, fx.y));
        }
        out[v] = o;
    }
}

}  // namespace mjx_nn
