// Shared device helpers for the admm-net B200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define ADMM_EPS 1e-8f

// Bounds / invariant checks for debug builds (python admm-net_b200/build.py --debug  ->  -DADMMNET_DEBUG): the GPU
// pool has compute-sanitizer closed, so index arithmetic of the shared-memory tilings is checked in-kernel instead.
// A failing check prints file:line and traps (the launch then reports an error instead of corrupting memory).
#ifdef ADMMNET_DEBUG
#include <cstdio>
#define ADMM_ASSERT(cond)                                                                              \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            printf("ADMM_ASSERT failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                                 \
            __trap();                                                                                  \
        }                                                                                              \
    } while (0)
#else
#define ADMM_ASSERT(cond) ((void)0)
#endif

namespace admmnet {

// ---------------------------------------------------------------- packed parameter layout (floats)
// One record per layer, written by admm-net_b200/params.py::pack_layer (must stay in sync).
enum ParamOff : int {
    P_RHO_PHI = 0,      // softplus(phiLayers.k.rho)                      admm_net.py:97
    P_RHO_H_EPS = 1,    // softplus(hLayers.k.rho) + eps                  admm_net.py:148,151
    P_SIG_PW = 2,       // sigmoid(hLayers.k.projection_weight)           admm_net.py:188
    P_C0 = 3,           // 1/(softplus(gLayers.k.lambda_param)^2+eps)     admm_net.py:269-271
    P_INV_RHO_G = 4,    // 1/(softplus(gLayers.k.rho)+eps)                admm_net.py:287-288
    P_THR = 5,          // sigmoid(gLayers.k.threshold)                   admm_net.py:321
    P_C1Z = 6,          // 1/(softplus(zLayers.k.lambda_param)^2+eps)     admm_net.py:424-426
    P_RHO_Z = 7,        // softplus(zLayers.k.rho)                        admm_net.py:406
    P_KNORM = 8,        // k/10                                           admm_net.py:457
    P_ZD2 = 9,          // residual_scale_net.2.bias
    P_VC2 = 10,         // value_net.2.bias
    P_V1 = 16,          // value_net.0.weight[16]
    P_VC1 = 32,         // value_net.0.bias[16]
    P_V2 = 48,          // value_net.2.weight[16]
    P_ZU1 = 64,         // residual_scale_net.0.weight [32][3]
    P_ZD1 = 160,        // residual_scale_net.0.bias[32]
    P_ZU2 = 192,        // residual_scale_net.2.weight[32]
    P_HB1 = 224,        // correction_net.0.bias[64]
    P_HW1T = 288,       // correction_net.0.weight transposed  [n][64]
    // P_HW2T = 288 + 64 n   correction_net.2.weight transposed [64][n]
    // P_HB2  = 288 + 128 n  correction_net.2.bias[n]
};
__host__ __device__ inline int param_stride(int n) { return (288 + 129 * n + 3) & ~3; }

// ---------------------------------------------------------------- small math
__device__ __forceinline__ float softplusf(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ float2 cconjmul(float2 a, float2 b) {  // conj(a) * b
    return make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cscale(float s, float2 a) { return make_float2(s * a.x, s * a.y); }
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__device__ __forceinline__ float2 cdiv(float2 a, float2 b) {
    float den = b.x * b.x + b.y * b.y;
    return make_float2((a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of up to 3 floats; `red` is >= 3*32 floats of shared scratch. All threads get the result.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[i * 32 + wid] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float x = lane < nw ? red[i * 32 + lane] : 0.f;
        v[i] = warp_sum(x);
    }
}
__device__ __forceinline__ float block_max(float v, float* red) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float x = lane < nw ? red[lane] : -INFINITY;
    return warp_max(x);
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2: two fp32 lanes per issue slot, same rounding as two FFMA).
// ptxas folds pk2(s, s) into a scalar-broadcast operand and lane swaps / sign changes into operand modifiers, so a
// complex multiply-add costs two issue slots instead of four.  Lane 0 = .x ("lo"), lane 1 = .y ("hi").
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 upk2(f32x2 v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 bc2(float s) { return pk2(s, s); }

// packed lower-triangular (row-major) index, j <= i
__host__ __device__ __forceinline__ int pk(int i, int j) { return (i * (i + 1)) / 2 + j; }
// Householder-vector store: column k holds rows k+1..d-1 (v[k+1] == 1 stored explicitly), columns concatenated
__host__ __device__ __forceinline__ int voff(int k, int d) { return k * (d - 1) - (k * (k - 1)) / 2; }

// adaptive dual step alpha (admm_net.py:443-474) for one signal
__device__ __forceinline__ float z_alpha(const float* __restrict__ P, float r, float mean_r) {
    const float rn = r / (mean_r + ADMM_EPS);
    const float f0 = P[P_KNORM], f1 = P[P_RHO_Z];
    float acc = P[P_ZD2];
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
        float hv = P[P_ZU1 + 3 * j] * f0 + P[P_ZU1 + 3 * j + 1] * f1 + P[P_ZU1 + 3 * j + 2] * rn + P[P_ZD1 + j];
        acc += P[P_ZU2 + j] * fmaxf(hv, 0.f);
    }
    const float sf = 0.5f + 1.5f * sigmoidf_(acc);
    return P[P_RHO_Z] * sf;
}

// eigenvalue map (admm_net.py:310-334)
__device__ __forceinline__ float eig_map(const float* __restrict__ P, float lam) {
    const float base = softplusf(lam - P[P_THR]);
    const float a = fabsf(lam);
    float acc = P[P_VC2];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += P[P_V2 + j] * fmaxf(P[P_V1 + j] * a + P[P_VC1 + j], 0.f);
    return base * sigmoidf_(acc);
}

}  // namespace admmnet
