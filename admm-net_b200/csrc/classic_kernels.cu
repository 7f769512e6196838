// Classical ADMM (reference admm.py:6-114) as evaluated for every reachable input (SURVEY.md §8a-9,
// App. A.3; oracle/classic_oracle.py::admm_linear_recursion):
//     M = diag(1/|b|^2) + rho 1 1^T ,  phi_0 = 0 ,  phi_k = M^{-1} (y/b + rho phi_{k-1})
// with M^{-1} v = D v - rho D 1 (1^T D v) / (1 + rho 1^T D 1),  D = diag(|b|^2)  (Sherman-Morrison).
// fp64 / complex128 like the reference.  One warp per signal; the kernel is HBM-bound
// (reads y,b, writes phi: 3 * 16 B * n per signal for complex128 I/O).
#include "common.cuh"

namespace admmnet {

template <typename CIn>
__device__ __forceinline__ double2 ld_c(const CIn* p, size_t i);
template <>
__device__ __forceinline__ double2 ld_c<double2>(const double2* p, size_t i) { return p[i]; }
template <>
__device__ __forceinline__ double2 ld_c<float2>(const float2* p, size_t i) {
    const float2 v = p[i];
    return make_double2((double)v.x, (double)v.y);
}

// One signal per group of GL lanes (ADMMNET_CLASSIC_GL = 8 | 16 | 32 selects; see the instantiations below).
// MAXE: elements per lane (n <= GL*MAXE).
template <typename CIn, int MAXE, int GL>
__global__ void __launch_bounds__(256)
k_classic(const CIn* __restrict__ y, const CIn* __restrict__ b, int B, int n, double rho, int n_iter,
          double2* __restrict__ phi_out) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int w = gtid / GL, lane = threadIdx.x % GL;
    const bool valid = w < B;
    const size_t base = (size_t)(valid ? w : 0) * n;
    double D[MAXE];
    double2 yb[MAXE], phi[MAXE];
    double dsum = 0.0;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int j = lane + GL * e;
        D[e] = 0.0;
        yb[e] = make_double2(0.0, 0.0);
        phi[e] = make_double2(0.0, 0.0);
        if (valid && j < n) {
            const double2 bj = ld_c<CIn>(b, base + j), yj = ld_c<CIn>(y, base + j);
            D[e] = bj.x * bj.x + bj.y * bj.y;                 // (b * conj(b)).real          admm.py:78
            const double rd = 1.0 / D[e];                     // inv(diag(b)) @ y = y * conj(b) / |b|^2
            yb[e] = make_double2((yj.x * bj.x + yj.y * bj.y) * rd, (yj.y * bj.x - yj.x * bj.y) * rd);
            dsum += D[e];
        }
    }
#pragma unroll
    for (int o = GL / 2; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    const double rden = rho / (1.0 + rho * dsum);
    for (int it = 0; it < n_iter; ++it) {
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            // v = y/b + rho*phi ; Dv = D*v
            phi[e].x = D[e] * fma(rho, phi[e].x, yb[e].x);
            phi[e].y = D[e] * fma(rho, phi[e].y, yb[e].y);
            sx += phi[e].x;
            sy += phi[e].y;
        }
#pragma unroll
        for (int o = GL / 2; o > 0; o >>= 1) {
            sx += __shfl_xor_sync(0xffffffffu, sx, o);
            sy += __shfl_xor_sync(0xffffffffu, sy, o);
        }
        const double fx = sx * rden, fy = sy * rden;           // rho * (1^T D v) / (1 + rho 1^T D 1)
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            phi[e].x = fma(-D[e], fx, phi[e].x);
            phi[e].y = fma(-D[e], fy, phi[e].y);
        }
    }
    if (!valid) return;
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int j = lane + GL * e;
        if (j < n) phi_out[base + j] = phi[e];
    }
}

template __global__ void k_classic<double2, 4, 32>(const double2*, const double2*, int, int, double, int, double2*);
template __global__ void k_classic<float2, 4, 32>(const float2*, const float2*, int, int, double, int, double2*);
template __global__ void k_classic<double2, 8, 32>(const double2*, const double2*, int, int, double, int, double2*);
template __global__ void k_classic<float2, 8, 32>(const float2*, const float2*, int, int, double, int, double2*);
// 8 lanes per signal (13 elements per lane, n <= 104): the per-iteration reduction is 3 shuffle steps shared by the
// 4 signals of a warp instead of 5 steps for one — the shuffle pipe (one warp instruction per clock per SM), not HBM,
// bounded the 32-lane form at n_iter = 5.
template __global__ void k_classic<double2, 13, 8>(const double2*, const double2*, int, int, double, int, double2*);
template __global__ void k_classic<float2, 13, 8>(const float2*, const float2*, int, int, double, int, double2*);
template __global__ void k_classic<double2, 7, 16>(const double2*, const double2*, int, int, double, int, double2*);
template __global__ void k_classic<float2, 7, 16>(const float2*, const float2*, int, int, double, int, double2*);

}  // namespace admmnet
