// Classical ADMM (reference admm.py:6-114) as evaluated for every reachable input (SURVEY.md §8a-9,
// App. A.3; oracle/classic_oracle.py::admm_linear_recursion):
//     M = diag(1/|b|^2) + rho 1 1^T ,  phi_0 = 0 ,  phi_k = M^{-1} (y/b + rho phi_{k-1})
// with M^{-1} v = D v - rho D 1 (1^T D v) / (1 + rho 1^T D 1),  D = diag(|b|^2)  (Sherman-Morrison).
// fp64 / complex128 like the reference.  D (y/b) = y conj(b) exactly, so the recursion is carried on
//     Dv_k = y conj(b) + (rho D) phi_{k-1},   phi_k = Dv_k - D f_k,   f_k = rho sum(Dv_k) / (1 + rho sum(D)) :
// no division per element and 6 fp64 operations per element and iteration (a zero b_j, where the reference divides by
// zero, gives NaN here too).
//
// Two forms:
//   k_classic    one signal per group of GL lanes, operands straight from global memory (any n <= 256, c64 or c128
//                input);
//   k_classic_p  persistent, one CTA per SM: tiles of 32 signals stream through shared memory with bulk asynchronous
//                copies (cp.async.bulk, completion on an mbarrier) two tiles deep, results leave through bulk stores -
//                memory traffic, fp64 arithmetic and the output stream of three consecutive tiles overlap.  The
//                kernel is HBM bound: 3204 B per signal (c64 y, b in; c128 phi out).
#include "common.cuh"
#include "tc.cuh"

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

// The recursion for one signal spread over GL lanes (MAXE elements per lane, element j = lane + GL e).
//   D[e] = |b_j|^2, yb[e] = y_j conj(b_j) on entry (0 for j >= n); phi[e] on exit.
template <int MAXE, int GL>
__device__ __forceinline__ void classic_iterate(const double (&D)[MAXE], const double2 (&dyb)[MAXE], double rho,
                                                int n_iter, double2 (&phi)[MAXE]) {
    double dsum = 0.0;
    double Dr[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        dsum += D[e];
        Dr[e] = rho * D[e];
        phi[e] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int o = GL / 2; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    const double rden = rho / (1.0 + rho * dsum);
    for (int it = 0; it < n_iter; ++it) {
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            phi[e].x = fma(Dr[e], phi[e].x, dyb[e].x);         // D (y/b + rho phi)
            phi[e].y = fma(Dr[e], phi[e].y, dyb[e].y);
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
}

__device__ __forceinline__ void classic_prepare(double2 bj, double2 yj, double& D, double2& dyb) {
    D = bj.x * bj.x + bj.y * bj.y;                             // (b * conj(b)).real          admm.py:78
    dyb = make_double2(yj.x * bj.x + yj.y * bj.y, yj.y * bj.x - yj.x * bj.y);     // y conj(b) = D * (y / b)
    if (D == 0.0) dyb = make_double2(__longlong_as_double(0x7ff8000000000000LL), __longlong_as_double(0x7ff8000000000000LL));
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
    double2 dyb[MAXE], phi[MAXE];
#pragma unroll
    for (int e = 0; e < MAXE; ++e) {
        const int j = lane + GL * e;
        D[e] = 0.0;
        dyb[e] = make_double2(0.0, 0.0);
        if (valid && j < n) classic_prepare(ld_c<CIn>(b, base + j), ld_c<CIn>(y, base + j), D[e], dyb[e]);
    }
    classic_iterate<MAXE, GL>(D, dyb, rho, n_iter, phi);
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
// 8 / 16 lanes per signal (13 / 7 elements per lane, n <= 104 / 112): the per-iteration reduction is 3 / 4 shuffle
// steps shared by the 4 / 2 signals of a warp instead of 5 steps for one.
template __global__ void k_classic<double2, 13, 8>(const double2*, const double2*, int, int, double, int, double2*);
template __global__ void k_classic<float2, 13, 8>(const float2*, const float2*, int, int, double, int, double2*);
template __global__ void k_classic<double2, 7, 16>(const double2*, const double2*, int, int, double, int, double2*);
template __global__ void k_classic<float2, 7, 16>(const float2*, const float2*, int, int, double, int, double2*);

// ---- persistent streaming form (complex64 input, n even)
// GL lanes per signal (MAXE elements per lane), NT threads: a tile holds S = (NT / 32) * (32 / GL) signals.
template <int GL, int NT>
__host__ __device__ constexpr int classic_p_tile() { return (NT / 32) * (32 / GL); }
template <int GL, int NT>
__host__ __device__ inline size_t classic_p_smem_bytes(int n) {
    // 2 input stages x (y, b) c64 + 2 output buffers c128 + barriers
    return (size_t)2 * 2 * classic_p_tile<GL, NT>() * n * sizeof(float2) + (size_t)2 * classic_p_tile<GL, NT>() * n * sizeof(double2) + 64;
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }

template <int GL, int MAXE, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
k_classic_p(const float2* __restrict__ y, const float2* __restrict__ b, int B, int n, double rho, int n_iter,
            double2* __restrict__ phi_out) {
    constexpr int CLP_S = classic_p_tile<GL, NT>();
    constexpr int SPW = 32 / GL;                               // signals per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const size_t in_elems = (size_t)CLP_S * n;
    double2* outb = reinterpret_cast<double2*>(smem_raw);                                  // [2][S*n]
    float2* inb = reinterpret_cast<float2*>(smem_raw + 2 * in_elems * sizeof(double2));    // [2 stages][y | b][S*n]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + 2 * in_elems * sizeof(double2) + 4 * in_elems * sizeof(float2));
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int ntiles = (B + CLP_S - 1) / CLP_S;
    const int first = blockIdx.x, step = gridDim.x;
    auto issue = [&](int tile, int stage) {                   // thread 0 only
        const int s0 = tile * CLP_S, ns = min(CLP_S, B - s0);
        const uint32_t bytes = (uint32_t)((size_t)ns * n * sizeof(float2));
        tc::mbar_expect_tx(&bar[stage], 2 * bytes);
        bulk_load(inb + (size_t)(2 * stage) * in_elems, y + (size_t)s0 * n, bytes, &bar[stage]);
        bulk_load(inb + (size_t)(2 * stage + 1) * in_elems, b + (size_t)s0 * n, bytes, &bar[stage]);
    };
    if (tid == 0) {
        tc::mbar_init(&bar[0], 1);
        tc::mbar_init(&bar[1], 1);
        tc::mbar_fence_init();
        if (first < ntiles) issue(first, 0);
        if (first + step < ntiles) issue(first + step, 1);
    }
    __syncthreads();
    const int g = lane / GL, l8 = lane % GL;                  // signal of the warp's SPW, lane within the signal
    int i = 0;
    for (int tile = first; tile < ntiles; tile += step, ++i) {
        const int stage = i & 1;
        const int s0 = tile * CLP_S, ns = min(CLP_S, B - s0);
        tc::mbar_wait(&bar[stage], (uint32_t)((i >> 1) & 1));
        const float2* ys = inb + (size_t)(2 * stage) * in_elems;
        const float2* bs = ys + in_elems;
        double2* os = outb + (size_t)stage * in_elems;
        const int sl = SPW * wid + g;                         // signal within the tile
        ADMM_ASSERT(sl < CLP_S && ns >= 1 && ns <= CLP_S && s0 + ns <= B && n <= GL * MAXE);
        const bool valid = sl < ns;
        double D[MAXE];
        double2 dyb[MAXE], phi[MAXE];
#pragma unroll
        for (int e = 0; e < MAXE; ++e) {
            const int j = l8 + GL * e;
            D[e] = 0.0;
            dyb[e] = make_double2(0.0, 0.0);
            if (valid && j < n) {
                const float2 bj = bs[(size_t)sl * n + j], yj = ys[(size_t)sl * n + j];
                classic_prepare(make_double2((double)bj.x, (double)bj.y), make_double2((double)yj.x, (double)yj.y), D[e], dyb[e]);
            }
        }
        classic_iterate<MAXE, GL>(D, dyb, rho, n_iter, phi);
        if (valid) {
#pragma unroll
            for (int e = 0; e < MAXE; ++e) {
                const int j = l8 + GL * e;
                if (j < n) os[(size_t)sl * n + j] = phi[e];
            }
        }
        tc::fence_async_smem();                                // generic-proxy writes of `os` -> visible to the bulk store
        __syncthreads();                                       // the stage's inputs are consumed, its outputs complete
        if (tid == 0) {
            bulk_store(phi_out + (size_t)s0 * n, os, (uint32_t)((size_t)ns * n * sizeof(double2)));
            bulk_commit();
            const int nxt = tile + 2 * step;
            if (nxt < ntiles) issue(nxt, stage);
            bulk_wait_read<1>();                               // the OTHER output buffer (next tile's) has been read out
        }
        __syncthreads();
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all stores complete before the CTA exits
}

}  // namespace admmnet
