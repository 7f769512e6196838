// Register-resident Householder tridiagonalisation of a Hermitian matrix (first half of torch.linalg.eigh,
// admm_net.py:303), the form k_head runs by default.
//
// The lower triangle lives in the REGISTERS of a 16 x 16 thread grid for the whole reduction (2-D cyclic ownership, so
// the work stays balanced while the trailing block shrinks), instead of being read and re-written in shared memory at
// every step:  thread (tr, tc) owns the elements A[r][c] whose REVERSED indices u = d-1-index satisfy u_r = 16 sr + tr,
// u_c = 16 sc + tc, u_r <= u_c (sr <= sc < NS).  Reversed, because the live block of step k is then the leading block
// u < m = d-1-k: a thread's live elements are its first slots and every loop bound is a compile-time unroll with a
// warp-uniform early exit.  Same mathematics as tridiag_smem (LAPACK zhetd2 'L': H_k = I - tau_k v_k v_k^H, v_k on rows
// k+1.., unit entry first) with the rank-2 update of step k-1 delayed into the pass of step k:
//     pass k :  A <- A - v w^H - w v^H  (pending, lower triangle only),   p = A v_k  through the Hermitian symmetry
//               (each stored element feeds its row sum and, conjugated, its column sum: 16 FMA per element, no
//               shared-memory traffic for A);  the thread column that owns the NEXT column of A hands it over in colb
//     reduce :  row partials (8 per index, the two thread columns of a warp pre-added by one shuffle) + column
//               partials (16 per index) -> p, two threads per index
//     finish :  warp 0 alone, registers + shuffles only: w = tau p - (tau/2)(p^H v) v, pending (v, w) published; the
//               next column with the pending update applied -> d_k, e_k, tau_k, v_k (stored to global straight away)
// Three block barriers per step.  Dead indices (u >= m) carry zero vectors, so no element needs a liveness predicate.
#pragma once
#include "common.cuh"

namespace admmnet {

template <int NS>
struct TrdLayout {
    static constexpr int NU = 16 * NS;        // index slots
    static constexpr int PR = NU + 8;         // stride of one warp's row partials
    static constexpr int PCS = 17;            // column partials per index (16 + 1: conflict-free for writer and reader)
    static constexpr int F2 = 8 * PR + NU * PCS + 2 * NU + 32;   // float2 entries of scratch (aliases the staging copy of A)
};
constexpr int TRD_SCRATCH_F2 = TrdLayout<8>::F2;

// A: column-major staging copy (leading dimension ld, lower triangle read), dead after the first barrier: the scratch
// aliases it (the caller sizes the region as max(d*ld, TRD_SCRATCH_F2)).  vw/vn: [128] each.  Vg: reflector store of
// this signal in global memory (voff layout).  256 threads.
template <int NS>
__device__ __forceinline__ void tridiag_reg(float2* __restrict__ A, int d, int ld, float4* __restrict__ vw,
                                            float2* __restrict__ vn, float2* __restrict__ tau_out, float* __restrict__ dd,
                                            float* __restrict__ ee, float2* __restrict__ Vg) {
    using L = TrdLayout<NS>;
    constexpr int NU = L::NU, PR = L::PR, PCS = L::PCS;
    constexpr int NEL = NS * (NS + 1) / 2;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tr = tid & 15, tc = tid >> 4;
#define TRD_IDX(sc, sr) ((sc) * ((sc) + 1) / 2 + (sr))
    float2 a[NEL];
#pragma unroll
    for (int sc = 0; sc < NS; ++sc) {
#pragma unroll
        for (int sr = 0; sr <= sc; ++sr) {
            const int ur = 16 * sr + tr, uc = 16 * sc + tc;
            float2 x = make_float2(0.f, 0.f);
            if (uc < d && ur <= uc) {
                x = A[(d - 1 - ur) + (size_t)(d - 1 - uc) * ld];
                if (ur == uc) x.y = 0.f;
                if (uc == d - 1) vn[ur] = x;          // the first column to be eliminated (step 0 reads it from vn)
            }
            a[TRD_IDX(sc, sr)] = x;
        }
    }
    __syncthreads();                          // the staging copy is dead: scratch aliases it from here on
    float2* partr = A;                        // [8][PR]
    float2* partc = partr + 8 * PR;           // [NU][PCS]
    float2* pbuf = partc + NU * PCS;          // [NU]  p = tau A v of the step being finished
    float2* colb = pbuf + NU;                 // [NU]  next column, handed over by its owners during the pass
    float2* red = colb + NU;                  // [8]   per-warp partials of p^H v
    float* red2 = reinterpret_cast<float*>(red + 8);       // [8] per-warp partials of the column norm
    float2* alpha_s = red + 16;               // [1]
    // (vw and vn need no clearing: the finishing sections of every step rewrite all NU entries before they are read)
    const int u = tid >> 1;                   // the index this thread pair reduces and finishes (h == 0 lane keeps it)
    const bool h0 = (tid & 1) == 0;
    if (tid < 8) red[tid] = make_float2(0.f, 0.f);
    if (tid < NU) pbuf[tid] = make_float2(0.f, 0.f);
    float2 tau_prev = make_float2(0.f, 0.f);  // replicated in every thread
    float2 pu = make_float2(0.f, 0.f), vu = make_float2(0.f, 0.f);     // p_u and v_u of the step being finished
    __syncthreads();
    for (int k = 0; k < d; ++k) {
        const int m = d - 1 - k;              // live indices of step k: u < m; the column being eliminated is u = m
        // ---- finish 1 (all threads): p^H v, w = p - (tau/2)(p^H v) v, pending (v, w) published; column k with the
        //      pending update applied (v at u = m is the unit entry of the previous reflector)
        float2 ac = make_float2(0.f, 0.f);
        {
            float dx = 0.f, dy = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float2 t = red[j]; dx += t.x; dy += t.y; }
            const float2 a2 = cmul(make_float2(-0.5f * tau_prev.x, -0.5f * tau_prev.y), make_float2(dx, dy));
            const float2 wu = cadd(pu, cmul(a2, vu));
            const float2 wm = cadd(pbuf[m], a2);
            float ssp = 0.f;
            if (h0 && u < NU) {
                vw[u] = (u < m) ? make_float4(vu.x, vu.y, wu.x, wu.y) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (u <= m) {
                    float2 x = (k == 0 ? vn : colb)[u];
                    // x -= v_u conj(w_m) + w_u
                    x.x = fmaf(-vu.x, wm.x, x.x); x.x = fmaf(-vu.y, wm.y, x.x); x.x -= wu.x;
                    x.y = fmaf(-vu.y, wm.x, x.y); x.y = fmaf(vu.x, wm.y, x.y);  x.y -= wu.y;
                    ac = x;
                    if (u + 2 <= m) ssp = fmaf(x.x, x.x, x.y * x.y);
                    if (u == m) dd[k] = x.x;
                    if (u + 1 == m) *alpha_s = x;
                }
            }
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) ssp += __shfl_xor_sync(0xffffffffu, ssp, o);
            if (lane == 0) red2[wid] = ssp;
        }
        if (m == 0) break;
        __syncthreads();
        // ---- finish 2 (all threads, the scalars redundantly): reflector k
        {
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) ss += red2[j];
            const float2 alpha = *alpha_s;
            float2 tau, scale;
            float beta;
            if (ss == 0.f && alpha.y == 0.f) {
                tau = make_float2(0.f, 0.f);
                scale = make_float2(0.f, 0.f);
                beta = alpha.x;
            } else {
                // (plain divisions, not reciprocals: a numerically rank-deficient trailing block leaves columns of
                //  denormal size, whose reciprocal overflows while the quotients stay finite)
                beta = -copysignf(sqrtf(alpha.x * alpha.x + alpha.y * alpha.y + ss), alpha.x);
                tau = make_float2((beta - alpha.x) / beta, -alpha.y / beta);
                scale = cdiv(make_float2(1.f, 0.f), make_float2(alpha.x - beta, alpha.y));
            }
            float2 vi = make_float2(0.f, 0.f);
            if (u + 1 == m) vi = make_float2(1.f, 0.f);
            else if (u + 1 < m) vi = cmul(ac, scale);
            vu = vi;
            if (h0 && u < NU) {
                vn[u] = vi;
                if (u + 1 <= m) Vg[voff(k, d) + (m - 1 - u)] = vi;
            }
            if (tid == 0) {
                tau_out[k] = tau;
                ee[k] = beta;
            }
            tau_prev = tau;
        }
        __syncthreads();
        // ---- pass k: pending rank-2 update of the live lower triangle + Hermitian mat-vec with v_k
        {
            int ncs = (m - (tc & ~1) + 15) >> 4;               // live column slots of this warp (uniform)
            ncs = ncs < 0 ? 0 : (ncs > NS ? NS : ncs);
            const int un = m - 1;                              // the next column to be eliminated
            float2 accr[NS];
#pragma unroll
            for (int sr = 0; sr < NS; ++sr) accr[sr] = make_float2(0.f, 0.f);
#pragma unroll
            for (int sc = 0; sc < NS; ++sc) {
                if (sc < ncs) {
                    const int uc = 16 * sc + tc;
                    const float4 cvw = vw[uc];
                    const float2 cvn = vn[uc];
                    float2 accc = make_float2(0.f, 0.f);
#pragma unroll
                    for (int sr = 0; sr <= sc; ++sr) {
                        if (sr < sc || tr <= tc) {
                            const int ur = 16 * sr + tr;
                            const float4 rvw = vw[ur];
                            const float2 rvn = vn[ur];
                            float2 x = a[TRD_IDX(sc, sr)];
                            // x -= v_r conj(w_c) + w_r conj(v_c)
                            x.x = fmaf(-rvw.x, cvw.z, x.x); x.x = fmaf(-rvw.y, cvw.w, x.x);
                            x.x = fmaf(-rvw.z, cvw.x, x.x); x.x = fmaf(-rvw.w, cvw.y, x.x);
                            x.y = fmaf(-rvw.y, cvw.z, x.y); x.y = fmaf(rvw.x, cvw.w, x.y);
                            x.y = fmaf(-rvw.w, cvw.x, x.y); x.y = fmaf(rvw.z, cvw.y, x.y);
                            const bool diag = (sr == sc) && (tr == tc);
                            if (diag) x.y = 0.f;
                            a[TRD_IDX(sc, sr)] = x;
                            // p_r += A[r][c] v_c
                            accr[sr].x = fmaf(x.x, cvn.x, accr[sr].x); accr[sr].x = fmaf(-x.y, cvn.y, accr[sr].x);
                            accr[sr].y = fmaf(x.x, cvn.y, accr[sr].y); accr[sr].y = fmaf(x.y, cvn.x, accr[sr].y);
                            // p_c += conj(A[r][c]) v_r   (off the diagonal)
                            if (!diag) {
                                accc.x = fmaf(x.x, rvn.x, accc.x); accc.x = fmaf(x.y, rvn.y, accc.x);
                                accc.y = fmaf(x.x, rvn.y, accc.y); accc.y = fmaf(-x.y, rvn.x, accc.y);
                            }
                        }
                    }
                    partc[uc * PCS + tr] = accc;
                    if (uc == un) {
#pragma unroll
                        for (int sr = 0; sr <= sc; ++sr)
                            if (sr < sc || tr <= tc) colb[16 * sr + tr] = a[TRD_IDX(sc, sr)];
                    }
                }
            }
#pragma unroll
            for (int sr = 0; sr < NS; ++sr) {
                accr[sr].x += __shfl_xor_sync(0xffffffffu, accr[sr].x, 16);
                accr[sr].y += __shfl_xor_sync(0xffffffffu, accr[sr].y, 16);
                if (lane < 16) partr[wid * PR + 16 * sr + tr] = accr[sr];
            }
        }
        __syncthreads();
        // ---- reduce the partials (two threads per index, 12 entries each) -> p = tau A v, per-warp partials of p^H v
        {
            const int h = tid & 1;
            float2 s = make_float2(0.f, 0.f);
            if (u < m) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float2 t = h ? partc[u * PCS + 4 + j] : partr[j * PR + u];
                    s.x += t.x; s.y += t.y;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 t = partc[u * PCS + (h ? 12 + j : j)];
                    s.x += t.x; s.y += t.y;
                }
            }
            s.x += __shfl_xor_sync(0xffffffffu, s.x, 1);
            s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
            float dx = 0.f, dy = 0.f;
            pu = make_float2(0.f, 0.f);
            if (h0 && u < m) {
                pu = cmul(tau_prev, s);
                pbuf[u] = pu;
                dx = fmaf(pu.x, vu.x, pu.y * vu.y);          // conj(p) * v
                dy = fmaf(pu.x, vu.y, -pu.y * vu.x);
            }
#pragma unroll
            for (int o = 16; o > 1; o >>= 1) {
                dx += __shfl_xor_sync(0xffffffffu, dx, o);
                dy += __shfl_xor_sync(0xffffffffu, dy, o);
            }
            if (lane == 0) red[wid] = make_float2(dx, dy);
        }
        __syncthreads();
    }
#undef TRD_IDX
    __syncthreads();
}

}  // namespace admmnet
