// ADMM-Net layer kernels for sm_100a.
//
// One unrolled layer (reference: admm_net.py:757-762) is five launches over a chunk of signals:
//   k_head   : Z += alpha*(G-C) of the previous layer (ZLayer, admm_net.py:388-412), phi-update
//              (PhiLayer, 79-105), H-update (HLayer, 134-194), build the Hermitian block matrix
//              (GLayer._build_block_matrix, 262-290) in shared memory and reduce it to real
//              tridiagonal form by Householder reflectors (first half of torch.linalg.eigh, 303).
//   k_ql     : implicit-shift QL on the tridiagonal, one thread per signal, fp64 scalars; emits
//              eigenvalues and the plane-rotation stream.
//   k_rot    : applies the rotation stream to I (rows are independent) -> real eigenvectors of T.
//   k_tail   : back-transforms with the reflectors (U = Q_H * Z), maps eigenvalues
//              (GLayer._eigenvalues_projection, 310-334), rebuilds G = U diag(l') U^H (336-354)
//              and the residual norm r = ||G - C||_F (ZLayer._compute_adaptive_step, 454).
//   k_mean   : batch mean of r (admm_net.py:459), deterministic.
// The last layer only needs the phi-update (k_final_phi): its H/G/Z updates never reach the output.
#include "common.cuh"
#include "trd_reg.cuh"

namespace admmnet {

// =====================================================================================
// Householder tridiagonalisation of a Hermitian matrix held in shared memory.
//   A : column-major, leading dimension ld (odd), both triangles valid on entry.
// On exit the strictly-lower part of column k (rows k+2..d-1) holds v_k (v_k[k+1] = 1 implicit),
// tau[k] the reflector scalars, dd/ee the real tridiagonal.  A = Q T Q^H, Q = H_0 H_1 ... H_{d-2},
// H_k = I - tau_k v_k v_k^H   (same convention as LAPACK zhetd2 'L').
// 256 threads: (r = tid & 127, half = tid >> 7).
// =====================================================================================
struct TriScratch {
    float4* vw;     // [128] pending rank-2 update, indexed by global row: (v.x, v.y, w.x, w.y)
    float2* vn;     // [128] reflector being formed, indexed by global row
    float2* part;   // [8][128] partial mat-vec sums
    float* red;     // [2][32] block-reduction scratch (double buffered)
    float2* scal;   // [4]
};

// One fused sweep over the trailing block (rows/cols k+1..d-1):
//     A <- A - v w^H - w v^H   (pending update of the previous step)     and     part[q][r] = sum_c A[r][c] vn[c]
// NROWS rows per thread (r, r+RSTRIDE); column groups q = tid / RSTRIDE, c = q, q+NQ, ...
template <int NROWS, int RSTRIDE, int NT, int PSTR>
__device__ __forceinline__ void tri_fused_pass(float2* __restrict__ A, int ld, int k, int m, const TriScratch& S) {
    constexpr int NQ = NT / RSTRIDE;
    const int tid = threadIdx.x;
    const int rr = tid & (RSTRIDE - 1), q = tid / RSTRIDE;
    const int g0 = k + 1;                       // global index of local row/col 0
    float2 vo[NROWS], wo[NROWS], acc[NROWS];
    bool on[NROWS];
#pragma unroll
    for (int i = 0; i < NROWS; ++i) {
        const int rl = rr + i * RSTRIDE;
        on[i] = rl < m;
        const float4 t = S.vw[g0 + (on[i] ? rl : 0)];
        vo[i] = make_float2(t.x, t.y);
        wo[i] = make_float2(t.z, t.w);
        acc[i] = make_float2(0.f, 0.f);
    }
    if (on[0]) {
        float2* a0 = A + (g0 + rr) + (size_t)g0 * ld;
        // CU columns per trip: all shared-memory loads first, then the FMA chains, then the stores
        constexpr int CU = NROWS >= 3 ? 2 : 4;
        int c = q;
        for (; c + (CU - 1) * NQ < m; c += CU * NQ) {
            float4 t[CU];
            float2 vn[CU], x[CU][NROWS];
#pragma unroll
            for (int u = 0; u < CU; ++u) {
                t[u] = S.vw[g0 + c + u * NQ];
                vn[u] = S.vn[g0 + c + u * NQ];
#pragma unroll
                for (int i = 0; i < NROWS; ++i)
                    if (i == 0 || on[i]) x[u][i] = a0[i * RSTRIDE + (size_t)(c + u * NQ) * ld];
            }
#pragma unroll
            for (int u = 0; u < CU; ++u) {
#pragma unroll
                for (int i = 0; i < NROWS; ++i) {
                    if (i == 0 || on[i]) {
                        float2 xx = x[u][i];
                        xx.x = fmaf(-vo[i].x, t[u].z, xx.x); xx.x = fmaf(-vo[i].y, t[u].w, xx.x);
                        xx.x = fmaf(-wo[i].x, t[u].x, xx.x); xx.x = fmaf(-wo[i].y, t[u].y, xx.x);
                        xx.y = fmaf(-vo[i].y, t[u].z, xx.y); xx.y = fmaf(vo[i].x, t[u].w, xx.y);
                        xx.y = fmaf(-wo[i].y, t[u].x, xx.y); xx.y = fmaf(wo[i].x, t[u].y, xx.y);
                        a0[i * RSTRIDE + (size_t)(c + u * NQ) * ld] = xx;
                        acc[i].x = fmaf(xx.x, vn[u].x, acc[i].x); acc[i].x = fmaf(-xx.y, vn[u].y, acc[i].x);
                        acc[i].y = fmaf(xx.x, vn[u].y, acc[i].y); acc[i].y = fmaf(xx.y, vn[u].x, acc[i].y);
                    }
                }
            }
        }
        for (; c < m; c += NQ) {
            const float4 t = S.vw[g0 + c];
            const float2 vn = S.vn[g0 + c];
#pragma unroll
            for (int i = 0; i < NROWS; ++i) {
                if (i == 0 || on[i]) {
                    float2* ap = a0 + i * RSTRIDE + (size_t)c * ld;
                    float2 x = *ap;
                    // x -= v_r conj(w_c) + w_r conj(v_c)
                    x.x = fmaf(-vo[i].x, t.z, x.x); x.x = fmaf(-vo[i].y, t.w, x.x);
                    x.x = fmaf(-wo[i].x, t.x, x.x); x.x = fmaf(-wo[i].y, t.y, x.x);
                    x.y = fmaf(-vo[i].y, t.z, x.y); x.y = fmaf(vo[i].x, t.w, x.y);
                    x.y = fmaf(-wo[i].y, t.x, x.y); x.y = fmaf(wo[i].x, t.y, x.y);
                    *ap = x;
                    acc[i].x = fmaf(x.x, vn.x, acc[i].x); acc[i].x = fmaf(-x.y, vn.y, acc[i].x);
                    acc[i].y = fmaf(x.x, vn.y, acc[i].y); acc[i].y = fmaf(x.y, vn.x, acc[i].y);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NROWS; ++i) {
        const int rl = rr + i * RSTRIDE;
        if (rl < PSTR) S.part[q * PSTR + rl] = acc[i];
    }
}

// Per Householder step k (trailing block = rows/cols k+1..d-1, size m):
//   warp 0 alone :  finish step k-1 (p = tau*sum(partials), w = p - (tau/2)(p^H v) v -> pending (v,w)),
//                   then column k with the pending update applied -> d[k], reflector k (tau, v_new).
//                   The rows it needs for the second half are the ones it just produced, so everything
//                   stays in registers / warp shuffles: no block barrier inside this section.
//   barrier
//   all 8 warps  :  fused pass  A <- A - v w^H - w v^H  and  partial products  A v_new
//   barrier
// i.e. two block barriers per step.
// Runs steps k = 0 .. k_stop-1.  k_stop == d-1: complete reduction (dd[d-1] included).  k_stop < d-1: stops
// with the rank-2 update of step k_stop-1 still pending in S.vw (rows k_stop..d-1); the caller applies it
// while handing the trailing block to the next stage (tri_store_trailing).
template <int NT, int PSTR>
__device__ void tridiag_smem(float2* __restrict__ A, int d, int ld, TriScratch S, float2* tau_out, float* dd,
                             float* ee, int k_stop) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 128) S.vw[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    float2 tau_prev = make_float2(0.f, 0.f);
    const bool partial = k_stop < d - 1;
    for (int k = 0; k <= k_stop; ++k) {
        const int m = d - k - 1;            // trailing size of step k (m == 0: only the epilogue of step d-2)
        const bool last = (k == k_stop);
        if (wid == 0) {
            // rows k..d-1 = local rows t = lane + 32 j of step k-1's trailing block
            float2 vr[4], w[4];
            const int mp = m + 1;
            // ---- finish step k-1
            if (k > 0) {
                constexpr int nq = NT / 32;                     // column groups of the pass that produced `part`
                float2 p[4];
                float dx = 0.f, dy = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = lane + 32 * j;
                    p[j] = make_float2(0.f, 0.f);
                    vr[j] = make_float2(0.f, 0.f);
                    if (t < mp) {
                        float2 sacc = make_float2(0.f, 0.f);
#pragma unroll
                        for (int qq = 0; qq < nq; ++qq) sacc = cadd(sacc, S.part[qq * PSTR + t]);
                        p[j] = cmul(tau_prev, sacc);
                        vr[j] = S.vn[k + t];
                        dx = fmaf(p[j].x, vr[j].x, dx); dx = fmaf(p[j].y, vr[j].y, dx);      // conj(p) * v
                        dy = fmaf(p[j].x, vr[j].y, dy); dy = fmaf(-p[j].y, vr[j].x, dy);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    dx += __shfl_xor_sync(0xffffffffu, dx, o);
                    dy += __shfl_xor_sync(0xffffffffu, dy, o);
                }
                const float2 a2 = cmul(make_float2(-0.5f * tau_prev.x, -0.5f * tau_prev.y), make_float2(dx, dy));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = lane + 32 * j;
                    w[j] = cadd(p[j], cmul(a2, vr[j]));
                    if (t < mp) S.vw[k + t] = make_float4(vr[j].x, vr[j].y, w[j].x, w[j].y);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) { vr[j] = make_float2(0.f, 0.f); w[j] = make_float2(0.f, 0.f); }
            }
            if (!(last && partial)) {
            // pending (v,w) of row k (local row 0: lane 0, slot 0)
            const float vkx = __shfl_sync(0xffffffffu, vr[0].x, 0), vky = __shfl_sync(0xffffffffu, vr[0].y, 0);
            const float wkx = __shfl_sync(0xffffffffu, w[0].x, 0), wky = __shfl_sync(0xffffffffu, w[0].y, 0);
            // ---- column k with the pending update applied
            float2 a[4];
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = lane + 32 * j;
                a[j] = make_float2(0.f, 0.f);
                if (t < mp) {
                    float2 x = A[(k + t) + (size_t)k * ld];
                    // x -= v_r conj(w_k) + w_r conj(v_k)
                    x.x = fmaf(-vr[j].x, wkx, x.x); x.x = fmaf(-vr[j].y, wky, x.x);
                    x.x = fmaf(-w[j].x, vkx, x.x);  x.x = fmaf(-w[j].y, vky, x.x);
                    x.y = fmaf(-vr[j].y, wkx, x.y); x.y = fmaf(vr[j].x, wky, x.y);
                    x.y = fmaf(-w[j].y, vkx, x.y);  x.y = fmaf(w[j].x, vky, x.y);
                    a[j] = x;
                    if (t >= 2) ss = fmaf(x.x, x.x, fmaf(x.y, x.y, ss));
                }
            }
            if (lane == 0) dd[k] = a[0].x;
            if (m > 0) {
                ss = warp_sum(ss);
                const float2 alpha = make_float2(__shfl_sync(0xffffffffu, a[0].x, 1), __shfl_sync(0xffffffffu, a[0].y, 1));
                float2 tau, scale;
                float beta;
                if (ss == 0.f && alpha.y == 0.f) {
                    tau = make_float2(0.f, 0.f);
                    scale = make_float2(0.f, 0.f);
                    beta = alpha.x;
                } else {
                    beta = -copysignf(sqrtf(alpha.x * alpha.x + alpha.y * alpha.y + ss), alpha.x);
                    tau = make_float2((beta - alpha.x) / beta, -alpha.y / beta);
                    scale = cdiv(make_float2(1.f, 0.f), make_float2(alpha.x - beta, alpha.y));
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = lane + 32 * j;
                    if (t >= 1 && t < mp) {
                        float2 vi = make_float2(1.f, 0.f);
                        if (t >= 2) {
                            vi = cmul(a[j], scale);
                            A[(k + t) + (size_t)k * ld] = vi;          // kept for the export of the reflectors
                        }
                        S.vn[k + t] = vi;
                    }
                }
                if (lane == 0) {
                    tau_out[k] = tau;
                    ee[k] = beta;
                }
                tau_prev = tau;
            }
            }
        }
        if (last) break;
        __syncthreads();
        // ---- fused pending update + mat-vec with the new reflector
        // rows -> lanes in stripes of 32 (at most 31 idle lanes), NT/32 column groups
        if (PSTR > 96 && m > 96) tri_fused_pass<4, 32, NT, PSTR>(A, ld, k, m, S);
        else if (PSTR > 64 && m > 64) tri_fused_pass<3, 32, NT, PSTR>(A, ld, k, m, S);
        else if (m > 32) tri_fused_pass<2, 32, NT, PSTR>(A, ld, k, m, S);
        else tri_fused_pass<1, 32, NT, PSTR>(A, ld, k, m, S);
        __syncthreads();
    }
    __syncthreads();
}

// Hand the trailing block (rows/cols k0..d-1, pending update applied) to the next stage: full, row-major
// [d2][d2] in global memory (Hermitian, so row-/column-major only differ by conjugation; we store A[r][c]).
__device__ void tri_store_trailing(const float2* __restrict__ A, int d, int ld, int k0, const TriScratch& S,
                                   float2* __restrict__ T) {
    const int d2 = d - k0;
    for (int idx = threadIdx.x; idx < d2 * d2; idx += blockDim.x) {
        const int c = idx / d2, r = idx % d2;            // consecutive threads -> consecutive rows (smem stride 1)
        float2 x = A[(k0 + r) + (size_t)(k0 + c) * ld];
        const float4 tr = S.vw[k0 + r], tc = S.vw[k0 + c];
        x.x = fmaf(-tr.x, tc.z, x.x); x.x = fmaf(-tr.y, tc.w, x.x);
        x.x = fmaf(-tr.z, tc.x, x.x); x.x = fmaf(-tr.w, tc.y, x.x);
        x.y = fmaf(-tr.y, tc.z, x.y); x.y = fmaf(tr.x, tc.w, x.y);
        x.y = fmaf(-tr.w, tc.x, x.y); x.y = fmaf(tr.z, tc.y, x.y);
        T[(size_t)c * d2 + r] = x;                         // column-major [c][r]
    }
}

// Export one stage: reflectors k = 0..nk-1 of the local matrix (order d, global offset k0 inside a matrix of
// order dg), tau[k], and d/e entries i = 0..ni-1 ([i][B] layout for k_ql).
__device__ void export_tridiag(const float2* __restrict__ A, int d, int ld, const float2* tau_s, const float* dd,
                               const float* ee, float2* __restrict__ Vg, float2* __restrict__ taug,
                               float* __restrict__ dT, float* __restrict__ eT, int B, int sig, int k0, int dg, int nk,
                               int ni) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    for (int k = wid; k < nk; k += nw) {
        const int off = voff(k0 + k, dg);
        for (int i = k + 1 + lane; i < d; i += 32)
            Vg[off + i - (k + 1)] = (i == k + 1) ? make_float2(1.f, 0.f) : A[i + (size_t)k * ld];
    }
    for (int i = tid; i < ni; i += blockDim.x) {
        const int gi = k0 + i;
        taug[gi] = gi < dg - 1 ? tau_s[i] : make_float2(0.f, 0.f);
        dT[(size_t)gi * B + sig] = dd[i];
        eT[(size_t)gi * B + sig] = gi < dg - 1 ? ee[i] : 0.f;
    }
}

// d, e, tau of a complete reduction (the register-resident form stores its reflectors itself)
__device__ void export_de(const float2* tau_s, const float* dd, const float* ee, float2* __restrict__ taug,
                          float* __restrict__ dT, float* __restrict__ eT, int B, int sig, int d) {
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        taug[i] = i < d - 1 ? tau_s[i] : make_float2(0.f, 0.f);
        dT[(size_t)i * B + sig] = dd[i];
        eT[(size_t)i * B + sig] = i < d - 1 ? ee[i] : 0.f;
    }
}

// shared-memory carve-up shared by k_head and k_tridiag
struct HeadSmem {
    float2* A;
    TriScratch S;
    float2* tau;
    float* dd;
    float* ee;
    float2* phi;    // [n]
    float2* gcol;   // [n]
    float2* zeta;   // [n]
    float2* phip;   // [n] previous phi
    float* hp;      // [n] previous h
    float* t;       // [n]
    float* hid;     // [64]
    float* h;       // [n]
};
// the staging copy of A doubles as the scratch of the register-resident tridiagonalisation (trd_reg.cuh)
__host__ __device__ inline size_t head_a_f2(int d, int ld) {
    const size_t a = (size_t)d * ld;
    return a > (size_t)TRD_SCRATCH_F2 ? a : (size_t)TRD_SCRATCH_F2;
}
__host__ __device__ inline size_t head_smem_bytes(int d, int ld) {
    size_t f2 = head_a_f2(d, ld) + 1 /*align*/ + 256 /*vw*/ + 128 /*vn*/ + 1024 /*part*/ + 4 + 128 /*tau*/ +
                4 * 128 /*phi,gcol,zeta,phip*/;
    size_t f1 = 96 + 128 + 128 /*dd,ee*/ + 128 /*hp*/ + 128 /*t*/ + 64 + 128 /*h*/;
    return f2 * sizeof(float2) + f1 * sizeof(float);
}
__device__ inline HeadSmem carve_head(unsigned char* base, int d, int ld) {
    HeadSmem s;
    float2* p2 = reinterpret_cast<float2*>(base);
    s.A = p2; p2 += head_a_f2(d, ld);
    p2 += head_a_f2(d, ld) & 1;                          // float4 alignment of vw
    s.S.vw = reinterpret_cast<float4*>(p2); p2 += 256;
    s.S.vn = p2; p2 += 128;
    s.S.part = p2; p2 += 1024;
    s.S.scal = p2; p2 += 4;
    s.tau = p2; p2 += 128;
    s.phi = p2; p2 += 128;
    s.gcol = p2; p2 += 128;
    s.zeta = p2; p2 += 128;
    s.phip = p2; p2 += 128;
    float* p1 = reinterpret_cast<float*>(p2);
    s.S.red = p1; p1 += 96;
    s.dd = p1; p1 += 128;
    s.ee = p1; p1 += 128;
    s.hp = p1; p1 += 128;
    s.t = p1; p1 += 128;
    s.hid = p1; p1 += 64;
    s.h = p1; p1 += 128;
    return s;
}

struct HeadArgs {
    const float2* y;
    const float2* b;
    const float* sigma;
    float2* Zp;          // [B][npk]
    float2* GV;          // [B][npk]: G (packed lower) of the previous layer on entry, reflectors on exit
    float2* phi_cur;     // [B][n]
    float* h_cur;        // [B][n]
    const float* r_prev; // [B]
    const float* mean_prev;
    const float* Pk;
    const float* Pkm1;
    float2* tau;         // [B][d]
    float* dT;           // [d][B]
    float* eT;           // [d][B]
    float2* Ttr;         // [B][d2*d2] trailing block handed to k_head2 (k1 < d-1)
    int B, n, d, ld, first;
    int k1;              // Householder steps done here; d-1 = everything
    const int* skip;     // optional [B]: 1 = signal already handled by k_arrow (layer 0)
};

// NS = 0: shared-memory tridiagonalisation, possibly continued by k_head2 stages (a.k1 < d-1);
// NS = 7 / 8: register-resident form (d <= 112 / 128), always complete.
template <int NS>
__global__ void __launch_bounds__(256, NS == 8 ? 1 : 2) k_head(HeadArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, d = a.d, ld = a.ld;
    HeadSmem s = carve_head(smem_raw, d, ld);
    const int tid = threadIdx.x;
    const int sig = blockIdx.x;
    if (a.skip && a.skip[sig]) return;
    const int npk = d * (d + 1) / 2;
    const float* __restrict__ P = a.Pk;
    float2* Zp = a.Zp + (size_t)sig * npk;
    float2* GV = a.GV + (size_t)sig * npk;
    const float inv_rho_g = P[P_INV_RHO_G];
    const float rho_h_eps = P[P_RHO_H_EPS];

    // ---- stage previous phi/h, clear captures
    for (int j = tid; j < n; j += 256) {
        s.gcol[j] = make_float2(0.f, 0.f);
        s.zeta[j] = make_float2(0.f, 0.f);
        s.t[j] = 0.f;
        if (!a.first) {
            s.phip[j] = a.phi_cur[(size_t)sig * n + j];
            s.hp[j] = a.h_cur[(size_t)sig * n + j];
        }
    }
    __syncthreads();
    // ---- dual update of the previous layer + A = -Z/(rho_g+eps)
    if (a.first) {
        for (int idx = tid; idx < d * ld; idx += 256) s.A[idx] = make_float2(0.f, 0.f);
        for (int idx = tid; idx < npk; idx += 256) Zp[idx] = make_float2(0.f, 0.f);   // Z_0 = 0 (admm_net.py:754)
    } else {
        const float alpha = z_alpha(a.Pkm1, a.r_prev[sig], *a.mean_prev);
        const float c1z = a.Pkm1[P_C1Z];
        // flat, coalesced sweep over the packed triangles; 4 independent loads in flight per thread
        for (int base = 0; base < npk; base += 4 * 256) {
            float2 zv[4], gv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 256 + tid;
                if (idx < npk) { zv[u] = Zp[idx]; gv[u] = GV[idx]; }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int idx = base + u * 256 + tid;
                if (idx >= npk) continue;
                int i = (int)((sqrtf(8.f * (float)idx + 1.f) - 1.f) * 0.5f);      // row of packed index
                while ((i + 1) * (i + 2) / 2 <= idx) ++i;
                while (i * (i + 1) / 2 > idx) --i;
                const int j = idx - i * (i + 1) / 2;
                float2 z = zv[u];
                const float2 g = gv[u];
                float2 c = make_float2(0.f, 0.f);
                if (i == j) c.x = (i < n) ? s.hp[i] : c1z;
                else if (i == n) c = cconj(s.phip[j]);
                z.x += alpha * (g.x - c.x);
                z.y += alpha * (g.y - c.y);
                if (i == j) z.y = 0.f;
                Zp[idx] = z;
                if (i == j && i < n) s.t[i] = g.x + z.x / rho_h_eps;
                if (i == n && j < n) {
                    s.gcol[j] = cconj(g);
                    s.zeta[j] = cconj(z);
                }
                const float2 av = make_float2(-inv_rho_g * z.x, -inv_rho_g * z.y);
                s.A[i + (size_t)j * ld] = av;
                if (i != j) s.A[j + (size_t)i * ld] = cconj(av);
            }
        }
    }
    __syncthreads();
    // ---- phi update (admm_net.py:94-103)
    const float rho_phi = P[P_RHO_PHI];
    for (int j = tid; j < n; j += 256) {
        const float2 bj = a.b[(size_t)sig * n + j], yj = a.y[(size_t)sig * n + j];
        const float ab = hypotf(bj.x, bj.y);
        const float bsq = ab * ab + ADMM_EPS;
        const float wgt = bsq / (1.f + rho_phi * bsq);
        const float2 yob = cdiv(yj, make_float2(bj.x + ADMM_EPS, bj.y));
        const float2 g = s.gcol[j], z = s.zeta[j];
        const float2 ph = make_float2(wgt * (yob.x + rho_phi * g.x + z.x), wgt * (yob.y + rho_phi * g.y + z.y));
        s.phi[j] = ph;
        a.phi_cur[(size_t)sig * n + j] = ph;
    }
    // ---- H update (admm_net.py:146-192): correction MLP n -> 64 -> n
    if (tid < 64) {
        float acc = P[P_HB1 + tid];
        const float* __restrict__ W1T = P + P_HW1T;
        for (int i = 0; i < n; ++i) acc += W1T[i * 64 + tid] * s.t[i];
        s.hid[tid] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    float tc = 0.f;
    if (tid < n) {
        const float* __restrict__ W2T = P + P_HW1T + 64 * n;
        float acc = P[P_HW1T + 128 * n + tid];
#pragma unroll 8
        for (int j = 0; j < 64; ++j) acc += W2T[j * n + tid] * s.hid[j];
        tc = s.t[tid] + 0.1f * tanhf(acc);
    }
    {
        const float linf = block_max(tid < n ? fabsf(tc) : 0.f, s.S.red);
        float sm[1] = {tid < n ? tc : 0.f};
        block_sum<1>(sm, s.S.red);
        const float sg = a.sigma[sig];
        const float Asig = 2.f * sqrtf((float)n) * sg + sg * sg;
        const float cv = Asig * linf + sm[0];
        const float scale = fminf(P[P_SIG_PW] / (cv + ADMM_EPS), 1.f);
        if (tid < n) {
            const float hv = tc * scale;
            s.h[tid] = hv;
            a.h_cur[(size_t)sig * n + tid] = hv;
        }
    }
    __syncthreads();
    // ---- A += [[diag(h), phi],[phi^H, c0]]
    for (int j = tid; j < n; j += 256) {
        s.A[j + (size_t)j * ld].x += s.h[j];
        const float2 ph = s.phi[j];
        float2* up = &s.A[j + (size_t)n * ld];   // A[j][n]
        up->x += ph.x; up->y += ph.y;
        float2* lo = &s.A[n + (size_t)j * ld];   // A[n][j]
        lo->x += ph.x; lo->y -= ph.y;
    }
    if (tid == 0) s.A[n + (size_t)n * ld].x += P[P_C0];
    __syncthreads();
    // ---- eigh, stage 1
    if (NS > 0) {
        tridiag_reg<(NS > 0 ? NS : 7)>(s.A, d, ld, s.S.vw, s.S.vn, s.tau, s.dd, s.ee, GV);
        export_de(s.tau, s.dd, s.ee, a.tau + (size_t)sig * d, a.dT, a.eT, a.B, sig, d);
        return;
    }
    tridiag_smem<256, 128>(s.A, d, ld, s.S, s.tau, s.dd, s.ee, a.k1);
    const bool full = a.k1 >= d - 1;
    export_tridiag(s.A, d, ld, s.tau, s.dd, s.ee, GV, a.tau + (size_t)sig * d, a.dT, a.eT, a.B, sig, 0, d,
                   full ? d - 1 : a.k1, full ? d : a.k1);
    if (!full) {
        const int d2 = d - a.k1;
        tri_store_trailing(s.A, d, ld, a.k1, s.S, a.Ttr + (size_t)sig * d2 * d2);
    }
}

// Later stages of the tridiagonalisation: the trailing block (order d2 <= 80, compacted by the previous
// stage) needs far less shared memory than the full matrix (62 / 42 / 24 KB at d2 = 80 / 60 / 40 against
// 101 KB), so 3 / 5 / 8 CTAs share an SM and hide each other's serial sections and barriers.
struct Head2Args {
    const float2* Tin;   // [B][d2*d2] column-major trailing block from the previous stage
    float2* Tout;        // [B][d3*d3] trailing block for the next stage (k_stop < d2-1)
    float2* GV;          // [B][npk] reflector store (offsets of the full matrix)
    float2* tau;         // [B][d]
    float* dT;
    float* eT;
    int B, d, d2, ld2, k0, k_stop;
    const int* skip;
};
template <int NT, int PSTR>
__host__ __device__ inline size_t head2_smem_bytes(int d2, int ld2) {
    return ((size_t)d2 * ld2 + 1 + 256 + 128 + (NT / 32) * PSTR + 4 + 128) * sizeof(float2) + (128 + 128) * sizeof(float);
}
template <int NT, int PSTR>
__global__ void __launch_bounds__(NT, NT == 128 ? 8 : 3) k_head2(Head2Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d2 = a.d2, ld2 = a.ld2;
    float2* p2 = reinterpret_cast<float2*>(smem_raw);
    float2* A = p2; p2 += (size_t)d2 * ld2;
    p2 += ((size_t)d2 * ld2) & 1;
    TriScratch S;
    S.vw = reinterpret_cast<float4*>(p2); p2 += 256;
    S.vn = p2; p2 += 128;
    S.part = p2; p2 += (NT / 32) * PSTR;
    S.scal = p2; p2 += 4;
    float2* tau_s = p2; p2 += 128;
    float* dd = reinterpret_cast<float*>(p2);
    float* ee = dd + 128;
    S.red = nullptr;
    const int sig = blockIdx.x;
    if (a.skip && a.skip[sig]) return;
    const float2* T = a.Tin + (size_t)sig * d2 * d2;
    for (int idx = threadIdx.x; idx < d2 * d2; idx += NT) {
        const int c = idx / d2, r = idx % d2;
        A[r + (size_t)c * ld2] = T[idx];
    }
    __syncthreads();
    tridiag_smem<NT, PSTR>(A, d2, ld2, S, tau_s, dd, ee, a.k_stop);
    const int npk = a.d * (a.d + 1) / 2;
    const bool full = a.k_stop >= d2 - 1;
    export_tridiag(A, d2, ld2, tau_s, dd, ee, a.GV + (size_t)sig * npk, a.tau + (size_t)sig * a.d, a.dT, a.eT, a.B, sig,
                   a.k0, a.d, full ? d2 - 1 : a.k_stop, full ? d2 : a.k_stop);
    if (!full) {
        const int d3 = d2 - a.k_stop;
        tri_store_trailing(A, d2, ld2, a.k_stop, S, a.Tout + (size_t)sig * d3 * d3);
    }
}

// Debug/unit entry: tridiagonalise arbitrary Hermitian matrices given as full row-major [B][d][d]
// (lower triangle is read).
template <int NS>
__global__ void __launch_bounds__(256, NS == 8 ? 1 : 2)
k_tridiag(const float2* __restrict__ Afull, int B, int d, int ld, float2* V, float2* tau, float* dT, float* eT,
          float2* Ttr, int k1) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HeadSmem s = carve_head(smem_raw, d, ld);
    const int sig = blockIdx.x;
    const int npk = d * (d + 1) / 2;
    const float2* Ag = Afull + (size_t)sig * d * d;
    for (int idx = threadIdx.x; idx < d * d; idx += blockDim.x) {
        const int i = idx / d, j = idx % d;
        if (j <= i) {
            float2 v = Ag[idx];
            if (i == j) v.y = 0.f;
            s.A[i + (size_t)j * ld] = v;
            if (i != j) s.A[j + (size_t)i * ld] = cconj(v);
        }
    }
    __syncthreads();
    if (NS > 0) {
        tridiag_reg<(NS > 0 ? NS : 7)>(s.A, d, ld, s.S.vw, s.S.vn, s.tau, s.dd, s.ee, V + (size_t)sig * npk);
        export_de(s.tau, s.dd, s.ee, tau + (size_t)sig * d, dT, eT, B, sig, d);
        return;
    }
    tridiag_smem<256, 128>(s.A, d, ld, s.S, s.tau, s.dd, s.ee, k1);
    const bool full = k1 >= d - 1;
    export_tridiag(s.A, d, ld, s.tau, s.dd, s.ee, V + (size_t)sig * npk, tau + (size_t)sig * d, dT, eT, B, sig, 0, d,
                   full ? d - 1 : k1, full ? d : k1);
    if (!full) {
        const int d2 = d - k1;
        tri_store_trailing(s.A, d, ld, k1, s.S, Ttr + (size_t)sig * d2 * d2);
    }
}

// =====================================================================================
// k_ql: implicit QL with Wilkinson shift on the real symmetric tridiagonal, one thread per signal.
// Scalars in fp64 (the rotation parameters decide the accuracy of f(A));
// rotations are emitted as fp32 (c,s).
// Stream format per signal (float2 entries): [header (m, cnt as int bits)] [cnt x (c,s)] ... [header m=-1].
// A sweep starting at m applies rotations to column pairs (i,i+1), i = m-1, m-2, ..., m-cnt.
// =====================================================================================
#define QL_THREADS 32
#define QL_MAXIT 60

// Written as a per-lane state machine (one plane rotation per trip of a common loop) so that the 32
// signals of a warp, whose sweeps have different lengths, all make progress on every trip instead of
// waiting at the reconvergence point of divergent inner loops.  The deflation test is fused into the
// sweep (every off-diagonal of the active block is rewritten by the sweep, so the smallest negligible
// index is known when the sweep ends); a scan is only needed at the start and in the rare case of
// three simultaneous deflations.
// Divide & conquer support: the tridiagonal is torn at `ntear` rows p (Cuppen): T = diag(T1', T2') + rho v v^T,
// rho = e[p-1], d[p-1] -= rho, d[p] -= rho, e[p-1] = 0.  QL then works on the independent blocks (its sweeps
// are ~1/nblocks as long), and k_merge glues the blocks back with secular-equation solves + a small GEMM.
#define DC_MAXTEAR 7
struct TearSpec {
    int n;
    int pos[DC_MAXTEAR];
    int absconv;      // 1: d[p-1] -= |rho|, d[p] -= |rho| (the convention of k_dc's merges: rank-one term 2|rho| z z^T,
                      //    z = (last row of Q1, sign(rho) first row of Q2)/sqrt(2)); 0: signed rho (k_merge)
};
__global__ void __launch_bounds__(QL_THREADS)
k_ql(const float* __restrict__ dT, const float* __restrict__ eT, int B, int d, float* __restrict__ lam,
     float2* __restrict__ rot, int rcap, int* __restrict__ nrot, int* __restrict__ status, TearSpec tears,
     double* __restrict__ rho_out, const int* __restrict__ skip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sd = reinterpret_cast<double*>(smem_raw);
    double* se = sd + (size_t)d * QL_THREADS;
    const int t = threadIdx.x;
    const int sig = blockIdx.x * QL_THREADS + t;
    const bool valid = sig < B && !(skip && skip[sig]);
#define D_(i) sd[(i) * QL_THREADS + t]
#define E_(i) se[(i) * QL_THREADS + t]
    double anorm = 0.0;
    for (int i = 0; i < d; ++i) {
        const double di = valid ? dT[(size_t)i * B + sig] : 0.0, ei = valid ? eT[(size_t)i * B + sig] : 0.0;
        D_(i) = di;
        E_(i) = ei;
        anorm = fmax(anorm, fmax(fabs(di), fabs(ei)));
    }
    for (int q = 0; q < tears.n; ++q) {
        const int p = tears.pos[q];
        const double rho = E_(p - 1);
        const double sub = tears.absconv ? fabs(rho) : rho;
        D_(p - 1) -= sub;
        D_(p) -= sub;
        E_(p - 1) = 0.0;
        if (valid) rho_out[(size_t)sig * DC_MAXTEAR + q] = rho;
    }
    const double eps = 3.0e-8, floor_abs = 1.0e-9 * anorm;
    float2* out = rot + (size_t)(valid ? sig : 0) * rcap;
    enum { SEARCH = 0, START = 1, ROTATE = 2, DONE = 3 };
    int phase = valid ? SEARCH : DONE;
    int l = 0, m = 0, ms = 0, i = 0, iter = 0, hdr = 0, cnt = 0, nrec = 0;
    int nm1 = 0, nm2 = 0;          // two smallest negligible indices seen in the current sweep
    bool have2 = false;
    double s = 1.0, c = 1.0, p = 0.0, g = 0.0;
    bool fail = false;
    while (!__all_sync(0xffffffffu, phase == DONE)) {
        if (phase == SEARCH) {      // first negligible off-diagonal at or after ms (e[d-1] counts as 0)
            bool found = false;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (!found) {
                    if (ms >= d - 1) found = true;
                    else if (fabs(E_(ms)) <= eps * (fabs(D_(ms)) + fabs(D_(ms + 1))) + floor_abs) found = true;
                    else ++ms;
                }
            }
            if (found) { m = ms; have2 = false; phase = START; }
        }
        if (phase == START) {       // block [l, m]: converged eigenvalue or a new sweep
            if (m == l) {
                ++l;
                iter = 0;
                if (l >= d) phase = DONE;
                else if (have2 && nm2 >= l) { m = nm2; have2 = false; }   // next block end known from the last sweep
                else { ms = l; phase = SEARCH; }
            } else if (iter++ == QL_MAXIT || nrec >= rcap - 2) {
                fail = true;
                phase = DONE;
            } else {                // Wilkinson shift, sweep i = m-1 .. l
                // g = d[m] - d[l] + e/(t + sign(t) sqrt(t^2+1)), t = delta/(2e)  ==  2e^2/(delta + sign(delta) hypot(delta, 2e))
                // The shift only steers convergence (any value is a valid QL step), so it is formed in fp32.
                const double dl0 = D_(l);
                const float el = (float)E_(l), delta = (float)(D_(l + 1) - dl0);
                const float hyp = sqrtf(fmaf(delta, delta, 4.f * el * el));
                const float den = delta + copysignf(hyp, delta);
                const float corr = den != 0.f ? __fdividef(2.f * el * el, den) : 0.f;
                g = D_(m) - dl0 + (double)corr;
                s = 1.0; c = 1.0; p = 0.0;
                i = m - 1;
                hdr = nrec++;
                cnt = 0;
                nm1 = m;
                nm2 = m;
                phase = ROTATE;
            }
        } else if (phase == ROTATE) {
            const double ei = E_(i);
            const double f = s * ei, bb = c * ei;
            const double h2 = f * f + g * g;
            bool end_sweep = false;
            if (h2 == 0.0) {
                E_(i + 1) = 0.0;
                D_(i + 1) -= p;
                E_(m) = 0.0;
                nm2 = nm1; nm1 = i + 1;
                end_sweep = true;
            } else {
                // 1/sqrt(h2): fp32 seed (2^-22) + one fp64 Newton step (-> 1e-13) instead of a sqrt and two divisions;
                // (c, s) leave as fp32 and the d/e recurrences only need to stay well below fp32 rounding
                double rinv;
                if (h2 > 1e-30 && h2 < 1e30) {
                    rinv = (double)rsqrtf((float)h2);
                    rinv = rinv * (1.5 - 0.5 * h2 * rinv * rinv);
                } else {
                    rinv = 1.0 / sqrt(h2);
                }
                const double enew = h2 * rinv;
                E_(i + 1) = enew;
                s = f * rinv;
                c = g * rinv;
                const double di1 = D_(i + 1);
                g = di1 - p;
                const double r = (D_(i) - g) * s + 2.0 * c * bb;
                p = s * r;
                const double dnew = g + p;
                D_(i + 1) = dnew;
                g = c * r - bb;
                // fused deflation test for e[i+1] (final for this sweep; e[m] itself is zeroed below)
                if (i + 1 < m && enew <= eps * (fabs(dnew) + fabs(D_(i + 2))) + floor_abs) { nm2 = nm1; nm1 = i + 1; }
                if (nrec < rcap - 1) out[nrec] = make_float2((float)c, (float)s);
                ++nrec;
                ++cnt;
                --i;
                if (i < l) {
                    const double dl = D_(l) - p;
                    D_(l) = dl;
                    E_(l) = g;
                    E_(m) = 0.0;
                    if (fabs(g) <= eps * (fabs(dl) + fabs(D_(l + 1))) + floor_abs) { nm2 = nm1; nm1 = l; }
                    end_sweep = true;
                }
            }
            if (end_sweep) {
                if (hdr < rcap - 1) out[hdr] = make_float2(__int_as_float(m), __int_as_float(cnt));
                if (nrec >= rcap - 1) { fail = true; phase = DONE; }
                else { m = nm1; have2 = true; phase = START; }       // smallest negligible index in [l, old m]
            }
        }
    }
    if (!valid) return;
    if (fail) {
        atomicOr(status, 1);
        nrec = 0;   // consumer leaves Z = I; the status word is the error report
    }
    out[nrec] = make_float2(__int_as_float(-1), __int_as_float(0));
    nrot[sig] = nrec + 1;
    for (int j = 0; j < d; ++j) lam[(size_t)sig * d + j] = (float)D_(j);
#undef D_
#undef E_
}

// =====================================================================================
// k_rot: Z = I * (product of the recorded plane rotations).  One CTA per signal, thread = row of Z.
// The rotation stream is staged through shared memory in double-buffered cp.async chunks.
// =====================================================================================
#define ROT_THREADS 32
#define ROT_CHUNK 1024   // granularity of rcap (entries)
#define ROT_STAGE 512    // float2 entries per cp.async stage (4 KB)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void rot4(float4& out, float4& cy, const float4 zi, const float2 e) {
    // z[i+1] = s*z[i] + c*f ;  z[i] = c*z[i] - s*f   (f = carry), four rows at once as two packed pairs
    // (FMUL2 + FFMA2: 8 issue slots instead of 16, same rounding)
    const f32x2 c2 = bc2(e.x), s2 = bc2(e.y), ns2 = bc2(-e.y);
    const f32x2 z01 = pk2(zi.x, zi.y), z23 = pk2(zi.z, zi.w);
    const f32x2 f01 = pk2(cy.x, cy.y), f23 = pk2(cy.z, cy.w);
    const float2 o01 = upk2(fma2(s2, z01, mul2(c2, f01))), o23 = upk2(fma2(s2, z23, mul2(c2, f23)));
    const float2 n01 = upk2(fma2(c2, z01, mul2(ns2, f01))), n23 = upk2(fma2(c2, z23, mul2(ns2, f23)));
    out = make_float4(o01.x, o01.y, o23.x, o23.y);
    cy = make_float4(n01.x, n01.y, n23.x, n23.y);
}

// One WARP per signal (CTA = 1 warp, ~46 KB of shared memory -> 4-5 CTAs per SM).  Z is kept TRANSPOSED in
// shared memory, zt[col][row] with ldr = 4*ceil(d/4) rows per column; lane l owns the four adjacent rows
// 4l..4l+3, so one rotation is one 128-bit load, 16 FP ops and one 128-bit store per lane, and one
// broadcast load of (c,s) feeds four independent chains.  Output: Z^T, i.e. Zt[c][r] row-major [d][d].
__global__ void __launch_bounds__(ROT_THREADS)
k_rot(const float2* __restrict__ rot, int rcap, const int* __restrict__ nrot, int d, float* __restrict__ Zt,
      const int* __restrict__ skip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* stage = reinterpret_cast<float2*>(smem_raw);                  // [2][ROT_STAGE]
    float* z = reinterpret_cast<float*>(stage + 2 * ROT_STAGE);           // [d][ldr]
    const int ldr = 4 * ((d + 3) / 4);
    const int lane = threadIdx.x;
    const int sig = blockIdx.x;
    if (skip && skip[sig]) return;
    const float2* src = rot + (size_t)sig * rcap;
    const int total = nrot[sig];
    for (int idx = lane; idx < d * ldr; idx += 32) z[idx] = 0.f;
    __syncwarp();
    for (int r = lane; r < d; r += 32) z[r * ldr + r] = 1.f;
    const int nchunks = (total + ROT_STAGE - 1) / ROT_STAGE;
    auto issue = [&](int ch) {
        const float2* g = src + (size_t)ch * ROT_STAGE;
        float2* sdst = stage + (ch & 1) * ROT_STAGE;
        // whole stages are always readable: rcap is a multiple of ROT_CHUNK (checked on the host)
#pragma unroll
        for (int q = lane; q < ROT_STAGE / 2; q += 32) cp_async16(sdst + 2 * q, g + 2 * q);
        cp_async_commit();
    };
    if (nchunks > 0) issue(0);
    const bool act = 4 * lane < ldr;
    float4* zq = reinterpret_cast<float4*>(z) + (act ? lane : 0);        // column c at zq[c * (ldr/4)]
    const int lq = ldr / 4;
    float4 carry = make_float4(0.f, 0.f, 0.f, 0.f);
    int remaining = 0, col = 0;
    bool done = false;
    for (int ch = 0; ch < nchunks; ++ch) {
        if (ch + 1 < nchunks) { issue(ch + 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncwarp();
        const float2* sbuf = stage + (ch & 1) * ROT_STAGE;
        const int cnt_here = min(ROT_STAGE, total - ch * ROT_STAGE);
        int q = 0;
        while (!done && q < cnt_here) {
            if (remaining == 0) {
                const float2 e = sbuf[q++];
                const int m = __float_as_int(e.x);
                if (m < 0) { done = true; break; }
                remaining = __float_as_int(e.y);
                col = m;                                    // carry holds z[.][col]
                carry = zq[col * lq];
                continue;
            }
            const int run = min(remaining, cnt_here - q);
            int t = 0;
            // batches of 4 rotations: loads first (independent of the carry chain), then the four
            // dependent rotations, then the stores
            for (; t + 4 <= run; t += 4) {
                const float2 e0 = sbuf[q + t], e1 = sbuf[q + t + 1], e2 = sbuf[q + t + 2], e3 = sbuf[q + t + 3];
                float4* zp = zq + (col - 1 - t) * lq;       // column ci = col-1-t
                const float4 z0 = zp[0], z1 = zp[-lq], z2 = zp[-2 * lq], z3 = zp[-3 * lq];
                float4 o0, o1, o2, o3;
                rot4(o0, carry, z0, e0);
                rot4(o1, carry, z1, e1);
                rot4(o2, carry, z2, e2);
                rot4(o3, carry, z3, e3);
                if (act) { zp[lq] = o0; zp[0] = o1; zp[-lq] = o2; zp[-2 * lq] = o3; }
            }
            for (; t < run; ++t) {
                const float2 e = sbuf[q + t];               // (c, s)
                float4* zp = zq + (col - 1 - t) * lq;
                const float4 zi = zp[0];
                float4 o;
                rot4(o, carry, zi, e);
                if (act) zp[lq] = o;
            }
            q += run;
            col -= run;
            remaining -= run;
            if (remaining == 0 && act) zq[col * lq] = carry;
        }
        __syncwarp();   // stage buffer (ch&1) is refilled by issue(ch+2)
    }
    __syncwarp();
    // global Z^T keeps the padded row pitch ldr (16-byte aligned rows: TMA box source for k_tail_tc)
    float4* outz = reinterpret_cast<float4*>(Zt + (size_t)sig * d * ldr);
    const float4* z4 = reinterpret_cast<const float4*>(z);
    for (int idx = lane; idx < d * ldr / 4; idx += 32) outz[idx] = z4[idx];
}

// k_rotf: k_rot with PAIRS of consecutive sweeps fused.  k_rot is bound by shared-memory bandwidth (one 128-bit load
// and one 128-bit store of a column of Z^T per rotation: 9 wavefronts of the 128 B/clk pipe, 4 signals per SM keep it
// saturated).  Sweep B follows sweep A over (almost) the same columns, and rotation b_j only needs a_{j-1} to be
// done, so B can trail A by one column with both carries in registers:
//     step i:  load col i;  a_i on (col i, carryA = col i+1) -> outA = final col i+1 of A;
//              b_{i+1} on (outA, carryB = col i+2) -> store col i+2;  carryB = new col i+1
// i.e. one load and one store per TWO rotations.  Steps outside a sweep's own range are exact pass-throughs, so any
// two ranges can be fused; pairs whose union is not shorter than 85 % of the two sweeps processed apart are not.
// The stream is staged through a ring of four 256-entry buffers (two landed ahead of the read position, one in
// flight), so both sweeps of a pair (<= 2(d+1) entries) are always resident.
#define ROTF_STG 256
__global__ void __launch_bounds__(ROT_THREADS)
k_rotf(const float2* __restrict__ rot, int rcap, const int* __restrict__ nrot, int d, float* __restrict__ Zt,
       const int* __restrict__ skip) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);                   // [4][ROTF_STG]
    float* z = reinterpret_cast<float*>(ring + 4 * ROTF_STG);             // [d][ldr]
    const int ldr = 4 * ((d + 3) / 4);
    const int lane = threadIdx.x;
    const int sig = blockIdx.x;
    if (skip && skip[sig]) return;
    const float2* src = rot + (size_t)sig * rcap;
    const int total = nrot[sig];
    for (int idx = lane; idx < d * ldr; idx += 32) z[idx] = 0.f;
    __syncwarp();
    for (int r = lane; r < d; r += 32) z[r * ldr + r] = 1.f;
    const int nstages = (total + ROTF_STG - 1) / ROTF_STG;
    auto issue = [&](int k) {                                             // stage k -> ring slot k & 3 (one group)
        if (k < nstages) {
            const float2* g = src + (size_t)k * ROTF_STG;
            float2* sdst = ring + (k & 3) * ROTF_STG;
#pragma unroll
            for (int q = lane; q < ROTF_STG / 2; q += 32) cp_async16(sdst + 2 * q, g + 2 * q);
        }
        cp_async_commit();
    };
    issue(0); issue(1); issue(2);
    cp_async_wait<1>();                                                   // stages 0 and 1 have landed
    __syncwarp();
    int cur = 0;
    const bool act = 4 * lane < ldr;
    float4* zq = reinterpret_cast<float4*>(z) + (act ? lane : 0);        // column c at zq[c * lq]
    const int lq = ldr / 4;
#define RING(e) ring[(e) & (4 * ROTF_STG - 1)]
    int q = 0;
    while (q < total) {
        while ((q / ROTF_STG) > cur) {                                    // entered a new stage: keep two landed ahead
            ++cur;
            __syncwarp();                                                 // every lane is done with stage cur-1
            issue(cur + 2);
            cp_async_wait<1>();
            __syncwarp();
        }
        const float2 hA = RING(q);
        const int mA = __float_as_int(hA.x);
        if (mA < 0) break;
        const int cntA = __float_as_int(hA.y);
        const int qa = q + 1;
        const int qB = qa + cntA;
        bool pair = false;
        int mB = 0, cntB = 0;
        if (cntA > 0 && qB < total) {
            const float2 hB = RING(qB);
            mB = __float_as_int(hB.x);
            cntB = __float_as_int(hB.y);
            if (mB >= 0 && cntB > 0) {
                const int top = max(mA, mB) - 1, end = min(mA - cntA, mB - cntB - 1);
                pair = 20 * (top - end + 1) <= 17 * (cntA + cntB);
            }
        }
        if (pair) {
            const int qb = qB + 1;
            const int lA = mA - cntA, lB = mB - cntB;
            const int itop = max(mA, mB) - 1, iend = min(lA, lB - 1);
            float4 cA = zq[(itop + 1) * lq], cB;
            {   // first step: A only, its output is B's first carry
                const float4 zi = zq[itop * lq];
                if (itop <= mA - 1 && itop >= lA) {
                    rot4(cB, cA, zi, RING(qa + mA - 1 - itop));
                } else { cB = cA; cA = zi; }
            }
            int i = itop - 1;
            // core: both sweeps active, two steps per trip, loads first
            const int core_lo = max(lA, lB - 1);
            const int core_hi = min(mA - 1, mB - 2);
            // flagged steps above the core
            for (; i >= iend && i > core_hi; --i) {
                const float4 zi = i >= 0 ? zq[i * lq] : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 oA, oB;
                if (i <= mA - 1 && i >= lA) rot4(oA, cA, zi, RING(qa + mA - 1 - i)); else { oA = cA; cA = zi; }
                if (i + 1 <= mB - 1 && i + 1 >= lB) rot4(oB, cB, oA, RING(qb + mB - 2 - i)); else { oB = cB; cB = oA; }
                if (act) zq[(i + 2) * lq] = oB;
            }
            // (requesting the next trip's operands ahead of the FMA chain was measured slower: 376 -> 426 ms)
            {
                // fast form of the core loop: when neither sweep's parameters wrap around the ring inside this pair
                // (4 pairs out of 5) plain pointers replace the masked index arithmetic (5 instead of 24 integer
                // instructions per trip, and a lone warp pays two cycles for every instruction it issues)
                const int ntr = i - 1 >= core_lo ? (i - core_lo + 1) / 2 : 0;
                const int ia = (qa + mA - 1 - i) & (4 * ROTF_STG - 1), ib = (qb + mB - 2 - i) & (4 * ROTF_STG - 1);
                if (ntr > 0 && ia + 2 * ntr <= 4 * ROTF_STG && ib + 2 * ntr <= 4 * ROTF_STG) {
                    const float2* pa = ring + ia;
                    const float2* pb = ring + ib;
                    float4* zp = zq + i * lq;
                    // two register sets: the loads of trip t+1 are issued before the FMA chain of trip t (pure
                    // reordering, no extra instructions; a load past the last trip reads valid, unused shared memory)
#define ROTF_LD(S, K)                                                                          \
    const float2 ea0##S = pa[2 * (K)], ea1##S = pa[2 * (K) + 1], eb0##S = pb[2 * (K)], eb1##S = pb[2 * (K) + 1]; \
    const float4 z0##S = zp[-2 * (K) * lq], z1##S = zp[-(2 * (K) + 1) * lq];
#define ROTF_DO(S, K)                                                                          \
    {                                                                                          \
        float4 oA0, oB0, oA1, oB1;                                                             \
        rot4(oA0, cA, z0##S, ea0##S);                                                          \
        rot4(oB0, cB, oA0, eb0##S);                                                            \
        rot4(oA1, cA, z1##S, ea1##S);                                                          \
        rot4(oB1, cB, oA1, eb1##S);                                                            \
        if (act) { zp[(2 - 2 * (K)) * lq] = oB0; zp[(1 - 2 * (K)) * lq] = oB1; }               \
    }
                    int tr = 0;
                    for (; tr + 4 <= ntr; tr += 4) {
                        ROTF_LD(X, 0)
                        ROTF_LD(Y, 1)
                        ROTF_LD(V, 2)
                        ROTF_LD(W, 3)
                        ROTF_DO(X, 0)
                        ROTF_DO(Y, 1)
                        ROTF_DO(V, 2)
                        ROTF_DO(W, 3)
                        pa += 8; pb += 8; zp -= 8 * lq;
                    }
                    if (tr + 2 <= ntr) {
                        ROTF_LD(X, 0)
                        ROTF_LD(Y, 1)
                        ROTF_DO(X, 0)
                        ROTF_DO(Y, 1)
                        pa += 4; pb += 4; zp -= 4 * lq;
                        tr += 2;
                    }
                    if (tr < ntr) {
                        ROTF_LD(X, 0)
                        ROTF_DO(X, 0)
                    }
#undef ROTF_LD
#undef ROTF_DO
                    i -= 2 * ntr;
                }
            }
            for (; i - 1 >= core_lo; i -= 2) {
                const float2 ea0 = RING(qa + mA - 1 - i), ea1 = RING(qa + mA - i);
                const float2 eb0 = RING(qb + mB - 2 - i), eb1 = RING(qb + mB - 1 - i);
                const float4 z0 = zq[i * lq], z1 = zq[(i - 1) * lq];
                float4 oA0, oB0, oA1, oB1;
                rot4(oA0, cA, z0, ea0);
                rot4(oB0, cB, oA0, eb0);
                rot4(oA1, cA, z1, ea1);
                rot4(oB1, cB, oA1, eb1);
                if (act) { zq[(i + 2) * lq] = oB0; zq[(i + 1) * lq] = oB1; }
            }
            // flagged steps below (and the odd core step)
            for (; i >= iend; --i) {
                const float4 zi = i >= 0 ? zq[i * lq] : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 oA, oB;
                if (i <= mA - 1 && i >= lA) rot4(oA, cA, zi, RING(qa + mA - 1 - i)); else { oA = cA; cA = zi; }
                if (i + 1 <= mB - 1 && i + 1 >= lB) rot4(oB, cB, oA, RING(qb + mB - 2 - i)); else { oB = cB; cB = oA; }
                if (act) zq[(i + 2) * lq] = oB;
            }
            if (act) {
                zq[(iend + 1) * lq] = cB;
                if (iend >= 0) zq[iend * lq] = cA;
            }
            q = qb + cntB;
        } else {
            // single sweep (as k_rot): batches of 4 rotations, loads first
            int col = mA;
            float4 carry = zq[col * lq];
            int t = 0;
            for (; t + 4 <= cntA; t += 4) {
                const float2 e0 = RING(qa + t), e1 = RING(qa + t + 1), e2 = RING(qa + t + 2), e3 = RING(qa + t + 3);
                float4* zp = zq + (col - 1 - t) * lq;
                const float4 z0 = zp[0], z1 = zp[-lq], z2 = zp[-2 * lq], z3 = zp[-3 * lq];
                float4 o0, o1, o2, o3;
                rot4(o0, carry, z0, e0);
                rot4(o1, carry, z1, e1);
                rot4(o2, carry, z2, e2);
                rot4(o3, carry, z3, e3);
                if (act) { zp[lq] = o0; zp[0] = o1; zp[-lq] = o2; zp[-2 * lq] = o3; }
            }
            for (; t < cntA; ++t) {
                const float2 e = RING(qa + t);
                float4* zp = zq + (col - 1 - t) * lq;
                const float4 zi = zp[0];
                float4 o;
                rot4(o, carry, zi, e);
                if (act) zp[lq] = o;
            }
            if (act) zq[(col - cntA) * lq] = carry;
            q = qB;
        }
    }
#undef RING
    __syncwarp();
    // global Z^T keeps the padded row pitch ldr (16-byte aligned rows: TMA box source for k_tail_tc)
    float4* outz = reinterpret_cast<float4*>(Zt + (size_t)sig * d * ldr);
    const float4* z4 = reinterpret_cast<const float4*>(z);
    for (int idx = lane; idx < d * ldr / 4; idx += 32) outz[idx] = z4[idx];
}

// ---- k_rotf_p: k_rotf with the coordinates of a signal split over SPLIT one-warp CTAs (VW coordinates per lane).  A plane
// rotation mixes two rows of Z^T coordinate by coordinate, so panels of coordinates are independent; a panel needs
// 1/SPLIT of the shared memory (7 CTAs per SM instead of 4 at SPLIT = 2) and half the packed operations per rotation,
// which matters because a lone warp on an SM sub-partition issues at most every other cycle: more resident warps, each
// with fewer instructions per rotation.  Every CTA of a signal streams the whole rotation list (L2 hits).
template <int VW> struct RotVec;
template <> struct RotVec<4> { using type = float4; };
template <> struct RotVec<2> { using type = float2; };
__device__ __forceinline__ void rotv(float4& out, float4& cy, const float4 zi, const float2 e) { rot4(out, cy, zi, e); }
__device__ __forceinline__ void rotv(float2& out, float2& cy, const float2 zi, const float2 e) {
    const f32x2 c2 = bc2(e.x), s2 = bc2(e.y), ns2 = bc2(-e.y);
    const f32x2 z = pk2(zi.x, zi.y), f = pk2(cy.x, cy.y);
    out = upk2(fma2(s2, z, mul2(c2, f)));
    cy = upk2(fma2(c2, z, mul2(ns2, f)));
}
template <int VW, int SPLIT>
__global__ void __launch_bounds__(ROT_THREADS)
k_rotf_p(const float2* __restrict__ rot, int rcap, const int* __restrict__ nrot, int d, float* __restrict__ Zt,
         const int* __restrict__ skip) {
    using VT = typename RotVec<VW>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* ring = reinterpret_cast<float2*>(smem_raw);                   // [4][ROTF_STG]
    float* z = reinterpret_cast<float*>(ring + 4 * ROTF_STG);             // [d][ldp]: this CTA's panel of coordinates
    const int ldr = 4 * ((d + 3) / 4);
    const int ldp = ldr / SPLIT;
    const int lane = threadIdx.x;
    const int sig = blockIdx.x / SPLIT, part = blockIdx.x % SPLIT;
    if (skip && skip[sig]) return;
    const float2* src = rot + (size_t)sig * rcap;
    const int total = nrot[sig];
    for (int idx = lane; idx < d * ldp; idx += 32) z[idx] = 0.f;
    __syncwarp();
    for (int r = lane; r < d; r += 32) {
        const int x = r - part * ldp;
        if (x >= 0 && x < ldp) z[r * ldp + x] = 1.f;
    }
    const int nstages = (total + ROTF_STG - 1) / ROTF_STG;
    auto issue = [&](int k) {                                             // stage k -> ring slot k & 3 (one group)
        if (k < nstages) {
            const float2* g = src + (size_t)k * ROTF_STG;
            float2* sdst = ring + (k & 3) * ROTF_STG;
#pragma unroll
            for (int q = lane; q < ROTF_STG / 2; q += 32) cp_async16(sdst + 2 * q, g + 2 * q);
        }
        cp_async_commit();
    };
    issue(0); issue(1); issue(2);
    cp_async_wait<1>();                                                   // stages 0 and 1 have landed
    __syncwarp();
    int cur = 0;
    const bool act = VW * lane < ldp;
    VT* zq = reinterpret_cast<VT*>(z) + (act ? lane : 0);                // column c at zq[c * lq]
    const int lq = ldp / VW;
#define RING(e) ring[(e) & (4 * ROTF_STG - 1)]
    int q = 0;
    while (q < total) {
        while ((q / ROTF_STG) > cur) {                                    // entered a new stage: keep two landed ahead
            ++cur;
            __syncwarp();                                                 // every lane is done with stage cur-1
            issue(cur + 2);
            cp_async_wait<1>();
            __syncwarp();
        }
        const float2 hA = RING(q);
        const int mA = __float_as_int(hA.x);
        if (mA < 0) break;
        const int cntA = __float_as_int(hA.y);
        const int qa = q + 1;
        const int qB = qa + cntA;
        bool pair = false;
        int mB = 0, cntB = 0;
        if (cntA > 0 && qB < total) {
            const float2 hB = RING(qB);
            mB = __float_as_int(hB.x);
            cntB = __float_as_int(hB.y);
            if (mB >= 0 && cntB > 0) {
                const int top = max(mA, mB) - 1, end = min(mA - cntA, mB - cntB - 1);
                pair = 20 * (top - end + 1) <= 17 * (cntA + cntB);
            }
        }
        if (pair) {
            const int qb = qB + 1;
            const int lA = mA - cntA, lB = mB - cntB;
            const int itop = max(mA, mB) - 1, iend = min(lA, lB - 1);
            VT cA = zq[(itop + 1) * lq], cB;
            {   // first step: A only, its output is B's first carry
                const VT zi = zq[itop * lq];
                if (itop <= mA - 1 && itop >= lA) {
                    rotv(cB, cA, zi, RING(qa + mA - 1 - itop));
                } else { cB = cA; cA = zi; }
            }
            int i = itop - 1;
            // core: both sweeps active, two steps per trip, loads first
            const int core_lo = max(lA, lB - 1);
            const int core_hi = min(mA - 1, mB - 2);
            // flagged steps above the core
            for (; i >= iend && i > core_hi; --i) {
                const VT zi = i >= 0 ? zq[i * lq] : VT{};
                VT oA, oB;
                if (i <= mA - 1 && i >= lA) rotv(oA, cA, zi, RING(qa + mA - 1 - i)); else { oA = cA; cA = zi; }
                if (i + 1 <= mB - 1 && i + 1 >= lB) rotv(oB, cB, oA, RING(qb + mB - 2 - i)); else { oB = cB; cB = oA; }
                if (act) zq[(i + 2) * lq] = oB;
            }
            // (requesting the next trip's operands ahead of the FMA chain was measured slower: 376 -> 426 ms)
            {
                // fast form of the core loop: when neither sweep's parameters wrap around the ring inside this pair
                // (4 pairs out of 5) plain pointers replace the masked index arithmetic (5 instead of 24 integer
                // instructions per trip, and a lone warp pays two cycles for every instruction it issues)
                const int ntr = i - 1 >= core_lo ? (i - core_lo + 1) / 2 : 0;
                const int ia = (qa + mA - 1 - i) & (4 * ROTF_STG - 1), ib = (qb + mB - 2 - i) & (4 * ROTF_STG - 1);
                if (ntr > 0 && ia + 2 * ntr <= 4 * ROTF_STG && ib + 2 * ntr <= 4 * ROTF_STG) {
                    const float2* pa = ring + ia;
                    const float2* pb = ring + ib;
                    VT* zp = zq + i * lq;
                    // two register sets: the loads of trip t+1 are issued before the FMA chain of trip t (pure
                    // reordering, no extra instructions; a load past the last trip reads valid, unused shared memory)
#define ROTF_LD(S, K)                                                                          \
    const float2 ea0##S = pa[2 * (K)], ea1##S = pa[2 * (K) + 1], eb0##S = pb[2 * (K)], eb1##S = pb[2 * (K) + 1]; \
    const VT z0##S = zp[-2 * (K) * lq], z1##S = zp[-(2 * (K) + 1) * lq];
#define ROTF_DO(S, K)                                                                          \
    {                                                                                          \
        VT oA0, oB0, oA1, oB1;                                                             \
        rotv(oA0, cA, z0##S, ea0##S);                                                          \
        rotv(oB0, cB, oA0, eb0##S);                                                            \
        rotv(oA1, cA, z1##S, ea1##S);                                                          \
        rotv(oB1, cB, oA1, eb1##S);                                                            \
        if (act) { zp[(2 - 2 * (K)) * lq] = oB0; zp[(1 - 2 * (K)) * lq] = oB1; }               \
    }
                    int tr = 0;
                    for (; tr + 4 <= ntr; tr += 4) {
                        ROTF_LD(X, 0)
                        ROTF_LD(Y, 1)
                        ROTF_LD(V, 2)
                        ROTF_LD(W, 3)
                        ROTF_DO(X, 0)
                        ROTF_DO(Y, 1)
                        ROTF_DO(V, 2)
                        ROTF_DO(W, 3)
                        pa += 8; pb += 8; zp -= 8 * lq;
                    }
                    if (tr + 2 <= ntr) {
                        ROTF_LD(X, 0)
                        ROTF_LD(Y, 1)
                        ROTF_DO(X, 0)
                        ROTF_DO(Y, 1)
                        pa += 4; pb += 4; zp -= 4 * lq;
                        tr += 2;
                    }
                    if (tr < ntr) {
                        ROTF_LD(X, 0)
                        ROTF_DO(X, 0)
                    }
#undef ROTF_LD
#undef ROTF_DO
                    i -= 2 * ntr;
                }
            }
            for (; i - 1 >= core_lo; i -= 2) {
                const float2 ea0 = RING(qa + mA - 1 - i), ea1 = RING(qa + mA - i);
                const float2 eb0 = RING(qb + mB - 2 - i), eb1 = RING(qb + mB - 1 - i);
                const VT z0 = zq[i * lq], z1 = zq[(i - 1) * lq];
                VT oA0, oB0, oA1, oB1;
                rotv(oA0, cA, z0, ea0);
                rotv(oB0, cB, oA0, eb0);
                rotv(oA1, cA, z1, ea1);
                rotv(oB1, cB, oA1, eb1);
                if (act) { zq[(i + 2) * lq] = oB0; zq[(i + 1) * lq] = oB1; }
            }
            // flagged steps below (and the odd core step)
            for (; i >= iend; --i) {
                const VT zi = i >= 0 ? zq[i * lq] : VT{};
                VT oA, oB;
                if (i <= mA - 1 && i >= lA) rotv(oA, cA, zi, RING(qa + mA - 1 - i)); else { oA = cA; cA = zi; }
                if (i + 1 <= mB - 1 && i + 1 >= lB) rotv(oB, cB, oA, RING(qb + mB - 2 - i)); else { oB = cB; cB = oA; }
                if (act) zq[(i + 2) * lq] = oB;
            }
            if (act) {
                zq[(iend + 1) * lq] = cB;
                if (iend >= 0) zq[iend * lq] = cA;
            }
            q = qb + cntB;
        } else {
            // single sweep (as k_rot): batches of 4 rotations, loads first
            int col = mA;
            VT carry = zq[col * lq];
            int t = 0;
            for (; t + 4 <= cntA; t += 4) {
                const float2 e0 = RING(qa + t), e1 = RING(qa + t + 1), e2 = RING(qa + t + 2), e3 = RING(qa + t + 3);
                VT* zp = zq + (col - 1 - t) * lq;
                const VT z0 = zp[0], z1 = zp[-lq], z2 = zp[-2 * lq], z3 = zp[-3 * lq];
                VT o0, o1, o2, o3;
                rotv(o0, carry, z0, e0);
                rotv(o1, carry, z1, e1);
                rotv(o2, carry, z2, e2);
                rotv(o3, carry, z3, e3);
                if (act) { zp[lq] = o0; zp[0] = o1; zp[-lq] = o2; zp[-2 * lq] = o3; }
            }
            for (; t < cntA; ++t) {
                const float2 e = RING(qa + t);
                VT* zp = zq + (col - 1 - t) * lq;
                const VT zi = zp[0];
                VT o;
                rotv(o, carry, zi, e);
                if (act) zp[lq] = o;
            }
            if (act) zq[(col - cntA) * lq] = carry;
            q = qB;
        }
    }
#undef RING
    __syncwarp();
    // global Z^T keeps the padded row pitch ldr (16-byte aligned rows: TMA box source for k_tail_tc)
    float* outz = Zt + (size_t)sig * d * ldr + part * ldp;
    const VT* zv = reinterpret_cast<const VT*>(z);
    for (int idx = lane; idx < d * lq; idx += 32) {
        const int row = idx / lq, v = idx - row * lq;
        *reinterpret_cast<VT*>(outz + (size_t)row * ldr + v * VW) = zv[idx];
    }
}

// =====================================================================================
// k_merge: one level of the divide & conquer merge tree.  Every range [a,b) of the level is the union of two
// already-solved blocks torn at row p; eigenpairs of diag(D) + rho z z^T (z = Q^T v, poles D = block eigenvalues
// rounded to fp32 so that pole differences are exact in fp64) come from the secular equation
//     1 + rho * sum_i z_i^2 / (d_i - lambda) = 0
// solved in fp64 in coordinates shifted to the nearer pole (safeguarded Newton), and the new eigenvectors are
// Q * W with W_ij = z_i / (d_i - lambda_j), a small fp32 GEMM.  Deflation: |z_i| negligible -> pair kept;
// exactly equal poles -> Givens rotation of the two columns.  One CTA per signal; Zt is Z^T ([c][r]).
// =====================================================================================
#define MG_THREADS 256
#define MG_MAXR 4
struct MergeArgs {
    const float* Zin;     // [B][d][d]
    float* Zout;          // [B][d][d]
    float* lam;           // [B][d] in/out
    const double* rho;    // [B][DC_MAXTEAR]
    int* status;
    int B, d, nr;
    int ra[MG_MAXR], rp[MG_MAXR], rb[MG_MAXR], rt[MG_MAXR];   // range [a,b), tear row p, tear index
    const int* skip;
};
__host__ __device__ inline size_t merge_smem_bytes(int d) {
    const int ldr = 4 * ((d + 3) / 4);
    return (size_t)2 * d * ldr * sizeof(float) + (size_t)8 * 128 * sizeof(double) + (size_t)8 * 128 * sizeof(int) + 256;   // (see carve in k_merge)
}

// Secular function pieces evaluated by a PAIR of lanes (lane parity `par` takes every other term, results combined
// with one shuffle): psi = sum_{i<=u} z_i^2/((d_i-org)-mu), phi = sum_{i>u}, and their derivatives.  T = float for
// the bulk of the iteration (the pole differences are still formed in fp64, then rounded), T = double for the polish.
struct SecVal { double psi, dpsi, phi, dphi; };
__device__ __forceinline__ float sec_rcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double sec_rcp(double x) {
    // 1/x: fp32 seed + two Newton steps (2^-23 -> 2^-46 -> below 2^-53); |x| stays well inside the fp32 range here
    // except next to a pole, where the plain division is used
    if (fabs(x) > 1e-30 && fabs(x) < 1e30) {
        double y = (double)__frcp_rn((float)x);
        y = y * (2.0 - x * y);
        y = y * (2.0 - x * y);
        return y;
    }
    return 1.0 / x;
}
template <typename T>
__device__ __forceinline__ SecVal secular_eval(const double* __restrict__ ksd, const double* __restrict__ ksz2, int k, int u,
                                               double org, double mu, int par) {
    T p0 = 0, p1 = 0, q0 = 0, q1 = 0;
    const int nl = k > 0 ? u + 1 : 0;
    int i = par;
    for (; i + 2 < nl; i += 4) {
        const T r0 = sec_rcp((T)((ksd[i] - org) - mu)), r1 = sec_rcp((T)((ksd[i + 2] - org) - mu));
        const T t0 = (T)ksz2[i] * r0, t1 = (T)ksz2[i + 2] * r1;
        p0 += t0; p1 += t1;
        q0 += t0 * r0; q1 += t1 * r1;
    }
    for (; i < nl; i += 2) {
        const T r0 = sec_rcp((T)((ksd[i] - org) - mu));
        const T t0 = (T)ksz2[i] * r0;
        p0 += t0; q0 += t0 * r0;
    }
    T f0 = 0, f1 = 0, g0 = 0, g1 = 0;
    for (; i + 2 < k; i += 4) {
        const T r0 = sec_rcp((T)((ksd[i] - org) - mu)), r1 = sec_rcp((T)((ksd[i + 2] - org) - mu));
        const T t0 = (T)ksz2[i] * r0, t1 = (T)ksz2[i + 2] * r1;
        f0 += t0; f1 += t1;
        g0 += t0 * r0; g1 += t1 * r1;
    }
    for (; i < k; i += 2) {
        const T r0 = sec_rcp((T)((ksd[i] - org) - mu));
        const T t0 = (T)ksz2[i] * r0;
        f0 += t0; g0 += t0 * r0;
    }
    SecVal v;
    v.psi = (double)(p0 + p1); v.dpsi = (double)(q0 + q1); v.phi = (double)(f0 + f1); v.dphi = (double)(g0 + g1);
    v.psi += __shfl_xor_sync(0xffffffffu, v.psi, 1);
    v.dpsi += __shfl_xor_sync(0xffffffffu, v.dpsi, 1);
    v.phi += __shfl_xor_sync(0xffffffffu, v.phi, 1);
    v.dphi += __shfl_xor_sync(0xffffffffu, v.dphi, 1);
    return v;
}
// One step of the two-pole rational interpolation (Bunch-Nielsen-Sorensen; what LAPACK's xLAED4 iterates): psi is
// modelled by a + s/(dl-x), phi by b + S/(dr-x), matching value and slope at mu; the model's root in (dl,dr) is
// the next iterate (monotone, quadratic).  `lastroot`: no pole to the right, one-pole model.
__device__ __forceinline__ double secular_step(const SecVal& v, double rho, double f, double mu, double dl, double dr,
                                               bool lastroot) {
    const double Dl = dl - mu, Dr = dr - mu;
    if (lastroot) {
        const double s = rho * v.dpsi * Dl * Dl, c0 = f - rho * v.dpsi * Dl;
        return Dl + s / c0;
    }
    const double s = rho * v.dpsi * Dl * Dl, S = rho * v.dphi * Dr * Dr;
    const double c0 = f - rho * v.dpsi * Dl - rho * v.dphi * Dr;
    const double A = c0, Bq = c0 * (Dl + Dr) + s + S, Cq = Dl * Dr * f;
    double disc = Bq * Bq - 4.0 * A * Cq;
    disc = disc > 0.0 ? sqrt(disc) : 0.0;
    if (A == 0.0) return Cq / Bq;
    return Bq > 0.0 ? 2.0 * Cq / (Bq + disc) : (Bq - disc) / (2.0 * A);
}

__global__ void __launch_bounds__(MG_THREADS, 2) k_merge(MergeArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d, ldr = 4 * ((d + 3) / 4);
    float* Zs = reinterpret_cast<float*>(smem_raw);                 // [d][ldr]  Z^T
    float* Ws = Zs + (size_t)d * ldr;                               // [d][ldr]  block-diagonal W, Ws[u][t]
    double* dbl = reinterpret_cast<double*>(Ws + (size_t)d * ldr);
    double* sd = dbl;             // [128] sorted poles
    double* sz = dbl + 128;       // [128] sorted z
    double* ksd = dbl + 256;      // [128] compact (non-deflated) poles
    double* ksz = dbl + 384;      // [128] compact z
    double* kmu = dbl + 512;      // [128] root offset from its origin pole
    double* korg = dbl + 640;     // [128] origin pole value of each root
    double* kwin = dbl + 768;     // [128] 1/||w||
    double* rsc = dbl + 896;      // [128] per-range scalars: rho_eff[4], sign[4]
    int* ib = reinterpret_cast<int*>(dbl + 1024);
    int* perm = ib;               // [128] sorted position -> original column
    int* kcol = ib + 128;         // [128] compact index -> original column
    int* kcnt = ib + 256;         // [MG_MAXR] non-deflated count per range
    int* rotl = ib + 384;         // [128][2] Givens column pairs
    float* rotcs = reinterpret_cast<float*>(ib + 640);   // [128][2]
    int* nrotl = ib + 896;        // [1]
    const int tid = threadIdx.x;
    const int sig = blockIdx.x;
    if (a.skip && a.skip[sig]) return;
    const float* Zg = a.Zin + (size_t)sig * d * ldr;        // global pitch = ldr (see k_rot)
    float* Zo = a.Zout + (size_t)sig * d * ldr;
    float* lam = a.lam + (size_t)sig * d;

    for (int idx = tid; idx < d * ldr; idx += MG_THREADS) Zs[idx] = Zg[idx];
    for (int idx = tid; idx < d * ldr; idx += MG_THREADS) Ws[idx] = 0.f;
    if (tid == 0) *nrotl = 0;
    __syncthreads();
    // range of every index
    int myr = -1;
    if (tid < d)
        for (int r = 0; r < a.nr; ++r)
            if (tid >= a.ra[r] && tid < a.rb[r]) myr = r;
    // ---- z = Q^T v: row p-1 plus row p of Q (Zs[c][r] = Q[r][c])
    double zi = 0.0, di = 0.0;
    if (myr >= 0) {
        const int p = a.rp[myr];
        zi = (double)Zs[tid * ldr + p - 1] + (double)Zs[tid * ldr + p];
        di = (double)lam[tid];
        kmu[tid] = zi;                 // scratch: unsorted z
    }
    __syncthreads();
    if (tid < a.nr) {
        double n2 = 0.0;
        for (int i = a.ra[tid]; i < a.rb[tid]; ++i) n2 += kmu[i] * kmu[i];
        const double rho = a.rho[(size_t)sig * DC_MAXTEAR + a.rt[tid]] * n2;
        rsc[tid] = fabs(rho);
        rsc[4 + tid] = rho < 0.0 ? -1.0 : 1.0;
        rsc[8 + tid] = n2 > 0.0 ? 1.0 / sqrt(n2) : 0.0;
    }
    __syncthreads();
    if (myr >= 0) {
        zi *= rsc[8 + myr];
        di *= rsc[4 + myr];            // rho < 0: solve for -T
        korg[tid] = di;                // scratch: unsorted signed poles
        kmu[tid] = zi;
    }
    __syncthreads();
    // ---- rank sort inside each range
    if (myr >= 0) {
        int rank = a.ra[myr];
        for (int j = a.ra[myr]; j < a.rb[myr]; ++j) {
            const double dj = korg[j];
            rank += (dj < di || (dj == di && j < tid)) ? 1 : 0;
        }
        sd[rank] = di;
        sz[rank] = zi;
        perm[rank] = tid;
    }
    __syncthreads();
    // ---- deflation scan, one thread per range (serial; rotations are rare)
    if (tid < a.nr) {
        const int ra = a.ra[tid], rb = a.rb[tid];
        int k = 0, prev = -1;
        const bool noop = !(rsc[tid] > 0.0);
        for (int t = ra; t < rb; ++t) {
            if (noop || fabs(sz[t]) <= 1e-9) { sz[t] = 0.0; continue; }
            if (prev >= 0 && sd[t] == sd[prev]) {
                const double r = hypot(sz[prev], sz[t]);
                const double c = sz[t] / r, s = sz[prev] / r;
                const int slot = atomicAdd(nrotl, 1);
                rotl[2 * slot] = perm[prev];
                rotl[2 * slot + 1] = perm[t];
                rotcs[2 * slot] = (float)c;
                rotcs[2 * slot + 1] = (float)s;
                sz[t] = r;
                sz[prev] = 0.0;
            }
            prev = t;
        }
        for (int t = ra; t < rb; ++t)
            if (sz[t] != 0.0) {
                ksd[ra + k] = sd[t];
                ksz[ra + k] = sz[t];
                kcol[ra + k] = perm[t];
                ++k;
            }
        kcnt[tid] = k;
    }
    __syncthreads();
    // ---- apply the (rare) Givens rotations to the columns, in order
    {
        const int nrot_ = *nrotl;
        for (int q = 0; q < nrot_; ++q) {
            const int cp = rotl[2 * q], ct = rotl[2 * q + 1];
            const float c = rotcs[2 * q], s = rotcs[2 * q + 1];
            for (int r = tid; r < d; r += MG_THREADS) {
                const float xp = Zs[cp * ldr + r], xt = Zs[ct * ldr + r];
                Zs[cp * ldr + r] = c * xp - s * xt;
                Zs[ct * ldr + r] = s * xp + c * xt;
            }
            __syncthreads();
        }
    }
    // ---- secular roots: a pair of lanes per root (slot = tid/2 = ra + u)
    if (myr >= 0) {
        const int ra = a.ra[myr], u = tid - ra;
        if (u < kcnt[myr]) kwin[tid] = ksz[ra + u] * ksz[ra + u];       // z^2 (kwin doubles as scratch)
    }
    __syncthreads();
    {
        const int slot = tid >> 1, par = tid & 1;
        int sr = -1;
        if (slot < d)
            for (int r = 0; r < a.nr; ++r)
                if (slot >= a.ra[r] && slot < a.rb[r]) sr = r;
        const int ra = sr >= 0 ? a.ra[sr] : 0, u = slot - ra, k = sr >= 0 ? kcnt[sr] : 0;
        const bool work = sr >= 0 && u < k;                  // uniform within the lane pair
        const double rho = work ? rsc[sr] : 1.0;
        const double* pd = ksd + ra;
        const double* pz2 = kwin + ra;
        double org = 0.0, lo = 0.0, hi = 1.0, mu = 0.5, newlam = 0.0, winv = 0.0;
        const int kk = work ? k : 0;
        if (work) {
            if (u < k - 1) {
                const double half = 0.5 * (pd[u + 1] - pd[u]);
                org = pd[u];
                lo = 0.0; hi = half;
            } else {
                double s2 = 0.0;
                for (int i = 0; i < k; ++i) s2 += pz2[i];
                org = pd[u]; lo = 0.0; hi = rho * s2 * 1.000001 + 1e-300;
            }
        }
        // which half of the interval holds the root decides the origin (nearer pole)
        double dl = 0.0, dr = 0.0;                 // neighbouring poles in shifted coordinates
        const bool lastroot = work && u == k - 1;
        {
            const SecVal v = secular_eval<float>(pd, pz2, kk, u, org, hi, par);
            const double fm = 1.0 + rho * (v.psi + v.phi);
            if (work && !lastroot) {
                const double gap = 2.0 * hi;
                if (fm < 0.0) { org = pd[u + 1]; lo = -hi; hi = 0.0; dl = -gap; dr = 0.0; }
                else { dl = 0.0; dr = gap; }
            }
        }
        mu = 0.5 * (lo + hi);
        // bulk of the iteration in fp32 arithmetic (differences still formed in fp64), then fp64 polish
        bool conv = !work;
        for (int it = 0; it < 30; ++it) {
            if (__all_sync(0xffffffffu, conv)) break;
            const SecVal v = secular_eval<float>(pd, pz2, kk, u, org, mu, par);
            if (!conv) {
                const double f = 1.0 + rho * (v.psi + v.phi);
                const double fa = 1.0 + rho * (fabs(v.psi) + fabs(v.phi));
                if (f > 0.0) hi = mu; else lo = mu;
                double nx = mu + secular_step(v, rho, f, mu, dl, dr, lastroot);
                if (!(nx > lo && nx < hi)) nx = 0.5 * (lo + hi);
                if (fabs(f) <= 4e-6 * fa || fabs(nx - mu) <= 2e-6 * fabs(nx) || nx == mu) conv = true;
                mu = nx;
            }
        }
        conv = !work;
        for (int it = 0; it < 5; ++it) {
            if (__all_sync(0xffffffffu, conv)) break;
            const SecVal v = secular_eval<double>(pd, pz2, kk, u, org, mu, par);
            if (!conv) {
                const double f = 1.0 + rho * (v.psi + v.phi);
                const double fa = 1.0 + rho * (fabs(v.psi) + fabs(v.phi));
                if (f > 0.0) hi = mu; else lo = mu;
                if (fabs(f) <= 2e-14 * fa) conv = true;
                else {
                    double nx = mu + secular_step(v, rho, f, mu, dl, dr, lastroot);
                    if (!(nx > lo && nx < hi)) nx = 0.5 * (lo + hi);
                    if (fabs(nx - mu) <= 1e-13 * fabs(nx) || (hi - lo) <= 4e-16 * fmax(fabs(lo), fabs(hi))) conv = true;
                    mu = nx;
                }
            }
        }
        // norm of the secular eigenvector (fp32 arithmetic on fp64-formed differences)
        {
            double n2 = 0.0;       // fp64: entries next to a pole reach 1/mu, far outside the fp32 range
            for (int i = par; i < kk; i += 2) {
                const double w = ksz[ra + i] * sec_rcp((pd[i] - org) - mu);
                n2 = fma(w, w, n2);
            }
            n2 += __shfl_xor_sync(0xffffffffu, n2, 1);
            winv = work ? 1.0 / sqrt(n2) : 0.0;
            newlam = org + mu;
        }
        __syncthreads();               // all reads of kwin (z^2) and sd are done
        if (work && par == 0) {
            kmu[slot] = mu;
            korg[slot] = org;
            sd[slot] = winv;           // sd is dead (compact copies are used from here on)
            lam[kcol[slot]] = (float)(rsc[4 + sr] * newlam);
        }
    }
    __syncthreads();
    // ---- W (block diagonal): Ws[ra+i][ra+t] = z_i / (d_i - lambda_t) / ||.||
    for (int r = 0; r < a.nr; ++r) {
        const int ra = a.ra[r], k = kcnt[r];
        for (int idx = tid; idx < k * k; idx += MG_THREADS) {
            const int i = idx / k, t = idx % k;
            const double den = (ksd[ra + i] - korg[ra + t]) - kmu[ra + t];
            Ws[(ra + i) * ldr + ra + t] = (float)(ksz[ra + i] * sec_rcp(den) * sd[ra + t]);
        }
    }
    __syncthreads();
    // ---- every column is copied first (deflated pairs keep theirs; rows outside a block are zero), then the
    //      non-deflated ones are overwritten by Q * W (4 rows x 4 roots per thread)
    for (int idx = tid; idx < d * ldr; idx += MG_THREADS) Zo[idx] = Zs[idx];
    __syncthreads();
    for (int r = 0; r < a.nr; ++r) {
        const int ra = a.ra[r], rb = a.rb[r], k = kcnt[r];
        if (k == 0) continue;
        const int row0 = ra & ~3;
        const int nrt = (rb - row0 + 3) / 4, nct = (k + 3) / 4;
        for (int tile = tid; tile < nrt * nct; tile += MG_THREADS) {
            const int rt = tile % nrt, ct = tile / nrt;
            const int r0 = row0 + 4 * rt, t0 = 4 * ct;
            float acc[4][4];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
            for (int u = 0; u < k; ++u) {
                const float4 q = *reinterpret_cast<const float4*>(Zs + (size_t)kcol[ra + u] * ldr + r0);
                const float* wrow = Ws + (size_t)(ra + u) * ldr + ra + t0;
                const float w0 = wrow[0], w1 = wrow[1], w2 = wrow[2], w3 = wrow[3];
                acc[0][0] = fmaf(q.x, w0, acc[0][0]); acc[0][1] = fmaf(q.x, w1, acc[0][1]);
                acc[0][2] = fmaf(q.x, w2, acc[0][2]); acc[0][3] = fmaf(q.x, w3, acc[0][3]);
                acc[1][0] = fmaf(q.y, w0, acc[1][0]); acc[1][1] = fmaf(q.y, w1, acc[1][1]);
                acc[1][2] = fmaf(q.y, w2, acc[1][2]); acc[1][3] = fmaf(q.y, w3, acc[1][3]);
                acc[2][0] = fmaf(q.z, w0, acc[2][0]); acc[2][1] = fmaf(q.z, w1, acc[2][1]);
                acc[2][2] = fmaf(q.z, w2, acc[2][2]); acc[2][3] = fmaf(q.z, w3, acc[2][3]);
                acc[3][0] = fmaf(q.w, w0, acc[3][0]); acc[3][1] = fmaf(q.w, w1, acc[3][1]);
                acc[3][2] = fmaf(q.w, w2, acc[3][2]); acc[3][3] = fmaf(q.w, w3, acc[3][3]);
            }
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                if (t0 + y >= k) continue;
                float* oc = Zo + (size_t)kcol[ra + t0 + y] * ldr;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int rr = r0 + x;
                    if (rr >= ra && rr < rb) oc[rr] = acc[x][y];
                }
            }
        }
    }
}

// G = U diag(l') U^H on 4x4 register tiles of the lower triangle, written packed to GV; returns this thread's share
// of ||G - C||_F^2 (C = [[diag(h), phi],[phi^H, c1z]]) when with_c.  U is column-major [k][ldu] in shared memory.
template <int NT, bool PACKED>
__device__ __forceinline__ float rebuild_lower(const float2* __restrict__ U, int ldu, const float* __restrict__ lamp,
                                               int d, int n, float2* __restrict__ GV, const float* __restrict__ hs,
                                               const float2* __restrict__ phis, float c1z, bool with_c) {
    const int tid = threadIdx.x;
    const int nt1 = (d + 3) / 4;
    const int ntiles = nt1 * (nt1 + 1) / 2;
    float rsq = 0.f;
    for (int t = tid; t < ntiles; t += NT) {
        int ti = (int)((sqrtf(8.f * (float)t + 1.f) - 1.f) * 0.5f);
        while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
        while (ti * (ti + 1) / 2 > t) --ti;
        const int tj = t - ti * (ti + 1) / 2;
        const float4* Ua = reinterpret_cast<const float4*>(U + 4 * ti);
        const float4* Ub = reinterpret_cast<const float4*>(U + 4 * tj);
        const int ld4 = ldu / 2;   // float4 stride per column
        float2 acc[4][4];
        if constexpr (PACKED) {
            // acc[x][y] = sum_k (l' u_x) conj(u_y) with packed FMAs and no operand shuffling: with a = l' u_x as the
            // (re, im) pair,  P += b.re * a  and  Q += b.im * a  (b = u_y as broadcast scalars), and at the end
            // acc = (P.re + Q.im, P.im - Q.re).
            f32x2 Pp[4][4], Qp[4][4];
    #pragma unroll
            for (int x = 0; x < 4; ++x)
    #pragma unroll
                for (int yv = 0; yv < 4; ++yv) { Pp[x][yv] = pk2(0.f, 0.f); Qp[x][yv] = pk2(0.f, 0.f); }
            for (int k = 0; k < d; ++k) {
                const f32x2 lp = bc2(lamp[k]);
                const float4 a01 = Ua[(size_t)k * ld4], a23 = Ua[(size_t)k * ld4 + 1];
                const float4 b01 = Ub[(size_t)k * ld4], b23 = Ub[(size_t)k * ld4 + 1];
                const f32x2 av[4] = {mul2(lp, pk2(a01.x, a01.y)), mul2(lp, pk2(a01.z, a01.w)), mul2(lp, pk2(a23.x, a23.y)),
                                     mul2(lp, pk2(a23.z, a23.w))};
                const float bre[4] = {b01.x, b01.z, b23.x, b23.z}, bim[4] = {b01.y, b01.w, b23.y, b23.w};
    #pragma unroll
                for (int x = 0; x < 4; ++x)
    #pragma unroll
                    for (int yv = 0; yv < 4; ++yv) {
                        Pp[x][yv] = fma2(bc2(bre[yv]), av[x], Pp[x][yv]);
                        Qp[x][yv] = fma2(bc2(bim[yv]), av[x], Qp[x][yv]);
                    }
            }
    #pragma unroll
            for (int x = 0; x < 4; ++x)
    #pragma unroll
                for (int yv = 0; yv < 4; ++yv) {
                    const float2 pp = upk2(Pp[x][yv]), qq = upk2(Qp[x][yv]);
                    acc[x][yv] = make_float2(pp.x + qq.y, pp.y - qq.x);
                }
        } else {
            // scalar form (k_arrow: its two-CTA register budget has no room for the second accumulator set)
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int yv = 0; yv < 4; ++yv) acc[x][yv] = make_float2(0.f, 0.f);
            for (int k = 0; k < d; ++k) {
                const float lp = lamp[k];
                const float4 a01 = Ua[(size_t)k * ld4], a23 = Ua[(size_t)k * ld4 + 1];
                const float4 b01 = Ub[(size_t)k * ld4], b23 = Ub[(size_t)k * ld4 + 1];
                const float2 av[4] = {make_float2(lp * a01.x, lp * a01.y), make_float2(lp * a01.z, lp * a01.w),
                                      make_float2(lp * a23.x, lp * a23.y), make_float2(lp * a23.z, lp * a23.w)};
                const float2 bv[4] = {make_float2(b01.x, b01.y), make_float2(b01.z, b01.w), make_float2(b23.x, b23.y),
                                      make_float2(b23.z, b23.w)};
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int yv = 0; yv < 4; ++yv) {
                        acc[x][yv].x = fmaf(av[x].x, bv[yv].x, acc[x][yv].x);
                        acc[x][yv].x = fmaf(av[x].y, bv[yv].y, acc[x][yv].x);
                        acc[x][yv].y = fmaf(av[x].y, bv[yv].x, acc[x][yv].y);
                        acc[x][yv].y = fmaf(-av[x].x, bv[yv].y, acc[x][yv].y);
                    }
            }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const int i = 4 * ti + x;
            if (i >= d) continue;
#pragma unroll
            for (int yv = 0; yv < 4; ++yv) {
                const int j = 4 * tj + yv;
                if (j > i) continue;
                float2 g = acc[x][yv];
                if (i == j) g.y = 0.f;
                GV[pk(i, j)] = g;
                if (with_c) {
                    float2 c = make_float2(0.f, 0.f);
                    if (i == j) c.x = (i < n) ? hs[i] : c1z;
                    else if (i == n) c = cconj(phis[j]);
                    const float rx = g.x - c.x, ry = g.y - c.y;
                    rsq += (i == j ? 1.f : 2.f) * (rx * rx + ry * ry);
                }
            }
        }
    }
    return rsq;
}

// =====================================================================================
// k_tail: U = Q_H Z, l' = f(l), G = U diag(l') U^H (lower triangle), r = ||G - C||_F.
// Thread mapping for the back-transformation: (column pair cp = tid>>3, row split s = tid&7),
// rows s+8j kept in registers for all d-1 reflectors -> no block barrier inside that phase.
// =====================================================================================
struct TailArgs {
    const float* Zr;        // [B][d][ldz]  Z transposed: Zr[c][r], row pitch ldz = 4*ceil(d/4)
    float2* GV;             // [B][npk]: reflectors on entry, G packed lower on exit
    const float2* tau;      // [B][d]
    const float* lam;       // [B][d]
    const float2* phi_cur;  // [B][n]
    const float* h_cur;     // [B][n]
    const float* Pk;
    float* r_out;           // [B]
    float2* U_out;          // optional [B][d][d] row-major eigenvectors (debug taps), may be null
    float* lamp_out;        // optional [B][d]
    int B, n, d, ldu, with_c;  // with_c=0: plain f(A) for the debug entry (no residual)
    const int* skip;
};
__host__ __device__ inline size_t tail_smem_bytes(int d, int ldu) {
    const size_t nv = (size_t)d * (d - 1) / 2;
    return ((size_t)d * ldu + nv + 2 + 128 /*tau*/ + 128 /*phi*/) * sizeof(float2) + (128 + 128 + 96) * sizeof(float);
}

template <int NR, int NT, int RL>
__global__ void __launch_bounds__(NT, 1) k_tail(TailArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, d = a.d, ldu = a.ldu;
    const int nv = d * (d - 1) / 2;
    float2* U = reinterpret_cast<float2*>(smem_raw);          // [d][ldu] column-major (also Z staging)
    float2* Vs = U + (size_t)d * ldu;                         // reflectors
    float2* taus = Vs + ((nv + 1) & ~1);
    float2* phis = taus + 128;
    float* lamp = reinterpret_cast<float*>(phis + 128);       // [128] mapped eigenvalues
    float* hs = lamp + 128;
    float* red = hs + 128;
    const int tid = threadIdx.x;
    const int sig = blockIdx.x;
    if (a.skip && a.skip[sig]) return;
    const int npk = d * (d + 1) / 2;
    float2* GV = a.GV + (size_t)sig * npk;
    const float* __restrict__ P = a.Pk;

    // ---- phase A: stage reflectors, Z, per-signal vectors
    for (int idx = tid; idx < nv; idx += NT) Vs[idx] = GV[idx];
    float* Zs = reinterpret_cast<float*>(U);
    const int ldz = 4 * ((d + 3) / 4);                        // row pitch of the global Z^T (k_rot)
    const float* Zg = a.Zr + (size_t)sig * d * ldz;
    for (int idx = tid; idx < d * ldz; idx += NT) Zs[idx] = Zg[idx];
    for (int i = tid; i < d; i += NT) {
        taus[i] = a.tau[(size_t)sig * d + i];
        const float l = a.lam[(size_t)sig * d + i];
        const float lp = a.with_c >= 0 ? eig_map(P, l) : l;
        lamp[i] = lp;
        if (a.lamp_out) a.lamp_out[(size_t)sig * d + i] = lp;
    }
    if (a.with_c > 0) {
        for (int j = tid; j < n; j += NT) {
            phis[j] = a.phi_cur[(size_t)sig * n + j];
            hs[j] = a.h_cur[(size_t)sig * n + j];
        }
    }
    __syncthreads();

    // ---- phase B: back-transformation in registers
    const int s8 = tid & (RL - 1), cp = tid / RL;        // RL lanes share a column pair, rows s8 + RL*j
    const int c0 = 2 * cp, c1 = 2 * cp + 1;
    // The two columns of a thread are kept as packed pairs across the columns: Mx[j] = (Re M0[j], Re M1[j]),
    // My[j] = (Im M0[j], Im M1[j]); with the reflector entry as a broadcast operand every complex multiply-add of the
    // pair of columns is two FFMA2 (same operation order and rounding as the scalar form).
    f32x2 Mx[NR], My[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        const int r = s8 + RL * j;
        Mx[j] = pk2((r < d && c0 < d) ? Zs[c0 * ldz + r] : 0.f, (r < d && c1 < d) ? Zs[c1 * ldz + r] : 0.f);
        My[j] = pk2(0.f, 0.f);
    }
    __syncthreads();   // Z staging area is dead from here on (becomes U)
    // Reflector k touches rows r > k.  Rows are held as r = s8 + 8j, so for the 8 reflectors with
    // (k+1)/8 == jm only the row slots j >= jm take part: slot jm under a per-lane predicate, slots > jm
    // unconditionally (compile-time bounds after unrolling jm -> no per-slot branches in the hot loop).
#pragma unroll
    for (int jm = NR - 1; jm >= 0; --jm) {
        const int khi = min(RL * jm + RL - 2, d - 2), klo = max(RL * jm - 1, 0);
        for (int k = khi; k >= klo; --k) {
            const float2 tk = taus[k];
            if (tk.x == 0.f && tk.y == 0.f) continue;
            const int vbase = voff(k, d) - (k + 1);           // Vs[vbase + r] valid for r >= k+1
            float2 vv[NR];
            f32x2 Dx = pk2(0.f, 0.f), Dy = pk2(0.f, 0.f);     // conj(v) . M for both columns
#pragma unroll
            for (int j = jm; j < NR; ++j) {
                const int r = s8 + RL * j;
                const bool on = (j == jm ? r > k : true) && (r < d);
                float2 v = make_float2(0.f, 0.f);
                if (on) v = Vs[vbase + r];
                vv[j] = v;
                Dx = fma2(bc2(v.x), Mx[j], Dx); Dx = fma2(bc2(v.y), My[j], Dx);
                Dy = fma2(bc2(v.x), My[j], Dy); Dy = fma2(bc2(-v.y), Mx[j], Dy);
            }
            float2 dx = upk2(Dx), dy = upk2(Dy);              // dx = (d0.x, d1.x), dy = (d0.y, d1.y)
#pragma unroll
            for (int o = 1; o < RL; o <<= 1) {
                dx.x += __shfl_xor_sync(0xffffffffu, dx.x, o);
                dy.x += __shfl_xor_sync(0xffffffffu, dy.x, o);
                dx.y += __shfl_xor_sync(0xffffffffu, dx.y, o);
                dy.y += __shfl_xor_sync(0xffffffffu, dy.y, o);
            }
            const float2 t0 = cmul(tk, make_float2(dx.x, dy.x)), t1 = cmul(tk, make_float2(dx.y, dy.y));
            const f32x2 nTx = pk2(-t0.x, -t1.x), Ty = pk2(t0.y, t1.y), nTy = pk2(-t0.y, -t1.y);
#pragma unroll
            for (int j = jm; j < NR; ++j) {
                const float2 v = vv[j];
                Mx[j] = fma2(bc2(v.x), nTx, Mx[j]); Mx[j] = fma2(bc2(v.y), Ty, Mx[j]);
                My[j] = fma2(bc2(v.y), nTx, My[j]); My[j] = fma2(bc2(v.x), nTy, My[j]);
            }
        }
    }
    float2 M0[NR], M1[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        const float2 mx = upk2(Mx[j]), my = upk2(My[j]);
        M0[j] = make_float2(mx.x, my.x);
        M1[j] = make_float2(mx.y, my.y);
    }
    // ---- phase C: U to shared memory (column-major, zero padded rows) and optional global tap
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        const int r = s8 + RL * j;
        if (r < ldu) {
            if (c0 < d) U[(size_t)c0 * ldu + r] = r < d ? M0[j] : make_float2(0.f, 0.f);
            if (c1 < d) U[(size_t)c1 * ldu + r] = r < d ? M1[j] : make_float2(0.f, 0.f);
        }
        if (a.U_out && r < d) {
            float2* Uo = a.U_out + (size_t)sig * d * d + (size_t)r * d;
            if (c0 < d) Uo[c0] = M0[j];
            if (c1 < d) Uo[c1] = M1[j];
        }
    }
    __syncthreads();

    // ---- phase D: G = U diag(l') U^H (lower triangle) and the residual norm
    const float rsq = rebuild_lower<NT, true>(U, ldu, lamp, d, n, GV, hs, phis, P ? P[P_C1Z] : 0.f, a.with_c > 0);
    if (a.with_c > 0) {
        float v[1] = {rsq};
        block_sum<1>(v, red);
        if (tid == 0) a.r_out[sig] = sqrtf(v[0]);
    }
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc));
}

// =====================================================================================
// k_tail_p: the production form of k_tail (layers >= 1 of admmnet_forward).  Persistent: one CTA per SM walks the
// signals sig = blockIdx.x, blockIdx.x + gridDim.x, ...; while a signal is being back-transformed and rebuilt, the
// NEXT signal's Z^T, reflectors, tau, eigenvalues, phi and h stream into a second set of shared-memory buffers with
// cp.async, so the HBM/L2 latency of the staging (11 % of k_tail's time at one CTA per SM) is hidden.
// Shared memory: U [d][ldu] c64 | V[2] reflectors | Zn (next Z^T, consumed into registers at the top of an
// iteration and refilled right away) | small per-signal vectors [2]  = 205 KB at d = 101.
// =====================================================================================
__host__ __device__ inline size_t tailp_smem_bytes(int d, int ldu) {
    const size_t nv2 = ((size_t)d * (d - 1) / 2 + 1) & ~(size_t)1;
    const size_t zf = (size_t)d * (4 * ((d + 3) / 4));
    return ((size_t)d * ldu + 2 * nv2 + 2 * 128 /*tau*/ + 2 * 128 /*phi*/) * sizeof(float2) +
           (zf + 2 * 128 /*lam*/ + 2 * 128 /*h*/ + 128 /*lamp*/ + 96) * sizeof(float);
}
template <int NR, int NT, int RL>
__global__ void __launch_bounds__(NT, 1) k_tail_p(TailArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, d = a.d, ldu = a.ldu;
    const int nv = d * (d - 1) / 2;
    const int nv2 = (nv + 1) & ~1;
    const int ldz = 4 * ((d + 3) / 4);                        // row pitch of the global Z^T (k_rot)
    const int zf = d * ldz;
    float2* U = reinterpret_cast<float2*>(smem_raw);          // [d][ldu] column-major
    float2* Vb = U + (size_t)d * ldu;                         // [2][nv2] reflectors
    float2* taub = Vb + 2 * (size_t)nv2;                      // [2][128]
    float2* phib = taub + 2 * 128;                            // [2][128]
    float* Zn = reinterpret_cast<float*>(phib + 2 * 128);     // [d*d] Z^T of the coming signal
    float* lamb = Zn + zf;                                    // [2][128] eigenvalues
    float* hb = lamb + 2 * 128;                               // [2][128]
    float* lamp = hb + 2 * 128;                               // [128] mapped eigenvalues
    float* red = lamp + 128;
    const int tid = threadIdx.x;
    const int npk = d * (d + 1) / 2;
    const float* __restrict__ P = a.Pk;
    const float c1z = P[P_C1Z];

    auto prefetch = [&](int sig, int buf) {                   // everything signal `sig` needs, one cp.async group
        const float2* gv = a.GV + (size_t)sig * npk;
        float2* vd = Vb + (size_t)buf * nv2;
        for (int idx = tid; idx < nv; idx += NT) cp_async8(vd + idx, gv + idx);
        const float* zg = a.Zr + (size_t)sig * zf;
        for (int idx = tid; idx < zf / 4; idx += NT) cp_async16(Zn + 4 * idx, zg + 4 * idx);
        for (int i = tid; i < d; i += NT) {
            cp_async8(taub + buf * 128 + i, a.tau + (size_t)sig * d + i);
            cp_async4(lamb + buf * 128 + i, a.lam + (size_t)sig * d + i);
        }
        for (int j = tid; j < n; j += NT) {
            cp_async8(phib + buf * 128 + j, a.phi_cur + (size_t)sig * n + j);
            cp_async4(hb + buf * 128 + j, a.h_cur + (size_t)sig * n + j);
        }
        cp_async_commit();
    };

    int buf = 0;
    if ((int)blockIdx.x < a.B) prefetch(blockIdx.x, 0);
    for (int sig = blockIdx.x; sig < a.B; sig += gridDim.x, buf ^= 1) {
        float2* GV = a.GV + (size_t)sig * npk;
        const float2* Vs = Vb + (size_t)buf * nv2;
        const float2* taus = taub + buf * 128;
        const float2* phis = phib + buf * 128;
        const float* hs = hb + buf * 128;
        cp_async_wait<0>();
        __syncthreads();              // this signal's data has landed; the previous signal's rebuild is finished
        // ---- Z into registers (packed pairs across the thread's two columns)
        const int s8 = tid & (RL - 1), cp = tid / RL;
        const int c0 = 2 * cp, c1 = 2 * cp + 1;
        f32x2 Mx[NR], My[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const int r = s8 + RL * j;
            Mx[j] = pk2((r < d && c0 < d) ? Zn[c0 * ldz + r] : 0.f, (r < d && c1 < d) ? Zn[c1 * ldz + r] : 0.f);
            My[j] = pk2(0.f, 0.f);
        }
        if (tid < d) lamp[tid] = eig_map(P, lamb[buf * 128 + tid]);
        __syncthreads();              // Zn is free again
        if (sig + (int)gridDim.x < a.B) prefetch(sig + gridDim.x, buf ^ 1);
        // ---- back-transformation (see k_tail)
#pragma unroll
        for (int jm = NR - 1; jm >= 0; --jm) {
            const int khi = min(RL * jm + RL - 2, d - 2), klo = max(RL * jm - 1, 0);
            for (int k = khi; k >= klo; --k) {
                const float2 tk = taus[k];
                if (tk.x == 0.f && tk.y == 0.f) continue;
                const int vbase = voff(k, d) - (k + 1);
                float2 vv[NR];
                f32x2 Dx = pk2(0.f, 0.f), Dy = pk2(0.f, 0.f);
#pragma unroll
                for (int j = jm; j < NR; ++j) {
                    const int r = s8 + RL * j;
                    const bool on = (j == jm ? r > k : true) && (r < d);
                    float2 v = make_float2(0.f, 0.f);
                    if (on) v = Vs[vbase + r];
                    vv[j] = v;
                    Dx = fma2(bc2(v.x), Mx[j], Dx); Dx = fma2(bc2(v.y), My[j], Dx);
                    Dy = fma2(bc2(v.x), My[j], Dy); Dy = fma2(bc2(-v.y), Mx[j], Dy);
                }
                float2 dx = upk2(Dx), dy = upk2(Dy);
#pragma unroll
                for (int o = 1; o < RL; o <<= 1) {
                    dx.x += __shfl_xor_sync(0xffffffffu, dx.x, o);
                    dy.x += __shfl_xor_sync(0xffffffffu, dy.x, o);
                    dx.y += __shfl_xor_sync(0xffffffffu, dx.y, o);
                    dy.y += __shfl_xor_sync(0xffffffffu, dy.y, o);
                }
                const float2 t0 = cmul(tk, make_float2(dx.x, dy.x)), t1 = cmul(tk, make_float2(dx.y, dy.y));
                const f32x2 nTx = pk2(-t0.x, -t1.x), Ty = pk2(t0.y, t1.y), nTy = pk2(-t0.y, -t1.y);
#pragma unroll
                for (int j = jm; j < NR; ++j) {
                    const float2 v = vv[j];
                    Mx[j] = fma2(bc2(v.x), nTx, Mx[j]); Mx[j] = fma2(bc2(v.y), Ty, Mx[j]);
                    My[j] = fma2(bc2(v.y), nTx, My[j]); My[j] = fma2(bc2(v.x), nTy, My[j]);
                }
            }
        }
        // ---- U to shared memory (column-major, zero padded rows)
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const int r = s8 + RL * j;
            const float2 mx = upk2(Mx[j]), my = upk2(My[j]);
            if (r < ldu) {
                if (c0 < d) U[(size_t)c0 * ldu + r] = r < d ? make_float2(mx.x, my.x) : make_float2(0.f, 0.f);
                if (c1 < d) U[(size_t)c1 * ldu + r] = r < d ? make_float2(mx.y, my.y) : make_float2(0.f, 0.f);
            }
        }
        __syncthreads();
        // ---- G = U diag(l') U^H (lower triangle) and the residual norm
        const float rsq = rebuild_lower<NT, true>(U, ldu, lamp, d, n, GV, hs, phis, c1z, true);
        float v[1] = {rsq};
        block_sum<1>(v, red);
        if (tid == 0) a.r_out[sig] = sqrtf(v[0]);
    }
}

// =====================================================================================
// k_mean: deterministic sum of r over the chunk (double), one CTA.
//   out_sum[0] = sum, and if mean_out != null, mean_out[0] = float(sum / count_total).
// =====================================================================================
__global__ void __launch_bounds__(1024) k_rsum(const float* __restrict__ r, int B, double* __restrict__ out_sum) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 1024) acc += (double)r[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = red[threadIdx.x];
        v = warp_sum(v);
        if (threadIdx.x == 0) out_sum[0] = v;
    }
}
__global__ void k_mean_from_sum(const double* __restrict__ sum, double count, float* __restrict__ mean_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) mean_out[0] = (float)(sum[0] / count);
}

// =====================================================================================
// k_final_phi: phi of the last layer (admm_net.py:764) from the state left by layer K-2.
// One warp per signal.
// =====================================================================================
struct FinalArgs {
    const float2* y;
    const float2* b;
    const float2* Zp;
    const float2* GV;
    const float2* phi_prev;
    const float* r_prev;
    const float* mean_prev;
    const float* Pk;     // last layer (rho_phi)
    const float* Pkm1;   // layer K-2 (alpha)
    float2* phi_out;
    int B, n, d, first;
};
__global__ void __launch_bounds__(256) k_final_phi(FinalArgs a) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= a.B) return;
    const int n = a.n, d = a.d, npk = d * (d + 1) / 2;
    const float rho_phi = a.Pk[P_RHO_PHI];
    float alpha = 0.f;
    if (!a.first) alpha = z_alpha(a.Pkm1, a.r_prev[w], *a.mean_prev);
    const float2* Zrow = a.Zp + (size_t)w * npk + pk(n, 0);
    const float2* Grow = a.GV + (size_t)w * npk + pk(n, 0);
    for (int j = lane; j < n; j += 32) {
        float2 g = make_float2(0.f, 0.f), z = make_float2(0.f, 0.f);
        if (!a.first) {
            const float2 gr = Grow[j], zr = Zrow[j];
            const float2 c = cconj(a.phi_prev[(size_t)w * n + j]);
            const float2 zn = make_float2(zr.x + alpha * (gr.x - c.x), zr.y + alpha * (gr.y - c.y));
            g = cconj(gr);
            z = cconj(zn);
        }
        const float2 bj = a.b[(size_t)w * n + j], yj = a.y[(size_t)w * n + j];
        const float ab = hypotf(bj.x, bj.y);
        const float bsq = ab * ab + ADMM_EPS;
        const float wgt = bsq / (1.f + rho_phi * bsq);
        const float2 yob = cdiv(yj, make_float2(bj.x + ADMM_EPS, bj.y));
        a.phi_out[(size_t)w * n + j] =
            make_float2(wgt * (yob.x + rho_phi * g.x + z.x), wgt * (yob.y + rho_phi * g.y + z.y));
    }
}

}  // namespace admmnet
