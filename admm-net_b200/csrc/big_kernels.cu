// k_big_layer / k_big_eigh: the size-agnostic form of one unrolled layer for matrix orders d = n+1 > 128
// (BASELINE.json configs[3]: n = 144, 196, 256).  The reference forward is size-agnostic (admm_net.py:726-764); the
// production kernels (k_head .. k_tail_tc) keep one matrix per CTA in shared memory / tensor memory and stop at d = 128.
//
// Same layer mathematics as k_head + k_tail (dual update of the previous layer, admm_net.py:388-412; phi-update 79-105;
// H-update 134-194; block matrix 262-290; eigh 292-308; eigenvalue map 310-334; rebuild 336-354; residual norm 454),
// same per-signal state (packed Z, packed G, phi, h, r).  The eigen-decomposition is a cyclic two-sided Jacobi method on
// the full Hermitian matrix, which lives with its eigenvector matrix (and an untouched copy for the final Rayleigh
// quotients) in a per-CTA global-memory scratch (3 d^2 complex, L2 resident: 1.6 MB at d = 257):  rounds of floor((d+1)/2) disjoint plane rotations (round-robin tournament), each round
//   (1) rotation parameters from (a_pp, a_qq, a_pq) -> shared memory;  columns p, q of A and of V rotated  (A <- A J)
//   (2) rows p, q of A rotated (A <- J^H A), a_pq := 0
// until a sweep meets ||off(A)||_F <= 2e-7 ||A||_F (quadratic convergence; 6-9 sweeps in fp32).  Jacobi needs no
// tridiagonal form, no shift strategy and is backward stable entry-wise, which is what a correctness path wants; it costs
// ~40x the flops of the Householder/QL pipeline, so it is NOT the benchmarked path.
// Persistent: CTA c handles signals c, c + grid, ...
#include "common.cuh"

namespace admmnet {

constexpr int BIG_NT = 512;
constexpr int BIG_NMAX = 256;        // n <= 256 (cfg 4's largest, 16 x 16)
constexpr int BIG_MAX_SWEEPS = 16;

struct BigArgs {
    const float2* y;
    const float2* b;
    const float* sigma;
    float2* Zp;          // [B][npk]
    float2* GV;          // [B][npk]: G (packed lower) of the previous layer on entry, of this layer on exit
    float2* phi_cur;     // [B][n]
    float* h_cur;        // [B][n]
    float* r;            // [B]: previous layer's residual norms on entry, this layer's on exit
    const float* mean_prev;
    const float* Pk;
    const float* Pkm1;
    float2* scratch;     // [grid][3][d*d]: A (rotated in place), V, and the untouched copy A0 for the Rayleigh quotients
    int* status;
    int B, n, d, first;
};

__host__ __device__ inline size_t big_scratch_f2(int d) { return (size_t)3 * d * d; }

struct BigSmem {
    float2 phi[BIG_NMAX], gcol[BIG_NMAX], zeta[BIG_NMAX], phip[BIG_NMAX];
    float hp[BIG_NMAX], t[BIG_NMAX], h[BIG_NMAX + 1], lamp[BIG_NMAX + 1];
    float hid[64];
    float red[96];
    float4 rot[(BIG_NMAX + 2) / 2];      // (c, s, cos(theta), sin(theta)) per pair of the round
    float offsq[BIG_NT / 32];
    int pp[(BIG_NMAX + 2) / 2], qq[(BIG_NMAX + 2) / 2];
};

// Cyclic two-sided Jacobi on the Hermitian matrix A (column-major, leading dimension d, both triangles valid), V = I on
// entry is accumulated.  On exit the columns of V hold the eigenvectors and s.lamp[k] the eigenvalue of column k, taken
// as the Rayleigh quotient v_k^H A0 v_k against the copy A0 of the input made here: the diagonal of the rotated A carries
// the rounding of ~d rotations per sweep (measured 2e-5 relative at d = 129..257), the quotient only that of one dot
// product (its error is quadratic in the eigenvector error).  All BIG_NT threads.
__device__ void big_jacobi(float2* __restrict__ A, float2* __restrict__ V, float2* __restrict__ A0, int d, BigSmem& s,
                           int* status) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = BIG_NT / 32;
    const int dd = (d + 1) & ~1, npair = dd / 2;
    // ||A||_F^2
    float fro;
    {
        float v[1] = {0.f};
        for (int idx = tid; idx < d * d; idx += BIG_NT) { const float2 x = A[idx]; A0[idx] = x; v[0] += x.x * x.x + x.y * x.y; }
        block_sum<1>(v, s.red);
        fro = v[0];
    }
    if (!(fro < INFINITY)) {                             // non-finite input
        if (tid == 0) atomicOr(status, 4);
        for (int i = tid; i < d; i += BIG_NT) s.lamp[i] = A[i + (size_t)i * d].x;
        __syncthreads();
        return;
    }
    const float stop = 4e-14f * fro;                   // (2e-7)^2 ||A||_F^2
    const float skip = 1e-18f * fro;                   // rotations below this are exact no-ops in fp32
    bool converged = false;
    for (int sweep = 0; sweep < BIG_MAX_SWEEPS; ++sweep) {
        float off_acc = 0.f;                           // sum of |a_pq|^2 met by this thread's warp (lane 0 keeps it)
        for (int r = 0; r < dd - 1; ++r) {
            // ---- (1) rotations of the round + column updates of A and V
            for (int i = wid; i < npair; i += nw) {
                int p, q;
                if (i == 0) { p = r; q = dd - 1; }
                else { p = (r + i) % (dd - 1); q = (r - i + (dd - 1)) % (dd - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                float4 rt = make_float4(1.f, 0.f, 1.f, 0.f);
                bool act = false;
                if (q < d) {
                    const float app = A[p + (size_t)p * d].x, aqq = A[q + (size_t)q * d].x;
                    const float2 apq = A[p + (size_t)q * d];
                    const float b2 = apq.x * apq.x + apq.y * apq.y;
                    if (lane == 0) off_acc += b2;
                    if (b2 > skip) {
                        const float ab = sqrtf(b2);
                        const float tau = (aqq - app) / (2.f * ab);
                        const float tt = copysignf(1.f, tau) / (fabsf(tau) + sqrtf(1.f + tau * tau));
                        const float c = 1.f / sqrtf(1.f + tt * tt);    // IEEE sqrt and division: c^2 + s^2 = 1 without the bias of rsqrtf
                        rt = make_float4(c, tt * c, apq.x / ab, apq.y / ab);
                        act = true;
                    }
                }
                if (lane == 0) { s.rot[i] = rt; s.pp[i] = p; s.qq[i] = act ? q : -1; }
                if (act) {
                    // J = [[c, s],[-s e^{-i th}, c e^{-i th}]]:  col_p' = c col_p - s e^{-i th} col_q,
                    //                                           col_q' = s col_p + c e^{-i th} col_q
                    const float2 em = make_float2(rt.z, -rt.w);                    // e^{-i theta}
                    const float2 se = cscale(rt.y, em), ce = cscale(rt.x, em);
                    float2* Ap = A + (size_t)p * d; float2* Aq = A + (size_t)q * d;
                    float2* Vp = V + (size_t)p * d; float2* Vq = V + (size_t)q * d;
                    for (int row = lane; row < d; row += 32) {
                        const float2 xp = Ap[row], xq = Aq[row];
                        Ap[row] = csub(cscale(rt.x, xp), cmul(se, xq));
                        Aq[row] = cadd(cscale(rt.y, xp), cmul(ce, xq));
                        const float2 vp = Vp[row], vq = Vq[row];
                        Vp[row] = csub(cscale(rt.x, vp), cmul(se, vq));
                        Vq[row] = cadd(cscale(rt.y, vp), cmul(ce, vq));
                    }
                }
            }
            __syncthreads();
            // ---- (2) row updates: row_p' = c row_p - s e^{i th} row_q,  row_q' = s row_p + c e^{i th} row_q
            for (int i = wid; i < npair; i += nw) {
                const int q = s.qq[i];
                if (q < 0) continue;
                const int p = s.pp[i];
                const float4 rt = s.rot[i];
                const float2 ep = make_float2(rt.z, rt.w);
                const float2 se = cscale(rt.y, ep), ce = cscale(rt.x, ep);
                for (int col = lane; col < d; col += 32) {
                    const float2 xp = A[p + (size_t)col * d], xq = A[q + (size_t)col * d];
                    float2 np_ = csub(cscale(rt.x, xp), cmul(se, xq));
                    float2 nq_ = cadd(cscale(rt.y, xp), cmul(ce, xq));
                    if (col == p) np_.y = 0.f;                       // diagonal stays real
                    if (col == q) { nq_.y = 0.f; np_ = make_float2(0.f, 0.f); }      // a_pq := 0
                    if (col == p) nq_ = make_float2(0.f, 0.f);                        // a_qp := 0
                    A[p + (size_t)col * d] = np_;
                    A[q + (size_t)col * d] = nq_;
                }
            }
            __syncthreads();
        }
        if (lane == 0) s.offsq[wid] = off_acc;
        __syncthreads();
        float tot = 0.f;
        for (int w = 0; w < nw; ++w) tot += s.offsq[w];
        __syncthreads();
        if (2.f * tot <= stop) { converged = true; break; }
    }
    if (!converged && tid == 0) atomicOr(status, 2);   // not converged within BIG_MAX_SWEEPS sweeps
    // the columns of V went through ~d rotations per sweep: renormalise them (removes the accumulated norm drift)
    for (int k = wid; k < d; k += nw) {
        float2* Vk = V + (size_t)k * d;
        float n2 = 0.f;
        for (int row = lane; row < d; row += 32) { const float2 v = Vk[row]; n2 = fmaf(v.x, v.x, fmaf(v.y, v.y, n2)); }
        n2 = warp_sum(n2);
        const float inv = 1.f / sqrtf(n2);
        for (int row = lane; row < d; row += 32) { const float2 v = Vk[row]; Vk[row] = make_float2(v.x * inv, v.y * inv); }
        __syncwarp();
        // Rayleigh quotient: lane owns rows lane, lane+32, ... of A0 v_k (A0 column-major: coalesced along rows)
        float rq = 0.f;
        for (int row = lane; row < d; row += 32) {
            float ax = 0.f, ay = 0.f;
            for (int j = 0; j < d; ++j) {
                const float2 m = A0[row + (size_t)j * d], v = Vk[j];
                ax = fmaf(m.x, v.x, fmaf(-m.y, v.y, ax));
                ay = fmaf(m.x, v.y, fmaf(m.y, v.x, ay));
            }
            const float2 v = Vk[row];
            rq = fmaf(v.x, ax, fmaf(v.y, ay, rq));
        }
        rq = warp_sum(rq);
        if (lane == 0) s.lamp[k] = rq;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BIG_NT, 2) k_big_layer(BigArgs a) {
    __shared__ BigSmem s;
    const int n = a.n, d = a.d, tid = threadIdx.x;
    const int npk = d * (d + 1) / 2;
    float2* A = a.scratch + (size_t)blockIdx.x * big_scratch_f2(d);
    float2* V = A + (size_t)d * d;
    float2* A0 = V + (size_t)d * d;
    const float* __restrict__ P = a.Pk;
    const float inv_rho_g = P[P_INV_RHO_G], rho_h_eps = P[P_RHO_H_EPS];
    for (int sig = blockIdx.x; sig < a.B; sig += gridDim.x) {
        float2* Zp = a.Zp + (size_t)sig * npk;
        float2* GV = a.GV + (size_t)sig * npk;
        for (int j = tid; j < n; j += BIG_NT) {
            s.gcol[j] = make_float2(0.f, 0.f);
            s.zeta[j] = make_float2(0.f, 0.f);
            s.t[j] = 0.f;
            if (!a.first) {
                s.phip[j] = a.phi_cur[(size_t)sig * n + j];
                s.hp[j] = a.h_cur[(size_t)sig * n + j];
            }
        }
        __syncthreads();
        // ---- dual update of the previous layer (admm_net.py:403-412), A = -Z/(rho_g+eps), V = I
        const float alpha = a.first ? 0.f : z_alpha(a.Pkm1, a.r[sig], *a.mean_prev);
        const float c1z = a.first ? 0.f : a.Pkm1[P_C1Z];
        for (int idx = tid; idx < d * d; idx += BIG_NT) {
            const int j = idx / d, i = idx - j * d;            // column j, row i
            V[idx] = make_float2(i == j ? 1.f : 0.f, 0.f);
            if (i < j) continue;                               // lower triangle drives, the mirror is written with it
            float2 z = make_float2(0.f, 0.f);
            if (!a.first) {
                z = Zp[pk(i, j)];
                const float2 g = GV[pk(i, j)];
                float2 c = make_float2(0.f, 0.f);
                if (i == j) c.x = (i < n) ? s.hp[i] : c1z;
                else if (i == n) c = cconj(s.phip[j]);
                z.x += alpha * (g.x - c.x);
                z.y += alpha * (g.y - c.y);
                if (i == j) z.y = 0.f;
                if (i == j && i < n) s.t[i] = g.x + z.x / rho_h_eps;
                if (i == n && j < n) { s.gcol[j] = cconj(g); s.zeta[j] = cconj(z); }
            }
            Zp[pk(i, j)] = z;                                   // layer 0: Z_0 = 0 (admm_net.py:754)
            const float2 av = make_float2(-inv_rho_g * z.x, -inv_rho_g * z.y);
            A[i + (size_t)j * d] = av;
            if (i != j) A[j + (size_t)i * d] = cconj(av);
        }
        __syncthreads();
        // ---- phi update (admm_net.py:94-103)
        const float rho_phi = P[P_RHO_PHI];
        for (int j = tid; j < n; j += BIG_NT) {
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
        // ---- H update (admm_net.py:146-192)
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
            const float linf = block_max(tid < n ? fabsf(tc) : 0.f, s.red);
            float sm[1] = {tid < n ? tc : 0.f};
            block_sum<1>(sm, s.red);
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
        for (int j = tid; j < n; j += BIG_NT) {
            A[j + (size_t)j * d].x += s.h[j];
            const float2 ph = s.phi[j];
            float2* up = &A[j + (size_t)n * d];
            up->x += ph.x; up->y += ph.y;
            float2* lo = &A[n + (size_t)j * d];
            lo->x += ph.x; lo->y -= ph.y;
        }
        if (tid == 0) A[n + (size_t)n * d].x += P[P_C0];
        __syncthreads();
        // ---- eigh + eigenvalue map
        big_jacobi(A, V, A0, d, s, a.status);
        for (int i = tid; i < d; i += BIG_NT) s.lamp[i] = eig_map(P, s.lamp[i]);
        __syncthreads();
        // ---- rebuild G = V diag(l') V^H (lower triangle), residual norm ||G - C||_F, packed store
        float rsq = 0.f;
        const float c1 = P[P_C1Z];
        for (int idx = tid; idx < npk; idx += BIG_NT) {
            int i = (int)((sqrtf(8.f * (float)idx + 1.f) - 1.f) * 0.5f);
            while ((i + 1) * (i + 2) / 2 <= idx) ++i;
            while (i * (i + 1) / 2 > idx) --i;
            const int j = idx - i * (i + 1) / 2;
            float gx = 0.f, gy = 0.f;
            for (int k = 0; k < d; ++k) {
                const float2 vi = V[i + (size_t)k * d], vj = V[j + (size_t)k * d];
                const float l = s.lamp[k];
                gx = fmaf(l, vi.x * vj.x + vi.y * vj.y, gx);
                gy = fmaf(l, vi.y * vj.x - vi.x * vj.y, gy);
            }
            if (i == j) gy = 0.f;
            GV[idx] = make_float2(gx, gy);
            float2 c = make_float2(0.f, 0.f);
            if (i == j) c.x = (i < n) ? s.h[i] : c1;
            else if (i == n) c = cconj(s.phi[j]);
            const float rx = gx - c.x, ry = gy - c.y;
            rsq += (i == j ? 1.f : 2.f) * (rx * rx + ry * ry);
        }
        {
            float v[1] = {rsq};
            block_sum<1>(v, s.red);
            if (tid == 0) a.r[sig] = sqrtf(v[0]);
        }
        __syncthreads();
    }
}

// Debug/unit entry (the eigh tap for d > 128): A complex64 [B][d][d] row-major, lower triangle read.
//   evals [B][d] (unsorted: Jacobi order), evecs [B][d][d] row-major (columns = eigenvectors),
//   fn_out [B][npk] packed lower triangle of f(A) (eigenvalue map of Pk, or the identity map when Pk == nullptr)
__global__ void __launch_bounds__(BIG_NT, 2)
k_big_eigh(const float2* __restrict__ Afull, int B, int d, float* evals, float2* evecs, float2* fn_out, const float* Pk,
           float2* scratch, int* status) {
    __shared__ BigSmem s;
    const int tid = threadIdx.x, npk = d * (d + 1) / 2;
    float2* A = scratch + (size_t)blockIdx.x * big_scratch_f2(d);
    float2* V = A + (size_t)d * d;
    float2* A0 = V + (size_t)d * d;
    for (int sig = blockIdx.x; sig < B; sig += gridDim.x) {
        const float2* Ag = Afull + (size_t)sig * d * d;
        for (int idx = tid; idx < d * d; idx += BIG_NT) {
            const int i = idx / d, j = idx - i * d;             // row i, column j of the row-major input
            V[idx] = make_float2(i == j ? 1.f : 0.f, 0.f);
            if (j > i) continue;
            float2 v = Ag[idx];
            if (i == j) v.y = 0.f;
            A[i + (size_t)j * d] = v;
            if (i != j) A[j + (size_t)i * d] = cconj(v);
        }
        __syncthreads();
        big_jacobi(A, V, A0, d, s, status);
        for (int i = tid; i < d; i += BIG_NT) {
            const float l = s.lamp[i];
            if (evals) evals[(size_t)sig * d + i] = l;
            s.lamp[i] = Pk ? eig_map(Pk, l) : l;
        }
        if (evecs)
            for (int idx = tid; idx < d * d; idx += BIG_NT) {
                const int i = idx / d, k = idx - i * d;
                evecs[(size_t)sig * d * d + idx] = V[i + (size_t)k * d];
            }
        __syncthreads();
        if (fn_out)
            for (int idx = tid; idx < npk; idx += BIG_NT) {
                int i = (int)((sqrtf(8.f * (float)idx + 1.f) - 1.f) * 0.5f);
                while ((i + 1) * (i + 2) / 2 <= idx) ++i;
                while (i * (i + 1) / 2 > idx) --i;
                const int j = idx - i * (i + 1) / 2;
                float gx = 0.f, gy = 0.f;
                for (int k = 0; k < d; ++k) {
                    const float2 vi = V[i + (size_t)k * d], vj = V[j + (size_t)k * d];
                    const float l = s.lamp[k];
                    gx = fmaf(l, vi.x * vj.x + vi.y * vj.y, gx);
                    gy = fmaf(l, vi.y * vj.x - vi.x * vj.y, gy);
                }
                if (i == j) gy = 0.f;
                fn_out[(size_t)sig * npk + idx] = make_float2(gx, gy);
            }
        __syncthreads();
    }
}

}  // namespace admmnet
