// Layer-0 shortcut of the unrolled net: with G = Z = 0 the matrix handed to eigh
// (admm_net.py:286-303) is the ARROWHEAD  A = [[diag(h), phi],[phi^H, c0]], whose eigen-decomposition needs no
// tridiagonalisation, QL sweep or back-transformation:
//   eigenvalues   roots of  g(l) = c0 - l + sum_i |phi_i|^2 / (l - h_i)   (one per interval between sorted poles h_i)
//   eigenvectors  x = [ phi_i / (l - h_i) ; 1 ] / ||.||
// One CTA per signal.  Roots: one thread per root, osculatory two-pole rational iteration (the scheme of LAPACK's
// xLAED4, with the linear term's slope split between the two sides) in coordinates shifted to the nearer pole, so
// that every difference l_j - h_i keeps full relative accuracy in fp32.  The moduli |phi_i| are then re-derived from
// the computed roots (Gu & Eisenstat): the vectors are the exact eigenvectors of a nearby arrowhead matrix and are
// orthogonal to rounding whatever the pole spacing.  G = U f(L) U^H and the residual norm reuse k_tail's rebuild.
// Signals the shortcut does not accept (poles closer than rounding, a vanishing |phi_i|, non-finite data) are left
// to the general pipeline: handled[sig] = 0 and those kernels skip every signal with handled[sig] = 1.
#include "common.cuh"

namespace admmnet {

#define AR_NT 384
#define AR_MAXIT 48

struct ArrowArgs {
    const float2* y;        // net mode (Pk != null): inputs of the layer-0 prologue
    const float2* b;
    const float* sigma;
    const float* h_in;      // tap mode (Pk == null): explicit arrowhead, h [B][n], phi [B][n], c0 [B]
    const float2* phi_in;
    const float* c0_in;
    float2* Zp;             // [B][npk]  zeroed (Z_0 = 0)
    float2* GV;             // [B][npk]  G packed lower
    float2* phi_cur;        // [B][n]
    float* h_cur;           // [B][n]
    const float* Pk;
    float* r_out;           // [B]
    int* handled;           // [B]
    float* lam_out;         // tap: [B][d] eigenvalues, ascending
    float2* U_out;          // tap: [B][d][d] row-major eigenvectors (column j pairs with lam_out[j])
    int B, n, d, ldu;
};
__host__ __device__ inline size_t arrow_smem_bytes(int d, int ldu) {
    return ((size_t)d * ldu + 2 * 128) * sizeof(float2) + (size_t)(10 * 128 + 64 + 96) * sizeof(float) + 2 * 128 * sizeof(int);
}

// g and the slopes of its left / right pole sums at x (shifted coordinates, origin `org`); poles [0, jsplit) lie left
__device__ __forceinline__ void arrow_eval(const float* __restrict__ sd, const float* __restrict__ sz2, int n, int jsplit,
                                           float org, float a0, float x, float& g, float& wl, float& wr, float& sabs) {
    float p0 = 0.f, p1 = 0.f, q0 = 0.f, q1 = 0.f;
    int i = 0;
    for (; i + 1 < jsplit; i += 2) {
        const float r0 = __fdividef(1.f, x - (sd[i] - org)), r1 = __fdividef(1.f, x - (sd[i + 1] - org));
        const float t0 = sz2[i] * r0, t1 = sz2[i + 1] * r1;
        p0 += t0; p1 += t1;
        q0 = fmaf(t0, r0, q0); q1 = fmaf(t1, r1, q1);
    }
    if (i < jsplit) {
        const float r0 = __fdividef(1.f, x - (sd[i] - org));
        const float t0 = sz2[i] * r0;
        p0 += t0; q0 = fmaf(t0, r0, q0);
        ++i;
    }
    float f0 = 0.f, f1 = 0.f, h0 = 0.f, h1 = 0.f;
    for (; i + 1 < n; i += 2) {
        const float r0 = __fdividef(1.f, x - (sd[i] - org)), r1 = __fdividef(1.f, x - (sd[i + 1] - org));
        const float t0 = sz2[i] * r0, t1 = sz2[i + 1] * r1;
        f0 += t0; f1 += t1;
        h0 = fmaf(t0, r0, h0); h1 = fmaf(t1, r1, h1);
    }
    if (i < n) {
        const float r0 = __fdividef(1.f, x - (sd[i] - org));
        const float t0 = sz2[i] * r0;
        f0 += t0; h0 = fmaf(t0, r0, h0);
    }
    const float psi = p0 + p1, phi = f0 + f1;
    g = (a0 - x) + (psi + phi);
    wl = -(q0 + q1) - 0.5f;          // d/dx of the left sum, plus half of the linear term's slope
    wr = -(h0 + h1) - 0.5f;
    sabs = fabsf(psi) + fabsf(phi);
}

__global__ void __launch_bounds__(AR_NT, 2) k_arrow(ArrowArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, d = a.d, ldu = a.ldu;
    float2* U = reinterpret_cast<float2*>(smem_raw);        // [d][ldu] column-major
    float2* phis = U + (size_t)d * ldu;                      // [128] phi, original order
    float2* zph = phis + 128;                                // [128] zhat_i * phase_i, sorted order
    float* hs = reinterpret_cast<float*>(zph + 128);         // [128] h, original order
    float* sd = hs + 128;                                    // [128] sorted poles
    float* sz2 = sd + 128;                                   // [128] |phi|^2 sorted, later zhat^2
    float* org = sz2 + 128;                                  // [128] origin pole of each root
    float* xs = org + 128;                                   // [128] root offset from its origin
    float* lamv = xs + 128;                                  // [128] eigenvalues
    float* lamp = lamv + 128;                                // [128] mapped eigenvalues
    float* nu = lamp + 128;                                  // [128] 1/||x_j||
    float* tc_s = nu + 128;                                  // [128] scratch
    float* spare = tc_s + 128;                               // [128]
    float* hid = spare + 128;                                // [64]
    float* red = hid + 64;                                   // [96]
    int* perm = reinterpret_cast<int*>(red + 96);            // [128] sorted position -> original index
    int* rnk = perm + 128;                                   // [128] original index -> sorted position
    const int tid = threadIdx.x;
    const int sig = blockIdx.x;
    const int npk = d * (d + 1) / 2;
    const float* __restrict__ P = a.Pk;
    const bool net = P != nullptr;
    float2* GV = a.GV ? a.GV + (size_t)sig * npk : nullptr;
    float c0;

    if (net) {
        // ---- layer-0 prologue: phi = w * y/(b+eps), h from the correction MLP at t = 0 (admm_net.py:94-103, 146-192)
        float2* Zp = a.Zp + (size_t)sig * npk;
        for (int idx = tid; idx < npk; idx += AR_NT) Zp[idx] = make_float2(0.f, 0.f);
        const float rho_phi = P[P_RHO_PHI];
        for (int j = tid; j < n; j += AR_NT) {
            const float2 bj = a.b[(size_t)sig * n + j], yj = a.y[(size_t)sig * n + j];
            const float ab = hypotf(bj.x, bj.y);
            const float bsq = ab * ab + ADMM_EPS;
            const float wgt = bsq / (1.f + rho_phi * bsq);
            const float2 yob = cdiv(yj, make_float2(bj.x + ADMM_EPS, bj.y));
            const float2 ph = make_float2(wgt * (yob.x + rho_phi * 0.f + 0.f), wgt * (yob.y + rho_phi * 0.f + 0.f));
            phis[j] = ph;
            a.phi_cur[(size_t)sig * n + j] = ph;
        }
        if (tid < 64) hid[tid] = fmaxf(P[P_HB1 + tid], 0.f);
        __syncthreads();
        float tc = 0.f;
        if (tid < n) {
            const float* __restrict__ W2T = P + P_HW1T + 64 * n;
            float acc = P[P_HW1T + 128 * n + tid];
#pragma unroll 8
            for (int j = 0; j < 64; ++j) acc += W2T[j * n + tid] * hid[j];
            tc = 0.f + 0.1f * tanhf(acc);
        }
        const float linf = block_max(tid < n ? fabsf(tc) : 0.f, red);
        float sm[1] = {tid < n ? tc : 0.f};
        block_sum<1>(sm, red);
        const float sg = a.sigma[sig];
        const float Asig = 2.f * sqrtf((float)n) * sg + sg * sg;
        const float cv = Asig * linf + sm[0];
        const float scale = fminf(P[P_SIG_PW] / (cv + ADMM_EPS), 1.f);
        if (tid < n) {
            const float hv = tc * scale;
            hs[tid] = hv;
            a.h_cur[(size_t)sig * n + tid] = hv;
        }
        c0 = P[P_C0];
    } else {
        for (int j = tid; j < n; j += AR_NT) {
            phis[j] = a.phi_in[(size_t)sig * n + j];
            hs[j] = a.h_in[(size_t)sig * n + j];
        }
        c0 = a.c0_in[sig];
    }
    __syncthreads();

    // ---- moduli, rank sort of the poles
    float zi2 = 0.f, hi_ = 0.f;
    if (tid < n) {
        const float2 ph = phis[tid];
        zi2 = fmaf(ph.x, ph.x, ph.y * ph.y);
        hi_ = hs[tid];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float hj = hs[j];
            rank += (hj < hi_ || (hj == hi_ && j < tid)) ? 1 : 0;
        }
        sd[rank] = hi_;
        sz2[rank] = zi2;
        perm[rank] = tid;
        rnk[tid] = rank;
    }
    float zn[1] = {zi2};
    block_sum<1>(zn, red);
    const float znorm2 = zn[0];
    const float hmax = block_max(tid < n ? fabsf(hi_) : 0.f, red);
    __syncthreads();
    const float znorm = sqrtf(znorm2);
    const float scl = fmaxf(fmaxf(hmax, fabsf(c0)), znorm);
    {
        int bad = 0;
        if (tid < n) {
            // Distinct poles and non-zero couplings are all the solver needs (pole differences of fp32 inputs are
            // exact, every l_j - h_i is formed from them); the floors only keep the arithmetic away from underflow.
            if (!(sz2[tid] > 1e-24f * scl * scl)) bad = 1;
            if (tid + 1 < n && !(sd[tid + 1] - sd[tid] > 1e-20f * scl)) bad = 1;
        }
        if (tid == 0 && !(scl < 1e18f && scl > 1e-18f)) bad = 1;                 // also catches NaN / Inf
        if (__syncthreads_or(bad)) {
            if (tid == 0) a.handled[sig] = 0;
            return;
        }
    }

    // ---- secular roots, one thread per root j = 0..n: root j lies in (sd[j-1], sd[j])
    int failed = 0;
    if (tid < d) {
        const int j = tid;
        float o, x, lo, hi, dL = 0.f, dR = 0.f;
        const bool first = (j == 0), last = (j == n);
        if (first || last) {          // all poles collapsed onto the nearest one bound the root from outside
            o = first ? sd[0] : sd[n - 1];
            const float ap = c0 - o;
            const float rt = sqrtf(fmaf(ap, ap, 4.f * znorm2));
            if (first) { x = ap > 0.f ? -2.f * znorm2 / (ap + rt) : 0.5f * (ap - rt); lo = x * 1.0001f - 1e-30f; hi = 0.f; }
            else { x = ap < 0.f ? 2.f * znorm2 / (rt - ap) : 0.5f * (ap + rt); hi = x * 1.0001f + 1e-30f; lo = 0.f; }
            if (x == 0.f) x = 0.5f * (lo + hi);
        } else {
            const float gap = sd[j] - sd[j - 1];
            o = sd[j - 1];
            float g, wl, wr, sa;
            arrow_eval(sd, sz2, n, j, o, c0 - o, 0.5f * gap, g, wl, wr, sa);
            if (g > 0.f) { o = sd[j]; lo = -0.5f * gap; hi = 0.f; dL = -gap; dR = 0.f; x = lo; }   // root in the right half
            else { lo = 0.f; hi = 0.5f * gap; dL = 0.f; dR = gap; x = hi; }
        }
        const float a0 = c0 - o;
        bool conv = false;
        for (int it = 0; it < AR_MAXIT && !conv; ++it) {
            float g, wl, wr, sa;
            arrow_eval(sd, sz2, n, j, o, a0, x, g, wl, wr, sa);
            if (g > 0.f) lo = x; else hi = x;
            if (fabsf(g) <= 1.2e-7f * (8.f * sa + fabsf(a0) + fabsf(x))) break;
            float eta;
            if (first || last) {
                const float w = wl + wr, D = x;                 // single pole at the origin
                const float den = g + w * D;
                eta = den != 0.f ? -g * D / den : 0.f;
            } else {
                const float DL = x - dL, DR = x - dR;
                const float s = -wl * DL * DL, S = -wr * DR * DR;
                const float C = g + wl * DL + wr * DR;
                const float a1 = C * (DL + DR) + s + S, a0q = DL * DR * g;
                const float disc = fmaxf(fmaf(a1, a1, -4.f * C * a0q), 0.f);
                const float q = a1 + copysignf(sqrtf(disc), a1);
                eta = q != 0.f ? -2.f * a0q / q : 0.f;
                float xn = x + eta;
                if (!(xn > lo && xn < hi) && C != 0.f && eta != 0.f) eta = a0q / (C * eta);   // the other root of the model
            }
            float xn = x + eta;
            if (!(xn > lo && xn < hi)) xn = 0.5f * (lo + hi);
            if (xn == x || fabsf(xn - x) <= 6e-8f * fabsf(xn)) conv = true;
            if (hi - lo <= 1.2e-7f * fmaxf(fabsf(lo), fabsf(hi))) conv = true;
            x = xn;
            if (it == AR_MAXIT - 1 && !conv) failed = 1;
        }
        if (!(x == x)) failed = 1;
        org[j] = o;
        xs[j] = x;
        lamv[j] = o + x;
    }
    if (__syncthreads_or(failed)) {
        if (tid == 0) a.handled[sig] = 0;
        return;
    }

    // ---- |phi_i| consistent with the computed roots (Gu-Eisenstat), phase restored
    if (tid < n) {
        const int i = tid;
        const float si = sd[i];
        float prod = ((si - org[0]) - xs[0]) * ((org[n] - si) + xs[n]);
        for (int j = 1; j <= i; ++j) prod *= ((si - org[j]) - xs[j]) / (si - sd[j - 1]);
        for (int j = i + 1; j < n; ++j) prod *= ((org[j] - si) + xs[j]) / (sd[j] - si);
        const float zh2 = fmaxf(prod, 0.f);
        const float2 ph = phis[perm[i]];
        const float inv = rsqrtf(fmaf(ph.x, ph.x, ph.y * ph.y));
        const float zh = sqrtf(zh2);
        zph[i] = make_float2(zh * ph.x * inv, zh * ph.y * inv);
        tc_s[i] = zh2;
    }
    __syncthreads();
    if (tid < d) {
        const int j = tid;
        const float o = org[j], x = xs[j];
        float s0 = 0.f, s1 = 0.f;
        int i = 0;
        for (; i + 1 < n; i += 2) {
            const float r0 = __fdividef(1.f, (o - sd[i]) + x), r1 = __fdividef(1.f, (o - sd[i + 1]) + x);
            s0 = fmaf(tc_s[i] * r0, r0, s0);
            s1 = fmaf(tc_s[i + 1] * r1, r1, s1);
        }
        if (i < n) {
            const float r0 = __fdividef(1.f, (o - sd[i]) + x);
            s0 = fmaf(tc_s[i] * r0, r0, s0);
        }
        nu[j] = rsqrtf(1.f + s0 + s1);
        const float l = lamv[j];
        lamp[j] = net ? eig_map(P, l) : l;
    }
    __syncthreads();
    // ---- U (column-major in shared memory, rows in original order, padding rows zero)
    for (int idx = tid; idx < d * ldu; idx += AR_NT) {
        const int j = idx / ldu, r = idx - j * ldu;
        float2 v = make_float2(0.f, 0.f);
        if (r < n) {
            const int i = rnk[r];
            const float w = nu[j] / ((org[j] - sd[i]) + xs[j]);
            const float2 z = zph[i];
            v = make_float2(z.x * w, z.y * w);
        } else if (r == n) {
            v.x = nu[j];
        }
        U[idx] = v;
    }
    __syncthreads();
    if (tid == 0) a.handled[sig] = 1;
    if (!net) {
        if (a.lam_out)
            for (int j = tid; j < d; j += AR_NT) a.lam_out[(size_t)sig * d + j] = lamv[j];
        if (a.U_out)
            for (int idx = tid; idx < d * d; idx += AR_NT) {
                const int r = idx / d, j = idx - r * d;
                a.U_out[(size_t)sig * d * d + idx] = U[(size_t)j * ldu + r];
            }
        return;
    }
    // ---- G = U f(L) U^H, r = ||G - C||_F
    const float rsq = rebuild_lower<AR_NT, false>(U, ldu, lamp, d, n, GV, hs, phis, P[P_C1Z], true);
    float v[1] = {rsq};
    block_sum<1>(v, red);
    if (tid == 0) a.r_out[sig] = sqrtf(v[0]);
}

}  // namespace admmnet
