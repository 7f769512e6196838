// On-device synthetic OFDM-radar signal generator.
// Restates the reference's per-sample recipe, generate_data.py:133-221 (`_generate_single_sample`,
// `_generate_communication_symbols`) with utils/mathUtils.py:4-21 (vander_vec), 24-50 (kr), 53-111 (pskmod,
// pskdemod, awgn), one warp per signal, in fp64 like the reference, cast to complex64 / fp32 on output
// (generate_data.py:196-201).  Randomness: counter-based Philox-4x32-10 keyed by (seed, signal index), so a batch is
// reproducible and independent of the launch geometry; the draws are statistically — not bitwise — numpy's.
#include "common.cuh"

namespace admmnet {

struct Philox {
    uint32_t key[2], ctr[4], out[4];
    int have;
    __device__ Philox(uint64_t seed, uint64_t stream) {
        key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
        ctr[0] = 0; ctr[1] = 0; ctr[2] = (uint32_t)stream; ctr[3] = (uint32_t)(stream >> 32);
        have = 0;
    }
    __device__ void round_once(uint32_t k0, uint32_t k1, uint32_t* c) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
    }
    __device__ void refill() {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k0 = key[0], k1 = key[1];
#pragma unroll
        for (int r = 0; r < 10; ++r) { round_once(k0, k1, c); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
        if (++ctr[0] == 0) ++ctr[1];
        have = 4;
    }
    __device__ uint32_t u32() { if (!have) refill(); return out[--have]; }
    __device__ double uniform() {      // (0,1), 53 bits
        const uint64_t a = u32() >> 5, b = u32() >> 6;
        return ((double)a * 67108864.0 + (double)b + 0.5) * (1.0 / 9007199254740992.0);
    }
    __device__ double2 normal2() {     // Box-Muller
        const double u1 = uniform(), u2 = uniform();
        const double r = sqrt(-2.0 * log(u1));
        double s, c;
        sincospi(2.0 * u2, &s, &c);
        return make_double2(r * c, r * s);
    }
};

struct GenArgs {
    float2* y;        // [B][n]
    float2* b;        // [B][n]
    float* sigma;     // [B]
    double* truth;    // optional [B][L][4]: tau, f, Re C, Im C
    float* ser;       // optional [B]: symbol error rate in percent (generate_data.py:220)
    int B, Nb, Nd, L;
    double snr_w_db, snr_w_hi_db, snr_demod_db;   // noise SNR drawn uniformly in [snr_w_db, snr_w_hi_db) when hi > lo (:164)
    unsigned long long seed;
};

#define GEN_MAXL 8
__global__ void __launch_bounds__(256) k_generate(GenArgs a) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= a.B) return;
    const int n = a.Nb * a.Nd, L = a.L;
    // per-signal scalars: every lane draws the same values from the signal's stream 0
    Philox g0(a.seed, (uint64_t)w * 64);
    double tau[GEN_MAXL], f[GEN_MAXL], cr[GEN_MAXL], ci[GEN_MAXL];
    for (int l = 0; l < L; ++l) tau[l] = 0.1 + 0.8 * g0.uniform();            // generate_data.py:30,138
    for (int l = 0; l < L; ++l) f[l] = -0.4 + 0.8 * g0.uniform();             // :31,139
    for (int l = 0; l < L; ++l) { const double2 z = g0.normal2(); cr[l] = 0.7 * z.x; ci[l] = 0.7 * z.y; }   // :142-144
    if (a.truth && lane == 0)
        for (int l = 0; l < L; ++l) {
            double* t = a.truth + ((size_t)w * L + l) * 4;
            t[0] = tau[l]; t[1] = f[l]; t[2] = cr[l]; t[3] = ci[l];
        }
    double snr_w = a.snr_w_db;
    if (a.snr_w_hi_db > a.snr_w_db) snr_w += (a.snr_w_hi_db - a.snr_w_db) * g0.uniform();
    const double npow = 1.0 / pow(10.0, a.snr_demod_db / 10.0);               // awgn on unit-power PSK
    // per-element work: lane handles elements lane, lane+32, ... with its own stream
    Philox g(a.seed, (uint64_t)w * 64 + 1 + lane);
    double ynorm2 = 0.0, er2 = 0.0;
    int nerr = 0;
    double2 yv[8], bv[8], wv[8];
    for (int e = 0; e < 8; ++e) {
        const int i = lane + 32 * e;
        yv[e] = bv[e] = wv[e] = make_double2(0.0, 0.0);
        if (i >= n) continue;
        const int p = i / a.Nd, q = i % a.Nd;                                  // a = kron(s(f), conj(d(tau)))
        double2 psi = make_double2(0.0, 0.0);
        for (int l = 0; l < L; ++l) {                                          // Psi = kr(S, conj(D)) @ C  (:152-156)
            double ss, sc, ds, dc;
            sincospi(2.0 * (double)p * f[l], &ss, &sc);
            sincospi(2.0 * (double)q * tau[l], &ds, &dc);
            const double ar = sc * dc + ss * ds, ai = ss * dc - sc * ds;       // s * conj(d)
            psi.x += ar * cr[l] - ai * ci[l];
            psi.y += ar * ci[l] + ai * cr[l];
        }
        const int sym = (int)(g.u32() & 3u);                                   // QPSK symbol (:208)
        double s_im, s_re;
        sincospi(0.5 * sym + 0.25, &s_im, &s_re);                              // pskmod(data, 4, pi/4)
        const double2 nz = g.normal2();
        const double rx = s_re + sqrt(npow * 0.5) * nz.x, ry = s_im + sqrt(npow * 0.5) * nz.y;   // awgn
        double ang = atan2(ry, rx) - 0.78539816339744830962 + 0.78539816339744830962;            // pskdemod
        ang = fmod(ang, 6.283185307179586);
        if (ang < 0) ang += 6.283185307179586;
        const int dsym = ((int)floor(ang * 4.0 / 6.283185307179586)) & 3;
        double b_im, b_re;
        sincospi(0.5 * dsym + 0.25, &b_im, &b_re);                             // re-modulated decision b
        const double e_re = s_re - b_re, e_im = s_im - b_im;                   // e = sig - b  (:219)
        // real_y = (b + e) * Psi = sig * Psi   (:162)
        const double yr = s_re * psi.x - s_im * psi.y, yi = s_re * psi.y + s_im * psi.x;
        ynorm2 += yr * yr + yi * yi;
        const double bden = b_re * b_re + b_im * b_im;                         // |e/b|^2 = |e|^2/|b|^2
        er2 += (e_re * e_re + e_im * e_im) / bden;
        nerr += (sqrt(e_re * e_re + e_im * e_im) > 1e-10);
        yv[e] = make_double2(yr, yi);
        bv[e] = make_double2(b_re, b_im);
        wv[e] = g.normal2();
    }
    ynorm2 = warp_sum(ynorm2);
    er2 = warp_sum(er2);
    nerr = (int)warp_sum((double)nerr);
    const double w_var = ynorm2 / (pow(10.0, snr_w / 10.0) * n);               // :166
    const double amp = sqrt(w_var) * sqrt(0.5);
    for (int e = 0; e < 8; ++e) {
        const int i = lane + 32 * e;
        if (i >= n) continue;
        a.y[(size_t)w * n + i] = make_float2((float)(yv[e].x + amp * wv[e].x), (float)(yv[e].y + amp * wv[e].y));
        a.b[(size_t)w * n + i] = make_float2((float)bv[e].x, (float)bv[e].y);
    }
    if (lane == 0) a.sigma[w] = (float)(sqrt(er2) + 1.0);                      // sigma = ||e/b|| + 1  (:171)
    if (lane == 0 && a.ser) a.ser[w] = (float)(100.0 * nerr / n);
}

}  // namespace admmnet
