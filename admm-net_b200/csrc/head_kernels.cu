// Learned peak-regression head of the full ADMMNet (reference admm_net.py:494-630, PeakSearchLayer.forward),
// SURVEY.md §8f rank 3.  One fused kernel: feature MLP 2n -> 128 -> 128, single-query multi-head attention over
// the n fixed position tokens (keys/values depend only on the parameters, so they are projected once on the host:
// admm-net_b200/params.py::pack_head), residual, peak MLP 128 -> 64 -> 32 -> 16, and the three per-target regressors.
// CTA = 128 threads = one thread per hidden unit, HS signals per CTA so that every weight (read once, coalesced,
// from L2) feeds HS FMAs.  The head is 0.03 % of the forward's flops; the kernel is a latency/L2-bound epilogue.
#include "common.cuh"

namespace admmnet {

#define HD 128          // hidden_dim
#define HS 8            // signals per CTA
#define HEADS 4

struct HeadArgs2 {
    const float2* phi;      // [B][n]
    const float* P;         // packed head parameters (params.py::pack_head)
    float* tau;             // [B][L]
    float* f;               // [B][L]
    float* conf;            // [B][L]
    int B, n, L;
};
// offsets into P (floats), all weights stored TRANSPOSED [in][out] for coalesced reads
struct HeadOff {
    int w1, b1, w2, b2, wq, bq, kt, v, wo, bo, p1, pb1, p2, pb2, p3, pb3, reg;   // reg: per target block
};
__host__ __device__ inline HeadOff head_offsets(int n, int L) {
    HeadOff o;
    int p = 0;
    o.w1 = p; p += 2 * n * HD; o.b1 = p; p += HD;
    o.w2 = p; p += HD * HD; o.b2 = p; p += HD;
    o.wq = p; p += HD * HD; o.bq = p; p += HD;
    o.kt = p; p += HD * n;          // K^T [HD][n]
    o.v = p; p += n * HD;           // V [n][HD]
    o.wo = p; p += HD * HD; o.bo = p; p += HD;
    o.p1 = p; p += HD * 64; o.pb1 = p; p += 64;
    o.p2 = p; p += 64 * 32; o.pb2 = p; p += 32;
    o.p3 = p; p += 32 * 16; o.pb3 = p; p += 16;
    o.reg = p;                      // per target t: tauW1T[16][32], taub1[32], tauW2[32], taub2, fW1T, fb1, fW2, fb2 ; then conf net
    (void)L;
    return o;
}
__host__ __device__ inline int head_reg_stride() { return 2 * (16 * 32 + 32 + 32 + 1); }
__host__ __device__ inline int head_param_count(int n, int L) {
    return head_offsets(n, L).reg + L * head_reg_stride() + (16 * 16 + 16 + 16 + 1);
}

// out[s][j] = act(bias[j] + sum_i in[s][i] * WT[i][j]) for j = tid < nout; in/out in shared memory
template <int ACT>   // 0 none, 1 relu
__device__ __forceinline__ void dense(const float* __restrict__ WT, const float* __restrict__ bias, const float* in,
                                      int ldin, float* out, int ldout, int nin, int nout) {
    const int j = threadIdx.x;
    if (j < nout) {
        float acc[HS];
        const float bj = bias[j];
#pragma unroll
        for (int s = 0; s < HS; ++s) acc[s] = bj;
        for (int i = 0; i < nin; ++i) {
            const float w = WT[(size_t)i * nout + j];
#pragma unroll
            for (int s = 0; s < HS; ++s) acc[s] = fmaf(in[s * ldin + i], w, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < HS; ++s) out[s * ldout + j] = ACT == 1 ? fmaxf(acc[s], 0.f) : acc[s];
    }
}

__global__ void __launch_bounds__(HD) k_peak_head(HeadArgs2 a) {
    extern __shared__ __align__(16) float hsm[];
    const int n = a.n, L = a.L, tid = threadIdx.x;
    const HeadOff o = head_offsets(n, L);
    const float* __restrict__ P = a.P;
    float* x0 = hsm;                       // [HS][2n]
    float* xa = x0 + HS * 2 * n;           // [HS][HD]
    float* xb = xa + HS * HD;              // [HS][HD]
    float* xc = xb + HS * HD;              // [HS][HD]
    float* sc = xc + HS * HD;              // [HS][HEADS][n] attention scores / weights
    const int s0 = blockIdx.x * HS;
    // features = cat(phi.real, phi.imag)   (admm_net.py:585-589)
    for (int idx = tid; idx < HS * n; idx += HD) {
        const int s = idx / n, i = idx % n;
        float2 v = make_float2(0.f, 0.f);
        if (s0 + s < a.B) v = a.phi[(size_t)(s0 + s) * n + i];
        x0[s * 2 * n + i] = v.x;
        x0[s * 2 * n + n + i] = v.y;
    }
    __syncthreads();
    dense<1>(P + o.w1, P + o.b1, x0, 2 * n, xa, HD, 2 * n, HD);      // feature_extractor.0 + ReLU
    __syncthreads();
    dense<1>(P + o.w2, P + o.b2, xa, HD, xb, HD, HD, HD);            // feature_extractor.2 + ReLU  -> x (xb)
    __syncthreads();
    dense<0>(P + o.wq, P + o.bq, xb, HD, xa, HD, HD, HD);            // q = in_proj_q(x)            -> xa
    __syncthreads();
    // scores[s][h][k] = q_h . K[k]_h / sqrt(32) ; thread = key k
    if (tid < n) {
        const float* __restrict__ KT = P + o.kt;
        for (int s = 0; s < HS; ++s) {
#pragma unroll
            for (int h = 0; h < HEADS; ++h) {
                float acc = 0.f;
                for (int c = 32 * h; c < 32 * h + 32; ++c) acc = fmaf(xa[s * HD + c], KT[(size_t)c * n + tid], acc);
                sc[(s * HEADS + h) * n + tid] = acc * 0.17677669529663687f;      // 1/sqrt(head_dim = 32)
            }
        }
    }
    __syncthreads();
    // softmax over the n keys: one warp per (signal, head) pair in turn
    {
        const int lane = tid & 31, wid = tid >> 5;
        for (int pair = wid; pair < HS * HEADS; pair += HD / 32) {
            float* row = sc + pair * n;
            float m = -INFINITY;
            for (int k = lane; k < n; k += 32) m = fmaxf(m, row[k]);
            m = warp_max(m);
            float sum = 0.f;
            for (int k = lane; k < n; k += 32) { const float e = expf(row[k] - m); row[k] = e; sum += e; }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            for (int k = lane; k < n; k += 32) row[k] *= inv;
        }
    }
    __syncthreads();
    // attended heads: xc[s][c] = sum_k w[s][h(c)][k] V[k][c]
    {
        const float* __restrict__ V = P + o.v;
        float acc[HS];
#pragma unroll
        for (int s = 0; s < HS; ++s) acc[s] = 0.f;
        const int h = tid >> 5;
        for (int k = 0; k < n; ++k) {
            const float v = V[(size_t)k * HD + tid];
#pragma unroll
            for (int s = 0; s < HS; ++s) acc[s] = fmaf(sc[(s * HEADS + h) * n + k], v, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < HS; ++s) xc[s * HD + tid] = acc[s];
    }
    __syncthreads();
    dense<0>(P + o.wo, P + o.bo, xc, HD, xa, HD, HD, HD);            // out_proj -> xa
    __syncthreads();
    for (int s = 0; s < HS; ++s) xa[s * HD + tid] += xb[s * HD + tid];   // x + attended   (admm_net.py:607)
    __syncthreads();
    dense<1>(P + o.p1, P + o.pb1, xa, HD, xc, HD, HD, 64);           // peak_extractor
    __syncthreads();
    dense<1>(P + o.p2, P + o.pb2, xc, HD, xb, HD, 64, 32);
    __syncthreads();
    dense<1>(P + o.p3, P + o.pb3, xb, HD, xc, HD, 32, 16);           // x_peak -> xc[s][0..15]
    __syncthreads();
    // regressors: thread = (signal s, target t, kind): kind 0 tau (sigmoid), 1 f (tanh), 2 confidence (sigmoid)
    for (int job = tid; job < HS * L * 3; job += HD) {
        const int s = job / (L * 3), t = (job / 3) % L, kind = job % 3;
        if (s0 + s >= a.B) continue;
        const float off = (float)t / (float)L;                        // query_offset (admm_net.py:617-618)
        float feat[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) feat[i] = xc[s * HD + i] + off;
        float outv;
        if (kind < 2) {
            const float* R = P + o.reg + t * head_reg_stride() + kind * (16 * 32 + 32 + 32 + 1);
            float acc = R[16 * 32 + 32 + 32];
            for (int j = 0; j < 32; ++j) {
                float hsum = R[16 * 32 + j];
#pragma unroll
                for (int i = 0; i < 16; ++i) hsum = fmaf(feat[i], R[i * 32 + j], hsum);
                acc = fmaf(fmaxf(hsum, 0.f), R[16 * 32 + 32 + j], acc);
            }
            outv = kind == 0 ? sigmoidf_(acc) : tanhf(acc);
        } else {
            const float* R = P + o.reg + L * head_reg_stride();
            float acc = R[16 * 16 + 16 + 16];
            for (int j = 0; j < 16; ++j) {
                float hsum = R[16 * 16 + j];
#pragma unroll
                for (int i = 0; i < 16; ++i) hsum = fmaf(feat[i], R[i * 16 + j], hsum);
                acc = fmaf(fmaxf(hsum, 0.f), R[16 * 16 + 16 + j], acc);
            }
            outv = sigmoidf_(acc);
        }
        float* dst = kind == 0 ? a.tau : (kind == 1 ? a.f : a.conf);
        dst[(size_t)(s0 + s) * L + t] = outv;
    }
}

}  // namespace admmnet
