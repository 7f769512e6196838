// k_dc: eigen-decomposition of the real symmetric tridiagonal matrix T = tridiag(e, d, e) (the middle of
// torch.linalg.eigh, admm_net.py:303) by divide & conquer, entirely in fp32 and entirely in shared memory: one CTA per
// signal replaces the k_ql + k_rotf pair (one-thread-per-signal fp64 QL chain + 8.5k plane rotations per matrix).
//
//   T is torn at EVERY off-diagonal (Cuppen): leaves are 1 x 1 (l_i = d_i - |e_{i-1}| - |e_i|, Q = I), and a binary
//   tree of ceil(log2 d) levels glues neighbouring blocks (the first level, 2 x 2 problems, in closed form).  One merge of the blocks [lo,p) and [p,hi):
//       diag(l) + rho z z^T,   rho = 2|e_{p-1}|,   z = (last row of Q1, sign(e_{p-1}) * first row of Q2) / sqrt(2)
//     - poles sorted by rank, deflation as LAPACK's xLAED2 (rho |z_i| <= tol: pair kept; (nearly) equal poles: Givens
//       rotation of the two columns - detected in parallel, carried out by a serial scan only when it occurs)
//     - one root of  -1/rho + sum_i z_i^2 / (l - d_i) = 0  per pole interval, a PAIR of lanes per root: osculatory
//       two-pole rational iteration with bracketing in coordinates shifted to the nearer pole (every l_j - d_i keeps
//       full relative accuracy in fp32), the scheme k_arrow uses for the layer-0 arrowhead
//     - Gu-Eisenstat: z re-derived from the computed roots, so the vectors w_j = (zhat_i / (d_i - l_j))_i are the
//       exact eigenvectors of a nearby problem: orthogonal to rounding for any pole spacing, no fp64 anywhere
//     - Q <- Q W on 4 x 4 register tiles (block diagonal: only coordinates lo..hi-1 are touched); W is formed on the
//       fly below the top level and materialised (in the free ping-pong buffer) at the top level, whose product goes
//       straight to global memory as Z^T [eigenvector][coordinate] with the row pitch k_tail_tc's TMA box expects.
//   CPU emulation of exactly this algorithm in float32: tests/dc_emulation.py (residual, orthogonality and eigenvalue
//   errors 1-6e-7 on random, clustered, graded, Wilkinson and degenerate matrices).
#include "common.cuh"

namespace admmnet {

constexpr int DCK_NT = 256;
constexpr int DCK_MAXIT = 48;
constexpr int DCP_NPH = 8;           // profiled phases per level (ADMMNET_DC_PROF=1): tables+z | sort+deflate+compact | close
                                     // poles | secular | Gu-Eisenstat | norms | W | GEMM+copy
constexpr int DCP_N = 128;           // counters: [level 1..7][phase] at 8*lev+ph, secular iterations / warps at 72+lev / 80+lev,
                                     // signals at 96

struct DcArgs {
    const float* dT;        // [d][B]
    const float* eT;        // [d][B]
    float* lam;             // [B][d]   eigenvalue of Z^T row c at lam[c] (unsorted)
    float* Zt;              // [B][d][ldz]
    int* status;
    const int* skip;
    long long* prof;        // tuning aid (nullptr: off)
    int B, d, ldz;
    // hybrid form: the lower levels were solved by k_ql + k_rotf on the 2^nhyb blocks of a torn T (TearSpec with
    // absconv = 1); Zt holds their block-diagonal Q^T, lam their eigenvalues, beta_in[sig][q] the torn off-diagonals at
    // rows tear_pos[q].  Only the top nhyb merge levels run here.  nhyb = 0: everything from 1 x 1 leaves.
    const double* beta_in;
    int nhyb, ntear, tear_stride;
    int tear_pos[7];
};
__host__ __device__ inline size_t dc_smem_bytes(int d, int ldz) {
    return (size_t)2 * d * ldz * sizeof(float) + (size_t)21 * 128 * sizeof(float) + 7 * 64 * sizeof(float) + 80 * sizeof(int);
}

// block of index i at a level with nb blocks: [r d / nb, (r+1) d / nb), split at (2r+1) d / (2 nb)
__device__ __forceinline__ int dc_blk(int i, int nb, int d) { return ((i + 1) * nb - 1) / d; }

// MUFU.RCP without __fdividef's range scaling (5 extra instructions per call): every denominator here is a difference of
// poles / roots of O(1e-7 .. 1e2) magnitude
__device__ __forceinline__ float dc_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// g(x) = a0 + sum_i z2_i / (x - (d_i - org)) by a pair of lanes (parity par), left sum over i < jsplit; returns the
// combined value, the slopes of the two sums and |psi| + |phi|
__device__ __forceinline__ void dc_eval(const float* __restrict__ sd, const float* __restrict__ sz2, int k, int jsplit,
                                        float org, float a0, float x, int par, float& g, float& wl, float& wr, float& sabs) {
    // four independent accumulation chains per sum: the loop is latency bound (LDS -> FADD -> MUFU -> FMUL per term)
    float p = 0.f, q = 0.f, f = 0.f, h = 0.f;
    int i = par;
    {
        float p1 = 0.f, q1 = 0.f, p2 = 0.f, q2 = 0.f, p3 = 0.f, q3 = 0.f;
        for (; i + 6 < jsplit; i += 8) {
            const float r0 = dc_rcp(x - (sd[i] - org)), r1 = dc_rcp(x - (sd[i + 2] - org));
            const float r2 = dc_rcp(x - (sd[i + 4] - org)), r3 = dc_rcp(x - (sd[i + 6] - org));
            const float t0 = sz2[i] * r0, t1 = sz2[i + 2] * r1, t2 = sz2[i + 4] * r2, t3 = sz2[i + 6] * r3;
            p += t0; p1 += t1; p2 += t2; p3 += t3;
            q = fmaf(t0, r0, q); q1 = fmaf(t1, r1, q1); q2 = fmaf(t2, r2, q2); q3 = fmaf(t3, r3, q3);
        }
        for (; i < jsplit; i += 2) {
            const float r = dc_rcp(x - (sd[i] - org));
            const float t = sz2[i] * r;
            p += t;
            q = fmaf(t, r, q);
        }
        p = (p + p1) + (p2 + p3);
        q = (q + q1) + (q2 + q3);
    }
    {
        float f1 = 0.f, h1 = 0.f, f2 = 0.f, h2 = 0.f, f3 = 0.f, h3 = 0.f;
        for (; i + 6 < k; i += 8) {
            const float r0 = dc_rcp(x - (sd[i] - org)), r1 = dc_rcp(x - (sd[i + 2] - org));
            const float r2 = dc_rcp(x - (sd[i + 4] - org)), r3 = dc_rcp(x - (sd[i + 6] - org));
            const float t0 = sz2[i] * r0, t1 = sz2[i + 2] * r1, t2 = sz2[i + 4] * r2, t3 = sz2[i + 6] * r3;
            f += t0; f1 += t1; f2 += t2; f3 += t3;
            h = fmaf(t0, r0, h); h1 = fmaf(t1, r1, h1); h2 = fmaf(t2, r2, h2); h3 = fmaf(t3, r3, h3);
        }
        for (; i < k; i += 2) {
            const float r = dc_rcp(x - (sd[i] - org));
            const float t = sz2[i] * r;
            f += t;
            h = fmaf(t, r, h);
        }
        f = (f + f1) + (f2 + f3);
        h = (h + h1) + (h2 + h3);
    }
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    q += __shfl_xor_sync(0xffffffffu, q, 1);
    f += __shfl_xor_sync(0xffffffffu, f, 1);
    h += __shfl_xor_sync(0xffffffffu, h, 1);
    g = a0 + (p + f);
    wl = -q;
    wr = -h;
    sabs = fabsf(p) + fabsf(f);
}

__global__ void __launch_bounds__(DCK_NT, 2) k_dc(DcArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int d = a.d, ldz = a.ldz;
    float* Qa = reinterpret_cast<float*>(smem_raw);          // [d][ldz]  Q^T: row = eigenvector, column = coordinate
    float* Qb = Qa + (size_t)d * ldz;
    float* lamv = Qb + (size_t)d * ldz;                       // [128] eigenvalue of column c
    float* zloc = lamv + 128;                                 // [128] z of column c; later: 1 = column rewritten by the GEMM
    float* es = zloc + 128;                                   // [128] off-diagonal e_i (between i and i+1)
    float* sd = es + 128;                                     // [128] sorted poles (per block, at lo + rank)
    float* sz = sd + 128;                                     // [128] sorted z
    float* ksd = sz + 128;                                    // [128] compact poles (at lo + u)
    float* ksz = ksd + 128;                                   // [128] compact z
    float* ksz2 = ksz + 128;                                  // [128] z^2
    float* orgv = ksz2 + 128;                                 // [128] origin pole value of root u
    float* mus = orgv + 128;                                  // [128] root offset from its origin
    float* nus = mus + 128;                                   // [128] 1 / ||w_u||
    float* zh = nus + 128;                                    // [128] Gu-Eisenstat z
    int* perm = reinterpret_cast<int*>(zh + 128);             // [128] sorted position -> column
    int* kflag = perm + 128;                                  // [128] kept (not deflated) by sorted position
    int* kcol = kflag + 128;                                  // [128] compact index -> column
    int* cpos = kcol + 128;                                   // [128] rotation records of the serial scan
    int* blkof = cpos + 128;                                  // [128] block of index i at this level
    float* csd = reinterpret_cast<float*>(blkof + 128);       // [128] column -> its pole (1e30 for a deflated column)
    float* czh = csd + 128;                                   // [128] column -> its Gu-Eisenstat z (0 for a deflated column)
    float* rho_r = czh + 256;                                 // [64] per block: signed off-diagonal (0: nothing to merge)
    int* kcnt = reinterpret_cast<int*>(rho_r + 64);           // [64] per block: non-deflated count
    int* serial = kcnt + 64;                                  // [64] per block: close poles found -> serial scan
    int* tb_lo = serial + 64;                                 // [64] block geometry of the level
    int* tb_p = tb_lo + 64;
    int* tb_hi = tb_p + 64;
    int* mixed = tb_hi + 64;                                  // [64] per block: a Givens rotation mixed columns of the two children
    int* toff = mixed + 64;                                   // [65] first GEMM tile of block rb (exclusive scan)
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int sig = blockIdx.x;
    if (a.skip && a.skip[sig]) return;
    float* Zg = a.Zt + (size_t)sig * d * ldz;

    const bool profiling = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long tlast = profiling ? clock64() : 0;
#define DC_MARK(LEV, PH)                                      \
    if (profiling) {                                          \
        const long long tnow = clock64();                     \
        atomicAdd((unsigned long long*)&a.prof[8 * (LEV) + (PH)], (unsigned long long)(tnow - tlast)); \
        tlast = tnow;                                         \
    }

    int nl = 0;
    while ((1 << nl) < d) ++nl;
    if (a.nhyb == 0) {
        // ---- leaves: every off-diagonal is a tear (leaf value d_i - |e_{i-1}| - |e_i|) ...
        if (tid < d) {
            const float di = a.dT[(size_t)tid * a.B + sig];
            const float e_own = tid < d - 1 ? a.eT[(size_t)tid * a.B + sig] : 0.f;
            const float e_prev = tid > 0 ? a.eT[(size_t)(tid - 1) * a.B + sig] : 0.f;
            lamv[tid] = di - fabsf(e_prev) - fabsf(e_own);
            es[tid] = e_own;
        }
        for (int idx = tid; idx < d * ldz / 4; idx += DCK_NT) {
            reinterpret_cast<float4*>(Qa)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
            reinterpret_cast<float4*>(Qb)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);   // a column's coordinates outside its own
        }                                                                            // block are read as zeros by every merge
        __syncthreads();
        // ... and the first merge level in closed form: its blocks have one or two rows, and the 2 x 2 problem
        //     [[l0 + |b|, b], [b, l1 + |b|]] is one Jacobi rotation (thread = block)
        const int nb1 = 1 << (nl - 1);
        if (tid < nb1) {
            const int lo = (tid * d) / nb1, hi = ((tid + 1) * d) / nb1;
            if (hi - lo == 2) {
                const float b = es[lo];
                float cs = 1.f, sn = 0.f;
                if (b != 0.f) {
                    const float aa = lamv[lo] + fabsf(b), cc = lamv[lo + 1] + fabsf(b);
                    const float theta = (cc - aa) / (2.f * b);
                    const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(fmaf(theta, theta, 1.f)));
                    cs = 1.f / sqrtf(fmaf(t, t, 1.f));
                    sn = t * cs;
                    lamv[lo] = aa - t * b;
                    lamv[lo + 1] = cc + t * b;
                }
                Qa[lo * ldz + lo] = cs;          Qa[lo * ldz + lo + 1] = -sn;       // row = eigenvector, column = coordinate
                Qa[(lo + 1) * ldz + lo] = sn;    Qa[(lo + 1) * ldz + lo + 1] = cs;
            } else if (hi - lo == 1) {
                Qa[lo * ldz + lo] = 1.f;
            }
        }
    } else {
        // ---- blocks solved by the QL pair: eigenvalues, block-diagonal Q^T and the torn off-diagonals from global memory
        if (tid < d) { lamv[tid] = a.lam[(size_t)sig * d + tid]; es[tid] = 0.f; }
        for (int idx = tid; idx < d * ldz / 4; idx += DCK_NT) {
            reinterpret_cast<float4*>(Qa)[idx] = reinterpret_cast<const float4*>(Zg)[idx];
            reinterpret_cast<float4*>(Qb)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncthreads();
        if (tid < a.ntear) {
            int tp = 1;                                          // (static indexing: a dynamically indexed kernel
#pragma unroll                                                   //  parameter array would be copied to local memory)
            for (int q = 0; q < 7; ++q) if (q == tid) tp = a.tear_pos[q];
            ADMM_ASSERT(tp >= 1 && tp < d);
            es[tp - 1] = (float)a.beta_in[(size_t)sig * a.tear_stride + tid];
        }
    }
    bool bad = false;
    __syncthreads();

    for (int lev = a.nhyb ? nl - a.nhyb + 1 : 2; lev <= nl; ++lev) {
        const int nb = 1 << (nl - lev);
        const bool top = lev == nl;
        // ---- P0: block tables of the level
        if (tid < 64) {
            int lo_ = 0, p_ = 0, hi_ = 0;
            float rr = 0.f;
            if (tid < nb) {
                lo_ = (tid * d) / nb; hi_ = ((tid + 1) * d) / nb; p_ = ((2 * tid + 1) * d) / (2 * nb);
                if (lo_ < p_ && p_ < hi_) rr = es[p_ - 1];
            }
            tb_lo[tid] = lo_; tb_p[tid] = p_; tb_hi[tid] = hi_; rho_r[tid] = rr; kcnt[tid] = 0; serial[tid] = 0; mixed[tid] = 0;
        }
        int r = 0;
        if (tid < d) { r = dc_blk(tid, nb, d); blkof[tid] = r; }
        __syncthreads();
        // ---- P1: z
        const int lo = tb_lo[r], hi = tb_hi[r], p = tb_p[r];
        const float beta = rho_r[r];
        const float rho = 2.f * fabsf(beta);
        const bool live = tid < d && rho > 0.f;               // rho_r != 0 implies lo < p < hi
        if (live) {
            const float zraw = tid < p ? Qa[tid * ldz + (p - 1)] : copysignf(1.f, beta) * Qa[tid * ldz + p];
            zloc[tid] = zraw * 0.70710678f;
        }
        __syncthreads();
        DC_MARK(lev, 0)
        // ---- P2: rank sort inside the block, deflation of negligible z (tolerance 8 eps max(|d|, |z|) over the
        //          block, as xLAED2), compaction of the kept poles - all from the thread's own two passes over the block
        float tol = 0.f;
        if (live) {
            const float di = lamv[tid], zi = zloc[tid];
            float dmax = 0.f, zmax = 0.f, dmax1 = 0.f, zmax1 = 0.f;
            int j = lo;
            for (; j + 1 < hi; j += 2) {
                dmax = fmaxf(dmax, fabsf(lamv[j])); dmax1 = fmaxf(dmax1, fabsf(lamv[j + 1]));
                zmax = fmaxf(zmax, fabsf(zloc[j])); zmax1 = fmaxf(zmax1, fabsf(zloc[j + 1]));
            }
            if (j < hi) { dmax = fmaxf(dmax, fabsf(lamv[j])); zmax = fmaxf(zmax, fabsf(zloc[j])); }
            tol = 8.f * 5.9604645e-8f * fmaxf(fmaxf(dmax, dmax1), fmaxf(zmax, zmax1));
            // rank among the poles of the block and compact index among the kept ones
            int rank = lo, ci = 0;
#pragma unroll 4
            for (int jj = lo; jj < hi; ++jj) {
                const float dj = lamv[jj];
                const int before = (dj < di || (dj == di && jj < tid)) ? 1 : 0;
                rank += before;
                ci += (before && rho * fabsf(zloc[jj]) > tol) ? 1 : 0;
            }
            const int keep = (rho * fabsf(zi) > tol) ? 1 : 0;
            ADMM_ASSERT(rank >= lo && rank < hi && lo + ci < hi && hi <= d && hi - lo <= 128);
            sd[rank] = di;
            sz[rank] = zi;
            perm[rank] = tid;
            kflag[rank] = keep;
            if (keep) {
                ksd[lo + ci] = di;
                ksz[lo + ci] = zi;
                ksz2[lo + ci] = zi * zi;
                kcol[lo + ci] = tid;
            }
            if (rank == hi - 1) kcnt[r] = ci + keep;
        }
        __syncthreads();
        DC_MARK(lev, 1)
        // ---- P5: (nearly) equal poles among the kept ones?  (thread = compact index)  Rare: the serial xLAED2 scan and
        //          the Givens rotations only run when some block of the level reports one
        {
            bool close = false;
            if (live) {
                const int u = tid - lo, k = kcnt[r];
                if (u + 1 < k) {
                    const float s_ = ksz[tid], c_ = ksz[tid + 1];
                    const float tau2 = c_ * c_ + s_ * s_;
                    const float tt = ksd[tid + 1] - ksd[tid];
                    close = fabsf(tt * c_ * s_) <= tol * tau2;
                    if (close) serial[r] = 1;
                }
            }
            if (__syncthreads_or(close)) {
                if (live && tid == lo && serial[r]) {
                    // (kflag already marks the negligible z; the previous kept pole and its z ride in registers, so an
                    //  iteration is two independent loads and, only when a rotation fires, a short dependent update)
                    int prev = -1, nrot = 0;
                    float zp = 0.f, dp = 0.f;
#pragma unroll 4
                    for (int t = lo; t < hi; ++t) {
                        float zt = sz[t], dt = sd[t];
                        if (!kflag[t]) continue;
                        if (prev >= 0) {
                            const float tau2 = fmaf(zt, zt, zp * zp);
                            if (fabsf((dt - dp) * zt * zp) <= tol * tau2) {   // |tt c s| <= tol with c, s = zt, -zp / tau
                                const float rt = rsqrtf(tau2);
                                const float c = zt * rt, s = -zp * rt;
                                // columns perm[prev], perm[t]: z[prev] -> 0, z[t] -> tau  (record; applied below)
                                // the record reuses the compact arrays of this block, which are rebuilt right after
                                ADMM_ASSERT(lo + nrot < hi && perm[prev] >= lo && perm[prev] < hi && perm[t] >= lo && perm[t] < hi);
                                ksd[lo + nrot] = c; ksz[lo + nrot] = s;
                                kcol[lo + nrot] = perm[prev]; cpos[lo + nrot] = perm[t];
                                ++nrot;
                                const float dpn = dp * c * c + dt * s * s;
                                dt = dp * s * s + dt * c * c;
                                zt = tau2 * rt;
                                sd[prev] = dpn; sz[prev] = 0.f; kflag[prev] = 0;
                                sd[t] = dt; sz[t] = zt;
                            }
                        }
                        prev = t; zp = zt; dp = dt;
                    }
                    serial[r] = nrot + 1;
                }
                __syncthreads();
                // apply the recorded rotations: a rotation mixes two columns coordinate by coordinate, so the thread
                // of coordinate x runs through its block's whole list on its own (no barrier between rotations)
                if (tid < d) {
                    const int rb = blkof[tid], nrot = serial[rb] - 1;
                    const int blo = tb_lo[rb];
                    for (int q = 0; q < nrot; ++q) {
                        const float c = ksd[blo + q], s = ksz[blo + q];
                        const int cp = kcol[blo + q], ct = cpos[blo + q];
                        ADMM_ASSERT(cp >= 0 && cp < d && ct >= 0 && ct < d && cp != ct);
                        const float qp = Qa[cp * ldz + tid], qt = Qa[ct * ldz + tid];
                        Qa[cp * ldz + tid] = c * qp + s * qt;
                        Qa[ct * ldz + tid] = -s * qp + c * qt;
                    }
                }
                __syncthreads();
                // rebuild the compact arrays of the scanned blocks
                if (live && serial[r] > 0) {
                    int ci = 0;
                    for (int t = lo; t < tid; ++t) ci += kflag[t];
                    if (kflag[tid]) {
                        ksd[lo + ci] = sd[tid];
                        ksz[lo + ci] = sz[tid];
                        ksz2[lo + ci] = sz[tid] * sz[tid];
                        kcol[lo + ci] = perm[tid];
                    }
                    // a Givens rotation may mix columns of the two children: the block's GEMM then takes every column
                    // of the block for every coordinate group
                    if (tid == hi - 1) { kcnt[r] = ci + kflag[tid]; mixed[r] = serial[r] > 1 ? 1 : 0; }
                }
                __syncthreads();
            }
        }
        DC_MARK(lev, 2)
        // ---- P6: secular roots, a pair of lanes per root (slot = tid / 2)
        {
            const int slot = tid >> 1, par = tid & 1;
            int slo = 0, k = 0, u = 0;
            float srho = 0.f;
            bool work = false;
            if (slot < d) {
                const int r2 = blkof[slot];
                ADMM_ASSERT(r2 >= 0 && r2 < nb && nb <= 64);
                slo = tb_lo[r2];
                srho = 2.f * fabsf(rho_r[r2]);
                k = kcnt[r2];
                u = slot - slo;
                work = srho > 0.f && u < k;
            }
            const float* pd = ksd + slo;
            const float* pz2 = ksz2 + slo;
            const int kk = work ? k : 0;
            const float a0 = work ? -1.f / srho : -1.f;
            const bool last = u == k - 1;
            float o = 0.f, x = 1.f, lo_ = 0.f, hi_ = 1.f, dL = 0.f, dR = 0.f;
            if (work) {
                if (last) {
                    float zs = 0.f;
                    for (int i = 0; i < k; ++i) zs += pz2[i];
                    o = pd[u];
                    lo_ = srho * pz2[u] * 0.9999f;
                    hi_ = srho * zs * 1.0001f + 1e-30f;
                    x = 0.5f * (lo_ + hi_);
                }
            }
            // first evaluation: at the middle of the pole interval (which half holds the root decides the origin: the
            // nearer pole; the same point in the new coordinates, so the values are reused), or at the middle of the
            // bracket of the last root
            float g, wl, wr, sa;
            {
                const bool mid = work && !last;
                const float gap = mid ? pd[u + 1] - pd[u] : 2.f;
                dc_eval(pd, pz2, kk, u + 1, mid ? pd[u] : o, a0, mid ? 0.5f * gap : x, par, g, wl, wr, sa);
                if (mid) {
                    if (g > 0.f) { o = pd[u + 1]; lo_ = -0.5f * gap; hi_ = 0.f; dL = -gap; dR = 0.f; x = lo_; }
                    else { o = pd[u]; lo_ = 0.f; hi_ = 0.5f * gap; dL = 0.f; dR = gap; x = hi_; }
                }
            }
            bool conv = !work;
            int its = 1;
            for (int it = 0; it < DCK_MAXIT; ++it) {
                if (!conv) {
                    if (g > 0.f) lo_ = x; else hi_ = x;
                    if (fabsf(g) <= 1.2e-7f * (8.f * sa + fabsf(a0))) conv = true;
                    else {
                        float eta;
                        if (last) {
                            const float w = wl + wr, D = x;
                            const float den = g + w * D;
                            eta = den != 0.f ? -g * D * dc_rcp(den) : 0.f;
                        } else {
                            const float DL = x - dL, DR = x - dR;
                            const float s = -wl * DL * DL, S = -wr * DR * DR;
                            const float Cc = g + wl * DL + wr * DR;
                            const float a1 = Cc * (DL + DR) + s + S, a0q = DL * DR * g;
                            const float disc = fmaxf(fmaf(a1, a1, -4.f * Cc * a0q), 0.f);
                            const float qq = a1 + copysignf(sqrtf(disc), a1);
                            eta = qq != 0.f ? -2.f * a0q * dc_rcp(qq) : 0.f;
                            const float xn0 = x + eta;
                            if (!(xn0 > lo_ && xn0 < hi_) && Cc != 0.f && eta != 0.f) eta = a0q * dc_rcp(Cc * eta);
                        }
                        float xn = x + eta;
                        if (!(xn > lo_ && xn < hi_)) xn = 0.5f * (lo_ + hi_);
                        if (xn == x || fabsf(xn - x) <= 6e-8f * fabsf(xn)) conv = true;
                        if (hi_ - lo_ <= 1.2e-7f * fmaxf(fabsf(lo_), fabsf(hi_))) conv = true;
                        x = xn;
                        if (it == DCK_MAXIT - 1 && !conv) bad = true;
                    }
                }
                if (__all_sync(0xffffffffu, conv)) break;
                ++its;
                dc_eval(pd, pz2, conv ? 0 : kk, u + 1, o, a0, x, par, g, wl, wr, sa);
            }
            if (work && !(x == x)) bad = true;
            if (work && par == 0) { orgv[slot] = o; mus[slot] = x; }
            if (a.prof != nullptr && blockIdx.x == 0 && lane == 0) {
                atomicAdd((unsigned long long*)&a.prof[72 + lev], (unsigned long long)its);
                atomicAdd((unsigned long long*)&a.prof[80 + lev], 1ull);
            }
        }
        __syncthreads();
        DC_MARK(lev, 3)
        // ---- P7: Gu-Eisenstat z (thread = compact pole i)
        //   zhat_i^2 = (l_{k-1} - d_i)/rho * prod_{j<i} (d_i - l_j)/(d_i - d_j) * prod_{j=i}^{k-2} (l_j - d_i)/(d_{j+1} - d_i)
        //   (a relative error of a few ulp per factor only moves the nearby problem the vectors are exact for)
        if (live) {
            const int u = tid - lo, k = kcnt[r];
            if (u < k) {
                const float* pd = ksd + lo;
                const float* po = orgv + lo;
                const float* pm = mus + lo;
                const float di = pd[u];
                float prod = ((po[k - 1] - di) + pm[k - 1]) / rho, prod1 = 1.f;
                int j = 0;
                for (; j + 1 < u; j += 2) {
                    prod *= ((di - po[j]) - pm[j]) * dc_rcp(di - pd[j]);
                    prod1 *= ((di - po[j + 1]) - pm[j + 1]) * dc_rcp(di - pd[j + 1]);
                }
                if (j < u) prod *= ((di - po[j]) - pm[j]) * dc_rcp(di - pd[j]);
                j = u;
                for (; j + 1 < k - 1; j += 2) {
                    prod *= ((po[j] - di) + pm[j]) * dc_rcp(pd[j + 1] - di);
                    prod1 *= ((po[j + 1] - di) + pm[j + 1]) * dc_rcp(pd[j + 2] - di);
                }
                if (j < k - 1) prod *= ((po[j] - di) + pm[j]) * dc_rcp(pd[j + 1] - di);
                zh[tid] = copysignf(sqrtf(fmaxf(prod * prod1, 0.f)), ksz[tid]);
            }
        } else if (wid == DCK_NT / 32 - 1) {
            // the last warp (never `live`: d <= 128 < 224) lays out the GEMM tiles of the level: block rb owns
            // ceil((hi - (lo & ~3)) / 4) coordinate groups x ceil(k / 4) root groups
            int cnt[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int rb = 2 * lane + h;
                const int k = rho_r[rb] != 0.f ? kcnt[rb] : 0;
                const int c0 = tb_lo[rb] & ~3;
                cnt[h] = ((tb_hi[rb] - c0 + 3) >> 2) * ((k + 3) >> 2);
            }
            const int own = cnt[0] + cnt[1];
            int incl = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            toff[2 * lane] = incl - own;
            toff[2 * lane + 1] = incl - own + cnt[0];
            if (lane == 31) toff[64] = incl;
        }
        __syncthreads();
        DC_MARK(lev, 4)
        // ---- P8: 1/||w_j||, new eigenvalue of the root's column (thread = compact root j); deflated columns keep
        //          theirs; zloc[column] := 1 when the GEMM rewrites the column
        if (tid < d) {
            if (live) {
                const int u = tid - lo, k = kcnt[r];
                if (u < k) {
                    const float* pd = ksd + lo;
                    const float* pz = zh + lo;
                    const float o = orgv[tid], x = mus[tid];
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                    int i = 0;
                    for (; i + 3 < k; i += 4) {
                        const float w0 = pz[i] * dc_rcp((pd[i] - o) - x), w1 = pz[i + 1] * dc_rcp((pd[i + 1] - o) - x);
                        const float w2 = pz[i + 2] * dc_rcp((pd[i + 2] - o) - x), w3 = pz[i + 3] * dc_rcp((pd[i + 3] - o) - x);
                        s0 = fmaf(w0, w0, s0);
                        s1 = fmaf(w1, w1, s1);
                        s2 = fmaf(w2, w2, s2);
                        s3 = fmaf(w3, w3, s3);
                    }
                    for (; i < k; ++i) { const float w0 = pz[i] * dc_rcp((pd[i] - o) - x); s0 = fmaf(w0, w0, s0); }
                    nus[tid] = rsqrtf((s0 + s1) + (s2 + s3));
                    lamv[kcol[tid]] = o + x;
                    // pole and z by COLUMN for the GEMM, which walks the columns of a child in their natural order
                    ADMM_ASSERT(kcol[tid] >= lo && kcol[tid] < hi && u < k && k <= hi - lo);
                    csd[kcol[tid]] = pd[u];
                    czh[kcol[tid]] = pz[u];
                }
                const int kf = kflag[tid];
                if (!kf) {                                           // deflated: (possibly rotated) pole, own column;
                    lamv[perm[tid]] = sd[tid];                       // its row of W is zero (1e30: no 0 * inf)
                    csd[perm[tid]] = 1e30f;
                    czh[perm[tid]] = 0.f;
                }
                zloc[perm[tid]] = kf ? 1.f : 0.f;
            } else {
                zloc[tid] = 0.f;
            }
        }
        __syncthreads();
        DC_MARK(lev, 5)
        // ---- P9: Q <- Q W.  Output row = column kcol[j] of the block, coordinates [lo, hi).
        float* out = top ? Zg : Qb;
        if (top) {
            // materialise W in Qb: row = column of Q (its pole), column j = root: zhat nu_j / (d - l_j), pitch ldz
            if (rho_r[0] != 0.f) {
                const int k = kcnt[0], kp = (k + 3) & ~3;
                float oj[4], mj[4], nj[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    const int j = lane + 32 * x;
                    const bool okj = j < k;
                    oj[x] = okj ? orgv[j] : 0.f;
                    mj[x] = okj ? mus[j] : 1.f;
                    nj[x] = okj ? nus[j] : 0.f;
                }
                for (int c = wid; c < d; c += DCK_NT / 32) {         // row = COLUMN c of Q (zero for a deflated one)
                    const float zi = czh[c], di = csd[c];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const int j = lane + 32 * x;
                        if (j < kp) Qb[c * ldz + j] = j < k ? zi * nj[x] * dc_rcp((di - oj[x]) - mj[x]) : 0.f;
                    }
                }
            }
            __syncthreads();
            DC_MARK(lev, 6)
        }
        {
            const int total = toff[64];
            for (int t = tid; t < total; t += DCK_NT) {
                int rb = 0;
#pragma unroll
                for (int s = 32; s > 0; s >>= 1) if (toff[rb + s] <= t) rb += s;
                const int blo = tb_lo[rb], bhi = tb_hi[rb];
                const int k = kcnt[rb];
                const int c0 = blo & ~3;
                const int nct = (bhi - c0 + 3) >> 2;
                const int tile = t - toff[rb];
                // side-major order - groups inside the first child, inside the second, then the (at most one per root
                // group) that straddles the split: the tiles of a warp run the same number of iterations
                const int bp = tb_p[rb], mx = mixed[rb];
                const int njt = (k + 3) >> 2;
                const int n1 = mx ? 0 : (bp - c0) >> 2, ns = mx ? nct : (((bp - c0) & 3) ? 1 : 0), n2 = nct - n1 - ns;
                int jt, ct;
                if (tile < n1 * njt) { jt = tile / n1; ct = tile - jt * n1; }
                else if (tile < (n1 + n2) * njt) { const int t2 = tile - n1 * njt; jt = t2 / n2; ct = n1 + ns + (t2 - jt * n2); }
                else { const int t3 = tile - (n1 + n2) * njt; jt = t3 / ns; ct = n1 + (t3 - jt * ns); }
                const int j0 = 4 * jt, cc = c0 + 4 * ct;
                ADMM_ASSERT(rb < nb && tile >= 0 && tile < nct * njt && jt >= 0 && jt < njt && ct >= 0 && ct < nct);
                ADMM_ASSERT(cc >= 0 && cc + 3 < ldz && j0 + 3 < ldz && k > 0 && k <= bhi - blo);
                float acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.f;
                // a column of the first child is zero on the second child's coordinates and vice versa: only the columns
                // of the tile's own child contribute (a coordinate group that straddles the split takes both); the
                // columns are walked in natural order - plain strided addresses, four iterations of loads in flight
                const int cbeg = (!mx && cc >= bp) ? bp : blo, cend = (!mx && cc + 3 < bp) ? bp : bhi;
                ADMM_ASSERT(cbeg >= blo && cend <= bhi && cbeg <= cend && bhi <= d);
                if (top) {
#pragma unroll 4
                    for (int c = cbeg; c < cend; ++c) {
                        const float4 qv = *reinterpret_cast<const float4*>(Qa + c * ldz + cc);
                        const float4 wv = *reinterpret_cast<const float4*>(Qb + c * ldz + j0);
                        acc[0][0] = fmaf(wv.x, qv.x, acc[0][0]); acc[0][1] = fmaf(wv.x, qv.y, acc[0][1]);
                        acc[0][2] = fmaf(wv.x, qv.z, acc[0][2]); acc[0][3] = fmaf(wv.x, qv.w, acc[0][3]);
                        acc[1][0] = fmaf(wv.y, qv.x, acc[1][0]); acc[1][1] = fmaf(wv.y, qv.y, acc[1][1]);
                        acc[1][2] = fmaf(wv.y, qv.z, acc[1][2]); acc[1][3] = fmaf(wv.y, qv.w, acc[1][3]);
                        acc[2][0] = fmaf(wv.z, qv.x, acc[2][0]); acc[2][1] = fmaf(wv.z, qv.y, acc[2][1]);
                        acc[2][2] = fmaf(wv.z, qv.z, acc[2][2]); acc[2][3] = fmaf(wv.z, qv.w, acc[2][3]);
                        acc[3][0] = fmaf(wv.w, qv.x, acc[3][0]); acc[3][1] = fmaf(wv.w, qv.y, acc[3][1]);
                        acc[3][2] = fmaf(wv.w, qv.z, acc[3][2]); acc[3][3] = fmaf(wv.w, qv.w, acc[3][3]);
                    }
                } else {
                    float oj[4], mj[4], nj[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) {
                        const bool okj = j0 + x < k;
                        oj[x] = okj ? orgv[blo + j0 + x] : 0.f;
                        mj[x] = okj ? mus[blo + j0 + x] : 1.f;
                        nj[x] = okj ? nus[blo + j0 + x] : 0.f;
                    }
#pragma unroll 4
                    for (int c = cbeg; c < cend; ++c) {
                        const float4 qv = *reinterpret_cast<const float4*>(Qa + c * ldz + cc);
                        const float di = csd[c], zi = czh[c];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const float w = zi * nj[x] * dc_rcp((di - oj[x]) - mj[x]);
                            acc[x][0] = fmaf(w, qv.x, acc[x][0]); acc[x][1] = fmaf(w, qv.y, acc[x][1]);
                            acc[x][2] = fmaf(w, qv.z, acc[x][2]); acc[x][3] = fmaf(w, qv.w, acc[x][3]);
                        }
                    }
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (j0 + x >= k) continue;
                    ADMM_ASSERT(kcol[blo + j0 + x] >= blo && kcol[blo + j0 + x] < bhi);
                    float* orow = out + (size_t)kcol[blo + j0 + x] * ldz;
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int c = cc + y;
                        if (c >= blo && c < bhi) orow[c] = acc[x][y];
                    }
                }
            }
        }
        // deflated columns and the columns of blocks without a merge are copied (block part only; at the top level
        // the whole row goes to global memory, and the pad columns of the TMA box are zeroed for every row)
        for (int c = wid; c < d; c += DCK_NT / 32) {
            const bool rewritten = zloc[c] != 0.f;
            if (top) {
                if (!rewritten) for (int x = lane; x < ldz; x += 32) Zg[(size_t)c * ldz + x] = x < d ? Qa[c * ldz + x] : 0.f;
                else if (lane < ldz - d) Zg[(size_t)c * ldz + d + lane] = 0.f;
            } else if (!rewritten) {
                const int rc = blkof[c];
                for (int x = tb_lo[rc] + lane; x < tb_hi[rc]; x += 32) Qb[c * ldz + x] = Qa[c * ldz + x];
            }
        }
        __syncthreads();
        DC_MARK(lev, 7)
        { float* t = Qa; Qa = Qb; Qb = t; }
    }
    if (tid < d) a.lam[(size_t)sig * d + tid] = lamv[tid];
    if (profiling) atomicAdd((unsigned long long*)&a.prof[96], 1ull);
    if (__syncthreads_or(bad) && tid == 0) atomicOr(a.status, 8);
#undef DC_MARK
}

}  // namespace admmnet
