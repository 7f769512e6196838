// k_tail_tc: the tensor-core form of k_tail_p (layers >= 1 of admmnet_forward, n = 100 class sizes, d <= 104).
//
//   U = Q_H Z  (back-transformation of the real eigenvectors of T, reference: second half of torch.linalg.eigh,
//               admm_net.py:303),   l' = f(l) (admm_net.py:310-334),
//   G = U diag(l') U^H (admm_net.py:343-352),   r = ||G - C||_F (admm_net.py:454)
//
// Both contractions run on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators and the running
// eigenvector matrix in tensor memory), at fp32 accuracy through the 3xTF32 split a*b ~ ah*bh + ah*bl + al*bh
// (ah = rna_tf32(a), al = a - ah).  Z^T arrives by TMA (cp.async.bulk.tensor.2d), one box per signal.
//
// Data layout.  X = U^T lives in TMEM: lane i = eigenvector i, columns = coordinates; the real plane occupies columns
// a (0..dp-1), the imaginary plane is stored column-REVERSED at 2dp-1-a, so the coordinates a >= a0 touched by a block
// of reflectors form ONE contiguous window [a0, 2dp-a0).  X is kept as an exact pair (hi, lo): hi = rna_tf32(x) in
// columns [0, 2dp), lo = x - hi in [2dp, 4dp); updates are accumulated into lo by the MMA (fp32 accumulator), then a
// re-split pass restores the invariant.  P (the block's projections) sits in [4dp, 4dp+96).
//
// Back-transformation, reflectors k = 0..d-2 (H_k = I - tau_k v_k v_k^H, U <- H_0 H_1 ... H_{d-2} Z):
//   * the trailing `rem` <= 8 reflectors only touch the last rem coordinates: applied per thread on its own row;
//   * the others in blocks of NB = 24 from the last block to the first, compact WY:  Q_j = I - V_j T_j V_j^H,
//         X <- X - P Y^T,   P = X conj(V_j)   (GEMM 1: M=128, N=48 [Pr|Pi], K = window),
//                           Y = V_j T_j       (GEMM 2: M=128, N = window, K = 24, for A = Pr and A = Pi).
//     Y is obtained without forming T: row a of Y solves  y (diag(1/tau) + striu(S)) = v_a,  S = V_j^H V_j, one
//     thread per (block, coordinate) - 256 independent substitutions for d = 101.
// Rebuild: W = sqrt(l')_i X (l' > 0 always, admm_net.py:321-331), staged K-major [coordinate][eigen index] in shared
// memory; G_r = Wr Wr^T + Wi Wi^T, G_i = Wi Wr^T - Wr Wi^T : 4 products x 3 split terms x dp/8 K-steps of M=128,
// N=dp MMAs into TMEM columns [0, 2dp).  Epilogue: lower triangle -> residual norm, packed store through shared
// memory (coalesced).
//
// Persistent: one CTA per SM (222 KB of shared memory, all 512 TMEM columns); the next signal's reflectors and small
// vectors stream in with cp.async while the current one is rebuilt, its Z^T by TMA during the epilogue.
//
// Overlap inside a signal: MMAs are asynchronous, so while GEMM 1 of block j runs the threads build the Y tiles of
// block j (buffer B), and while GEMM 2 runs they build the V tiles of block j-1 (buffer A).  The MMAs of this kernel
// are operand-bandwidth bound (tf32 K = 8: every MMA reads a 4 KB A slab from TMEM at 64 B/clk), so GEMM 1 reads the
// X window only twice: A = X_hi against the 96-row tile [V_hi ; V_lo] (N = 96) and A = X_lo against its first 48 rows.
#include "common.cuh"
#include "tc.cuh"

namespace admmnet {

constexpr int TC_NB = 24;            // reflectors per block (P occupies 4*NB = 96 TMEM columns)
constexpr int TC_NT = 256;           // 8 worker warps: warp w works on TMEM lanes 32*(w&3).., column groups of parity w>>2
constexpr int TC_THREADS = TC_NT + 32;   // + one warp whose first lane only issues the MMAs (tcgen05.mma issue blocks while
                                         //   the tensor pipe's queue is full: a worker doing it would stall its warp's share)
constexpr int TC_MAXBLK = 5;
constexpr int TC_LOCAL_MAX = 8;

struct TailTcPlan {
    int d, dp, ldz;                  // matrix order, d rounded up to 8, row pitch of Z^T
    int nblk, nloc;                  // WY blocks, trailing reflectors applied per thread
    int k0[TC_MAXBLK], nb[TC_MAXBLK], a0[TC_MAXBLK];   // first reflector, count, window start of each block
    int yoff[TC_MAXBLK];             // byte offset (from the shared-memory base) of the block's rows of Y = V T
    int yrow0[TC_MAXBLK];            // first work item (row of Y) of the block (block 0 comes last), yrows = total
    int yrows;
    // shared-memory byte offsets: buffer A (V tiles) | buffer B (Y tiles; Z^T box early) | S (Gram; later Y rows of block 0)
    int off_a, off_b, off_s, off_z, off_vs, off_small, total;
    int lbo_w;                       // K-chunk stride of a W tile (odd multiple of 16 bytes)
};

// Shared-memory map (d = 101: 227 072 bytes).  R0 = [A | B | S] is also the home of the four W tiles of the rebuild and
// of the packed-G staging; the rows of Y of the blocks live where the tiles of their block do not reach:
//   block 0 -> S (after the Gram matrices have been consumed), block 1 -> tail of A, blocks 2, 3 -> tail of B.
inline __host__ bool tail_tc_plan(int d, TailTcPlan& p) {
    if (d < 33 || d > 104) return false;
    p.d = d; p.dp = (d + 7) & ~7; p.ldz = 4 * ((d + 3) / 4);
    const int nref = d - 1;
    int nfull = nref / TC_NB, rem = nref - nfull * TC_NB;
    p.nblk = nfull; p.nloc = rem;
    if (rem > TC_LOCAL_MAX) { p.nblk = nfull + 1; p.nloc = 0; }
    if (p.nblk < 1 || p.nblk > 4) return false;
    auto al = [](int x) { return (x + 127) & ~127; };
    int ybytes[TC_MAXBLK] = {0, 0, 0, 0, 0}, tbytes[TC_MAXBLK] = {0, 0, 0, 0, 0};
    p.yrows = 0;
    for (int jj = 0; jj < p.nblk; ++jj) {
        // work-item order of the rows of Y: blocks 1..nblk-1 first, block 0 LAST (its rows are stored over the Gram
        // matrices, after every reader is done; a thread's block-0 row is therefore always its last item)
        const int j = (jj + 1) % p.nblk;
        p.k0[j] = j * TC_NB;
        p.nb[j] = (j < nfull) ? TC_NB : rem;
        p.a0[j] = (p.k0[j] + 1) & ~7;
        p.yrow0[j] = p.yrows;
        p.yrows += d - 1 - p.k0[j];
        ybytes[j] = (d - 1 - p.k0[j]) * TC_NB * 8;
        tbytes[j] = 96 * 2 * (p.dp - p.a0[j]) * 4;              // V tile (96 rows) = the four Y tiles of the block
    }
    // more rows than threads: the second-round items must all be block-0 rows, given to threads whose first item was not
    if (p.yrows > TC_NT && (p.yrows > 2 * TC_NT || p.yrow0[0] > TC_NT || p.yrows - TC_NT > p.yrow0[0])) return false;
    const int szT = al(tbytes[0]);
    const int szS = al(ybytes[0] > p.nblk * TC_NB * TC_NB * 8 ? ybytes[0] : p.nblk * TC_NB * TC_NB * 8);
    p.lbo_w = 16 * (p.dp | 1);
    const int wbytes = 4 * (p.dp / 4) * p.lbo_w;                // 4 W tiles
    p.off_a = 1024;
    p.off_b = p.off_a + szT;
    p.off_s = p.off_b + szT;
    p.off_z = p.off_b;                                          // Z^T box: head of buffer B (consumed before any Y tile)
    int r0_end = p.off_s + szS;
    if (r0_end < p.off_a + al(wbytes)) r0_end = p.off_a + al(wbytes);
    // rows of Y
    p.yoff[0] = p.off_s;
    if (p.nblk > 1) {
        p.yoff[1] = p.off_a + szT - al(ybytes[1]);
        if (p.yoff[1] < p.off_a + tbytes[1]) return false;      // must clear the V tile of block 1
    }
    int tailb = p.off_b + szT;
    for (int j = 2; j < p.nblk; ++j) {
        tailb -= al(ybytes[j]);
        p.yoff[j] = tailb;
    }
    if (p.nblk > 2) {
        if (tailb < p.off_b + tbytes[2]) return false;          // must clear the Y tiles of blocks 2, 3 ...
        if (tailb < p.off_z + d * p.ldz * 4) return false;      // ... and the Z^T box
    }
    const int gstage = d * (d + 1) / 2 * 8;
    if (p.off_a + gstage > p.off_z) return false;               // packed-G staging must not overlap the next Z^T box
    p.off_vs = r0_end;
    p.off_small = al(p.off_vs + d * (d - 1) / 2 * 8);
    p.total = p.off_small + 2 * 128 * 8 /*tau*/ + 2 * 128 * 4 /*lam*/ + 2 * 128 * 8 /*phi*/ + 2 * 128 * 4 /*h*/ + 512;
    return p.total <= 227 * 1024;
}

struct TailTcArgs {
    const float* Zr;        // [B][d][ldz]
    float2* GV;             // [B][npk]: reflectors on entry, G packed lower on exit
    const float2* tau;      // [B][d]
    const float* lam;       // [B][d]
    const float2* phi_cur;  // [B][n]
    const float* h_cur;     // [B][n]
    const float* Pk;
    float* r_out;           // [B]
    int B, n, with_c;       // with_c = 0: plain f(A) (debug tap), no residual
    long long* prof;        // optional [16]: clock cycles per phase, summed over the signals of CTA 0 (tuning aid)
    TailTcPlan plan;
};
enum TcPhase { TCP_WAIT = 0, TCP_S1, TCP_GRAM, TCP_YSOLVE, TCP_VTILE, TCP_GEMM1, TCP_PSPLIT_YTILE, TCP_GEMM2, TCP_RESPLIT,
               TCP_WSTAGE, TCP_REBUILD, TCP_EPILOGUE, TCP_SIGNALS, TCP_NPHASE };

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_tail_tc(TailTcArgs a, const __grid_constant__ CUtensorMap tmZ) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const TailTcPlan& pl = a.plan;
    const int d = pl.d, dp = pl.dp, ldz = pl.ldz, n = a.n;
    const int DP2 = 2 * dp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool worker = warp < TC_NT / 32, issuer = tid == TC_NT;
    const int q = warp & 3, hh = (warp >> 2) & 1;
    const int row = 32 * q + lane;                       // TMEM lane = eigenvector index (phases on X), row of G (epilogue)
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;

    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_z = bar_mma + 1;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 16);
    float* red = reinterpret_cast<float*>(smem + 64);    // [32]
    float* bufA = reinterpret_cast<float*>(smem + pl.off_a);           // V tiles (96 rows x window, K-major)
    float* bufB = reinterpret_cast<float*>(smem + pl.off_b);           // Y tiles of the block
    float2* Sg = reinterpret_cast<float2*>(smem + pl.off_s);           // Gram matrices [blk][NB][NB]
    float* Zt = reinterpret_cast<float*>(smem + pl.off_z);             // [d][ldz] TMA box
    float* epar = reinterpret_cast<float*>(smem + 256);                // [64] eigenvalue-map parameters
    float2* Vs = reinterpret_cast<float2*>(smem + pl.off_vs);          // packed reflectors
    float2* taub = reinterpret_cast<float2*>(smem + pl.off_small);     // [2][128]
    float2* phib = taub + 2 * 128;                                     // [2][128]
    float* lamb = reinterpret_cast<float*>(phib + 2 * 128);            // [2][128]
    float* hb = lamb + 2 * 128;                                        // [2][128]
    float2* Gs = reinterpret_cast<float2*>(smem + pl.off_a);           // packed G staging (alias, late)

    const int npk = d * (d + 1) / 2, nv = d * (d - 1) / 2;
    const float* __restrict__ P = a.Pk;
    const float c1z = (a.with_c > 0) ? P[P_C1Z] : 0.f;

    if (tid == 0) {
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_z, 1);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmZ);
    }
    if (warp == 0) tc::tmem_alloc(tslot, 512);
    if (tid < 64) epar[tid] = P[tid];                    // P_THR, P_VC2, P_V1, P_VC1, P_V2 all lie below index 64
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = *tslot;
    const uint32_t tXhi = tbase, tXlo = tbase + DP2, tP = tbase + 2 * DP2;
    uint32_t ph_mma = 0, ph_z = 0;

    auto prefetch_small = [&](int sig, int buf) {
        if (!worker) return;
        for (int i = tid; i < d; i += TC_NT) {
            cp_async8(taub + buf * 128 + i, a.tau + (size_t)sig * d + i);
            cp_async4(lamb + buf * 128 + i, a.lam + (size_t)sig * d + i);
        }
        if (a.with_c > 0)
            for (int j = tid; j < n; j += TC_NT) {
                cp_async8(phib + buf * 128 + j, a.phi_cur + (size_t)sig * n + j);
                cp_async4(hb + buf * 128 + j, a.h_cur + (size_t)sig * n + j);
            }
    };
    auto prefetch_vs = [&](int sig) {
        if (!worker) return;
        const float2* gv = a.GV + (size_t)sig * npk;
        for (int idx = tid; idx < nv; idx += TC_NT) cp_async8(Vs + idx, gv + idx);
    };
    auto load_z = [&](int sig) {                          // one thread: TMA box [d][ldz] of signal sig
        tc::fence_async_smem();
        tc::mbar_expect_tx(bar_z, (uint32_t)(d * ldz * sizeof(float)));
        tc::tma_load_2d(Zt, &tmZ, 0, sig * d, bar_z);
    };
    // wait for all MMAs committed so far; every thread observes every phase
    auto wait_mma = [&]() {
        tc::mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1;
        tc::tc_fence_after_sync();
    };
    // TMEM / shared-memory writes of all threads -> visible to the MMAs the issuing thread is about to launch
    auto publish = [&]() {
        tc::tmem_wait_st();
        tc::tmem_wait_ld();
        tc::fence_async_smem();
        tc::tc_fence_before_sync();
        __syncthreads();
        tc::tc_fence_after_sync();
    };

    long long pacc[TCP_NPHASE];
#pragma unroll
    for (int i = 0; i < TCP_NPHASE; ++i) pacc[i] = 0;
    const bool profiling = a.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long tlast = profiling ? clock64() : 0;
#define TC_MARK(PH)                                 \
    if (profiling) {                                \
        const long long tnow = clock64();           \
        pacc[PH] += tnow - tlast;                   \
        tlast = tnow;                               \
    }

    int buf = 0;
    if ((int)blockIdx.x < a.B) {
        prefetch_small(blockIdx.x, 0);
        prefetch_vs(blockIdx.x);
        cp_async_commit();
        if (tid == 0) load_z(blockIdx.x);
    }
    for (int sig = blockIdx.x; sig < a.B; sig += gridDim.x, buf ^= 1) {
        const float2* taus = taub + buf * 128;
        const float2* phis = phib + buf * 128;
        const float* hs = hb + buf * 128;
        const bool has_next = sig + (int)gridDim.x < a.B;
        cp_async_wait<0>();
        tc::mbar_wait(bar_z, ph_z);
        ph_z ^= 1;
        __syncthreads();
        TC_MARK(TCP_WAIT)

        // ================= S1: X <- Z^T (trailing reflectors applied per row), split into TMEM =================
        float sl = 0.f;                                                   // sqrt(l'_row)
        if (row < d) {
            const float l = lamb[buf * 128 + row];
            sl = sqrtf(eig_map(epar, l));
        }
        if (worker) {
            const int aL = d - pl.nloc;                                   // first coordinate touched locally
            float ur[TC_LOCAL_MAX], ui[TC_LOCAL_MAX];
#pragma unroll
            for (int e = 0; e < TC_LOCAL_MAX; ++e) {
                ur[e] = (row < d && e < pl.nloc) ? Zt[row * ldz + aL + e] : 0.f;
                ui[e] = 0.f;
            }
            // reflectors k = d-2 down to d-1-nloc; v_k lives on coordinates k+1.. = local index e >= k+1-aL
            for (int k = d - 2; k >= d - 1 - pl.nloc; --k) {
                const float2 tk = taus[k];
                const int vb = voff(k, d) - (k + 1);
                float dx = 0.f, dy = 0.f;                                 // v^H u
#pragma unroll
                for (int e = 0; e < TC_LOCAL_MAX; ++e) {
                    const int r = aL + e;
                    if (r > k && r < d) {
                        const float2 v = Vs[vb + r];
                        dx += v.x * ur[e] + v.y * ui[e];
                        dy += v.x * ui[e] - v.y * ur[e];
                    }
                }
                const float2 t = cmul(tk, make_float2(dx, dy));
#pragma unroll
                for (int e = 0; e < TC_LOCAL_MAX; ++e) {
                    const int r = aL + e;
                    if (r > k && r < d) {
                        const float2 v = Vs[vb + r];
                        ur[e] -= v.x * t.x - v.y * t.y;
                        ui[e] -= v.x * t.y + v.y * t.x;
                    }
                }
            }
            // column groups of 8: real plane g -> columns 8g.., imaginary plane -> reversed columns.  Groups that do not
            // reach the locally transformed coordinates take the fast path (two 128-bit loads / plain zeros).
            const int ngr = dp / 8;
            for (int g = hh; g < 2 * ngr; g += 2) {
                uint32_t vh[8], vl[8];
                const bool im = g >= ngr;
                const int amax = im ? (DP2 - 1 - 8 * g) : (8 * g + 7);        // largest coordinate in the group
                if (pl.nloc == 0 || amax < aL) {
                    if (im) {
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) { vh[jj] = 0u; vl[jj] = 0u; }
                    } else {
                        float x[8];
                        if (row < d) {
                            const float4 z0 = *reinterpret_cast<const float4*>(Zt + row * ldz + 8 * g);
                            const float4 z1 = (8 * g + 4 < ldz) ? *reinterpret_cast<const float4*>(Zt + row * ldz + 8 * g + 4)
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
                            x[0] = z0.x; x[1] = z0.y; x[2] = z0.z; x[3] = z0.w;
                            x[4] = z1.x; x[5] = z1.y; x[6] = z1.z; x[7] = z1.w;
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) x[jj] = 0.f;
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float xv = (8 * g + jj < d) ? x[jj] : 0.f;   // pad columns of the box are zero anyway
                            const float h = rna_tf32(xv);
                            vh[jj] = __float_as_uint(h);
                            vl[jj] = __float_as_uint(xv - h);
                        }
                    }
                } else {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int acol = im ? (DP2 - 1 - (8 * g + jj)) : (8 * g + jj);      // coordinate of TMEM column 8g+jj
                        float x = 0.f;
                        if (row < d && acol < d) {
                            if (acol >= aL) {
                                // local-reflector coordinates: registers (static index through the unrolled select)
                                float xr = 0.f, xi = 0.f;
#pragma unroll
                                for (int e = 0; e < TC_LOCAL_MAX; ++e)
                                    if (acol - aL == e) { xr = ur[e]; xi = ui[e]; }
                                x = im ? xi : xr;
                            } else if (!im) {
                                x = Zt[row * ldz + acol];
                            }
                        }
                        const float h = rna_tf32(x);
                        vh[jj] = __float_as_uint(h);
                        vl[jj] = __float_as_uint(x - h);
                    }
                }
                tc::tmem_st8(tXhi + lane_off + 8 * g, vh);
                tc::tmem_st8(tXlo + lane_off + 8 * g, vl);
            }
        }

        TC_MARK(TCP_S1)
        // ================= S2: Gram matrices and Y = V T for all blocks (SIMT) =================
        // S_j[c1][c2] = v_c1^H v_c2 for c1 < c2.  One item = (block, pair of rows c1 = 2p, 2p+1, strip of four columns
        // c2 = 4s..4s+3) with 4s+3 > 2p: 42 items per block, each a 2x4 register tile (6 loads for 8 complex MACs per
        // coordinate), packed FFMA2:  conj(a) b = a.x*(b.x,b.y) + a.y*(b.y,-b.x)  kept as P += a.x*b, Q += a.y*b.
        // Four adjacent lanes share an item (coordinates r = r0 + lane4, +4, ...) and combine with two shuffle steps.
        if (worker) {
            const int l4 = lane & 3;
            const int nitem = pl.nblk * 42;
            for (int it0 = 0; it0 < nitem; it0 += TC_NT / 4) {
                const int item = it0 + (tid >> 2);
                const bool on = item < nitem;
                f32x2 Pa[2][4], Qa[2][4];
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) { Pa[t][u] = pk2(0.f, 0.f); Qa[t][u] = pk2(0.f, 0.f); }
                int j = 0, c1 = 0, c2 = 0, nbj = 0;
                if (on) {
                    j = item / 42;
                    int rem_i = item - j * 42, p2 = 0;
                    for (;; ++p2) {
                        const int cnt = 6 - ((2 * p2 + 1) >> 2);
                        if (rem_i < cnt) break;
                        rem_i -= cnt;
                    }
                    const int s4 = ((2 * p2 + 1) >> 2) + rem_i;                // strip index
                    const int k0 = pl.k0[j];
                    nbj = pl.nb[j];
                    c1 = 2 * p2; c2 = 4 * s4;
                    const float2* va0 = Vs + voff(k0 + c1, d) - (k0 + c1 + 1);
                    const float2* va1 = Vs + voff(k0 + c1 + 1, d) - (k0 + c1 + 2);
                    const float2* vb[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) vb[u] = Vs + voff(k0 + c2 + u, d) - (k0 + c2 + u + 1);
                    const int rfull = k0 + c2 + 4;                             // from here on all four columns are stored
                    if (c2 + 3 < nbj && rfull <= d) {
                        // the three leading rows (column c2+u exists for r >= k0+c2+u+1): lane 0 of the group
                        if (l4 == 0) {
#pragma unroll
                            for (int e = 0; e < 3; ++e) {
                                const int r = k0 + c2 + 1 + e;
                                if (r < d) {
                                    const float2 a0 = (r > k0 + c1) ? va0[r] : make_float2(0.f, 0.f);
                                    const float2 a1 = (r > k0 + c1 + 1) ? va1[r] : make_float2(0.f, 0.f);
#pragma unroll
                                    for (int u = 0; u < 3; ++u) {
                                        if (u <= e) {
                                            const float2 b = vb[u][r];
                                            const f32x2 bp = pk2(b.x, b.y);
                                            Pa[0][u] = fma2(bc2(a0.x), bp, Pa[0][u]); Qa[0][u] = fma2(bc2(a0.y), bp, Qa[0][u]);
                                            Pa[1][u] = fma2(bc2(a1.x), bp, Pa[1][u]); Qa[1][u] = fma2(bc2(a1.y), bp, Qa[1][u]);
                                        }
                                    }
                                }
                            }
                        }
                        for (int r = rfull + l4; r < d; r += 4) {
                            const float2 a0 = va0[r], a1 = va1[r];
                            const f32x2 a0x = bc2(a0.x), a0y = bc2(a0.y), a1x = bc2(a1.x), a1y = bc2(a1.y);
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float2 b = vb[u][r];
                                const f32x2 bp = pk2(b.x, b.y);
                                Pa[0][u] = fma2(a0x, bp, Pa[0][u]); Qa[0][u] = fma2(a0y, bp, Qa[0][u]);
                                Pa[1][u] = fma2(a1x, bp, Pa[1][u]); Qa[1][u] = fma2(a1y, bp, Qa[1][u]);
                            }
                        }
                    } else {
                        for (int r = k0 + c2 + 1 + l4; r < d; r += 4) {
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                const int ca = c1 + t;
                                if (ca >= nbj || r <= k0 + ca) continue;
                                const float2 av = (t ? va1 : va0)[r];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (c2 + u >= nbj || r <= k0 + c2 + u) continue;
                                    const float2 b = vb[u][r];
                                    const f32x2 bp = pk2(b.x, b.y);
                                    Pa[t][u] = fma2(bc2(av.x), bp, Pa[t][u]); Qa[t][u] = fma2(bc2(av.y), bp, Qa[t][u]);
                                }
                            }
                        }
                    }
                }
                // combine the four lanes of the group (all 32 lanes take part in the shuffles)
#pragma unroll
                for (int t = 0; t < 2; ++t)
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float2 pp = upk2(Pa[t][u]), qq = upk2(Qa[t][u]);
                        float sx = pp.x + qq.y, sy = pp.y - qq.x;              // conj(a) b
                        sx += __shfl_xor_sync(0xffffffffu, sx, 1); sy += __shfl_xor_sync(0xffffffffu, sy, 1);
                        sx += __shfl_xor_sync(0xffffffffu, sx, 2); sy += __shfl_xor_sync(0xffffffffu, sy, 2);
                        if (on && l4 == 0) {
                            const bool live = (c1 + t < c2 + u) && (c2 + u < nbj);
                            Sg[j * TC_NB * TC_NB + (c1 + t) * TC_NB + c2 + u] = live ? make_float2(sx, sy) : make_float2(0.f, 0.f);
                        }
                    }
            }
        }
        __syncthreads();
        TC_MARK(TCP_GRAM)
        // rows of Y = V T: one work item per (block, coordinate); the results stay in registers across a barrier
        // because block 0's rows are stored over the Gram matrices
        {
            float2 y[TC_NB];
            int keep_row = -1;                                             // this thread's row of block 0 (stored last)
            for (int it0 = 0; it0 < pl.yrows; it0 += TC_NT) {
                const int item = it0 + tid;
                if (!worker || item >= pl.yrows) continue;
                int j = 0;
                if (item < pl.yrow0[0]) {
                    j = 1;
#pragma unroll
                    for (int jj = 2; jj < 4; ++jj)
                        if (jj < pl.nblk && item >= pl.yrow0[jj]) j = jj;
                }
                const int k0 = pl.k0[j];
                const int arow = k0 + 1 + (item - pl.yrow0[j]);            // coordinate of this row of Y
                const float2* S = Sg + j * TC_NB * TC_NB;
#pragma unroll
                for (int c = 0; c < TC_NB; ++c) {
                    const int k = k0 + c;
                    float2 v = make_float2(0.f, 0.f);
                    float2 tk = make_float2(0.f, 0.f);
                    if (c < pl.nb[j]) {
                        tk = taus[k];
                        if (arow > k) v = Vs[voff(k, d) - (k + 1) + arow];
                    }
                    // v - sum_m y_m S[m][c]:  y s = y.x*(s.x,s.y) + y.y*(-s.y,s.x), kept as P += y.x*s, Q += y.y*s (two
                    // independent chains each by the parity of m), packed FFMA2
                    f32x2 P0 = pk2(0.f, 0.f), Q0 = pk2(0.f, 0.f), P1 = pk2(0.f, 0.f), Q1 = pk2(0.f, 0.f);
#pragma unroll
                    for (int m = 0; m < c; ++m) {
                        const float2 sv = S[m * TC_NB + c];
                        const f32x2 sp = pk2(sv.x, sv.y);
                        if (m & 1) { P1 = fma2(bc2(y[m].x), sp, P1); Q1 = fma2(bc2(y[m].y), sp, Q1); }
                        else       { P0 = fma2(bc2(y[m].x), sp, P0); Q0 = fma2(bc2(y[m].y), sp, Q0); }
                    }
                    const float2 p0 = upk2(P0), q0 = upk2(Q0), p1 = upk2(P1), q1 = upk2(Q1);
                    const float sx = v.x - ((p0.x + p1.x) - (q0.y + q1.y)), sy = v.y - ((p0.y + p1.y) + (q0.x + q1.x));
                    y[c] = cmul(tk, make_float2(sx, sy));
                }
                if (j == 0) {
                    keep_row = item - pl.yrow0[0];                         // always the thread's last item
                } else {
                    float4* dst = reinterpret_cast<float4*>(smem + pl.yoff[j]) + (size_t)(item - pl.yrow0[j]) * (TC_NB / 2);
#pragma unroll
                    for (int c = 0; c < TC_NB; c += 2) dst[c / 2] = make_float4(y[c].x, y[c].y, y[c + 1].x, y[c + 1].y);
                }
            }
            __syncthreads();                   // every reader of the Gram matrices is done: block 0's rows go over them
            if (keep_row >= 0) {
                float4* dst = reinterpret_cast<float4*>(smem + pl.yoff[0]) + (size_t)keep_row * (TC_NB / 2);
#pragma unroll
                for (int c = 0; c < TC_NB; c += 2) dst[c / 2] = make_float4(y[c].x, y[c].y, y[c + 1].x, y[c + 1].y);
            }
        }
        TC_MARK(TCP_YSOLVE)

        // ================= blocks, last to first =================
        // tile builders (all threads).  V tile of block j in buffer A: 96 rows x K2 window positions, K-major (LBO = 96*16):
        //   rows  0..23 [ Vr(a) | Vi(rev a)] hi   rows 24..47 [-Vi(a) | Vr(rev a)] hi   rows 48..95 the same, lo parts
        // lane -> (c & 7, coordinate & 3): the 32 stores of a warp fall into 32 different banks
        auto build_vtile = [&](int j) {
            if (!worker) return;
            const int k0 = pl.k0[j], nbj = pl.nb[j], a0 = pl.a0[j];
            const int Na = dp - a0, K2 = 2 * Na;
            for (int it = warp; it < 3 * (Na / 4); it += TC_NT / 32) {
                const int chi = it % 3, ag = it / 3;
                const int c = 8 * chi + (lane >> 2), ar = 4 * ag + (lane & 3), acol = a0 + ar, k = k0 + c;
                float2 v = make_float2(0.f, 0.f);
                if (c < nbj && acol > k && acol < d) v = Vs[voff(k, d) - (k + 1) + acol];
                const int k1 = ar, k2 = K2 - 1 - ar;
                const int o_r1 = (k1 >> 2) * 384 + c * 4 + (k1 & 3), o_r2 = (k2 >> 2) * 384 + c * 4 + (k2 & 3);
                ADMM_ASSERT(o_r1 >= 0 && o_r2 >= 0 && (o_r1 + 288) * 4 < pl.off_b - pl.off_a && (o_r2 + 288) * 4 < pl.off_b - pl.off_a);
                ADMM_ASSERT(!(c < nbj && acol > k && acol < d) || (voff(k, d) - (k + 1) + acol >= 0 && voff(k, d) - (k + 1) + acol < nv));
                const float hr = rna_tf32(v.x), hi_ = rna_tf32(v.y);
                const float lr = v.x - hr, li = v.y - hi_;
                bufA[o_r1] = hr;        bufA[o_r1 + 192] = lr;         // Pr row, real-plane position (hi | lo rows)
                bufA[o_r2] = hi_;       bufA[o_r2 + 192] = li;         // Pr row, imaginary-plane position
                bufA[o_r1 + 96] = -hi_; bufA[o_r1 + 288] = -li;        // Pi row, real-plane position
                bufA[o_r2 + 96] = hr;   bufA[o_r2 + 288] = lr;         // Pi row, imaginary-plane position
            }
        };
        // Y tiles of block j in buffer B (K = c, LBO = K2*16), order YPr_hi, YPr_lo, YPi_hi, YPi_lo:
        //   tile for A = Pr: rows [-Yr(a) | -Yi(rev a)],  tile for A = Pi: rows [+Yi(a) | -Yr(rev a)]
        // lane -> (coordinate & 7, c & 3): conflict-free stores
        auto build_ytile = [&](int j) {
            if (!worker) return;
            const int k0 = pl.k0[j], a0 = pl.a0[j];
            const int Na = dp - a0, K2 = 2 * Na;
            const float2* Yj = reinterpret_cast<const float2*>(smem + pl.yoff[j]);
            float* YPr_hi = bufB;
            float* YPr_lo = bufB + 6 * K2 * 4;
            float* YPi_hi = bufB + 2 * 6 * K2 * 4;
            float* YPi_lo = bufB + 3 * 6 * K2 * 4;
            for (int it = warp; it < 6 * (Na / 8); it += TC_NT / 32) {
                const int chi = it % 6, ag = it / 6;
                const int c = 4 * chi + (lane & 3), ar = 8 * ag + (lane >> 2), acol = a0 + ar;
                float2 yv = make_float2(0.f, 0.f);
                if (acol > k0 && acol < d) yv = Yj[(size_t)(acol - (k0 + 1)) * TC_NB + c];
                const int n1 = ar, n2 = K2 - 1 - ar;
                const int o1 = (c >> 2) * (K2 * 4) + n1 * 4 + (c & 3), o2 = (c >> 2) * (K2 * 4) + n2 * 4 + (c & 3);
                ADMM_ASSERT(o1 >= 0 && o2 >= 0 && (3 * 6 * K2 * 4 + o1) * 4 < pl.off_s - pl.off_b && (3 * 6 * K2 * 4 + o2) * 4 < pl.off_s - pl.off_b);
                const float hr = rna_tf32(yv.x), hi_ = rna_tf32(yv.y);
                const float lr = yv.x - hr, li = yv.y - hi_;
                YPr_hi[o1] = -hr;  YPr_lo[o1] = -lr;
                YPr_hi[o2] = -hi_; YPr_lo[o2] = -li;
                YPi_hi[o1] = hi_;  YPi_lo[o1] = li;
                YPi_hi[o2] = -hr;  YPi_lo[o2] = -lr;
            }
        };
        build_vtile(pl.nblk - 1);
        publish();
        TC_MARK(TCP_VTILE)
        for (int j = pl.nblk - 1; j >= 0; --j) {
            const int a0 = pl.a0[j];
            const int Na = dp - a0, K2 = 2 * Na;
            if (j == 0 && has_next) {            // buffer A holds the last V tile: the reflectors and tau are free
                prefetch_vs(sig + gridDim.x);
                prefetch_small(sig + gridDim.x, buf ^ 1);
            }
            cp_async_commit();
            // ---- GEMM 1: D[0,96) = X_hi(window) [V_hi ; V_lo]^T, then D[0,48) += X_lo(window) V_hi^T
            if (issuer) {
                const uint64_t vd = tc::smem_desc(tc::smem_u32(bufA), 1536, 128);
                const int nks = K2 / 8;
                tc::mma_chain_ts(tP, tXhi + a0, vd, 192, nks, tc::idesc_tf32(128, 96), 0u);
                tc::mma_chain_ts(tP, tXlo + a0, vd, 192, nks, tc::idesc_tf32(128, 48), 1u);
                tc::mma_commit(bar_mma);
            }
            build_ytile(j);                      // overlaps GEMM 1
            TC_MARK(TCP_PSPLIT_YTILE)
            wait_mma();
            TC_MARK(TCP_GEMM1)
            // ---- P = D[c] + D[48+c], split in place: hi -> [0,48), lo -> [48,96)
            for (int g = 3 * hh; worker && g < 3 * hh + 3; ++g) {
                uint32_t v0[8], v1[8];
                tc::tmem_ld8(tP + lane_off + 8 * g, v0);
                tc::tmem_ld8(tP + lane_off + 48 + 8 * g, v1);
                tc::tmem_wait_ld();
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const float x = __uint_as_float(v0[jj]) + __uint_as_float(v1[jj]);
                    const float h = rna_tf32(x);
                    v0[jj] = __float_as_uint(h);
                    v1[jj] = __float_as_uint(x - h);
                }
                tc::tmem_st8(tP + lane_off + 8 * g, v0);
                tc::tmem_st8(tP + lane_off + 48 + 8 * g, v1);
            }
            publish();
            // ---- GEMM 2: X_lo(window) += Pr * YPr^T + Pi * YPi^T, 3 split terms each
            if (issuer) {
                const uint32_t idesc = tc::idesc_tf32(128, K2);
                const uint32_t lbo = (uint32_t)K2 * 16, step = 2 * lbo / 16;
                const uint32_t yb = tc::smem_u32(bufB), tsz = 6u * K2 * 16;
                const uint64_t yprh = tc::smem_desc(yb, lbo, 128), yprl = tc::smem_desc(yb + tsz, lbo, 128);
                const uint64_t ypih = tc::smem_desc(yb + 2 * tsz, lbo, 128), ypil = tc::smem_desc(yb + 3 * tsz, lbo, 128);
                const uint32_t dw = tXlo + a0;
                tc::mma_chain_ts(dw, tP, yprh, step, TC_NB / 8, idesc, 1u);                 // Pr_hi * YPr_hi
                tc::mma_chain_ts(dw, tP, yprl, step, TC_NB / 8, idesc, 1u);                 // Pr_hi * YPr_lo
                tc::mma_chain_ts(dw, tP + 48, yprh, step, TC_NB / 8, idesc, 1u);            // Pr_lo * YPr_hi
                tc::mma_chain_ts(dw, tP + TC_NB, ypih, step, TC_NB / 8, idesc, 1u);         // Pi_hi * YPi_hi
                tc::mma_chain_ts(dw, tP + TC_NB, ypil, step, TC_NB / 8, idesc, 1u);         // Pi_hi * YPi_lo
                tc::mma_chain_ts(dw, tP + 48 + TC_NB, ypih, step, TC_NB / 8, idesc, 1u);    // Pi_lo * YPi_hi
                tc::mma_commit(bar_mma);
            }
            if (j > 0) build_vtile(j - 1);       // overlaps GEMM 2 (GEMM 1 of this block has completed: buffer A is free)
            TC_MARK(TCP_VTILE)
            wait_mma();
            TC_MARK(TCP_GEMM2)
            // ---- re-split the window: x = hi + lo, hi' = rna(x), lo' = x - hi'   (last block: done by the W staging)
            if (j > 0) {
                // two register sets: the loads of the next column group are in flight while this one is re-split
                const int g0 = a0 / 8 + hh, g1 = worker ? (DP2 - a0) / 8 : 0;
                uint32_t ah[8], al[8], bh[8], bl[8];
                if (g0 < g1) {
                    tc::tmem_ld8(tXhi + lane_off + 8 * g0, ah);
                    tc::tmem_ld8(tXlo + lane_off + 8 * g0, al);
                }
                for (int g = g0; g < g1; g += 4) {
                    tc::tmem_wait_ld();
                    if (g + 2 < g1) {
                        tc::tmem_ld8(tXhi + lane_off + 8 * (g + 2), bh);
                        tc::tmem_ld8(tXlo + lane_off + 8 * (g + 2), bl);
                    }
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const float x = __uint_as_float(ah[jj]) + __uint_as_float(al[jj]);
                        const float h = rna_tf32(x);
                        ah[jj] = __float_as_uint(h);
                        al[jj] = __float_as_uint(x - h);
                    }
                    tc::tmem_st8(tXhi + lane_off + 8 * g, ah);
                    tc::tmem_st8(tXlo + lane_off + 8 * g, al);
                    if (g + 2 < g1) {
                        tc::tmem_wait_ld();
                        if (g + 4 < g1) {
                            tc::tmem_ld8(tXhi + lane_off + 8 * (g + 4), ah);
                            tc::tmem_ld8(tXlo + lane_off + 8 * (g + 4), al);
                        }
#pragma unroll
                        for (int jj = 0; jj < 8; ++jj) {
                            const float x = __uint_as_float(bh[jj]) + __uint_as_float(bl[jj]);
                            const float h = rna_tf32(x);
                            bh[jj] = __float_as_uint(h);
                            bl[jj] = __float_as_uint(x - h);
                        }
                        tc::tmem_st8(tXhi + lane_off + 8 * (g + 2), bh);
                        tc::tmem_st8(tXlo + lane_off + 8 * (g + 2), bl);
                    }
                }
                publish();                       // re-split X and the next V tile -> visible to GEMM 1 of block j-1
            }
            TC_MARK(TCP_RESPLIT)
        }

        // ================= rebuild: W tiles, G = W W^H on the tensor cores =================
        {
            const int lbw = pl.lbo_w / 4;                                      // floats per K chunk
            const int tsz = (dp / 4) * lbw;                                    // floats per W tile
            float* Wt = bufA;                                                  // Wr_hi, Wr_lo, Wi_hi, Wi_lo
            const int ngr = dp / 8;
            const int kbase = (row >> 2) * lbw + (row & 3);
            // (tcgen05.ld is warp-collective: every lane takes part, only the shared-memory stores are predicated;
            //  two register sets keep the next group's loads in flight while this one is scaled, split and stored)
            auto stage_group = [&](int g, const uint32_t (&vh)[8], const uint32_t (&vl)[8]) {
                const bool im = g >= ngr;
                float* Th = Wt + (im ? 2 * tsz : 0);
                float* Tl = Th + tsz;
                if (row < dp) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int acol = im ? (DP2 - 1 - (8 * g + jj)) : (8 * g + jj);
                        const float w = sl * (__uint_as_float(vh[jj]) + __uint_as_float(vl[jj]));
                        const float h = rna_tf32(w);
                        ADMM_ASSERT(acol >= 0 && acol < dp && (kbase + acol * 4) < tsz);
                        Th[kbase + acol * 4] = h;
                        Tl[kbase + acol * 4] = w - h;
                    }
                }
            };
            if (worker) {
                uint32_t ah[8], al[8], bh[8], bl[8];
                const int g1 = 2 * ngr;
                tc::tmem_ld8(tXhi + lane_off + 8 * hh, ah);
                tc::tmem_ld8(tXlo + lane_off + 8 * hh, al);
                for (int g = hh; g < g1; g += 4) {
                    tc::tmem_wait_ld();
                    if (g + 2 < g1) {
                        tc::tmem_ld8(tXhi + lane_off + 8 * (g + 2), bh);
                        tc::tmem_ld8(tXlo + lane_off + 8 * (g + 2), bl);
                    }
                    stage_group(g, ah, al);
                    if (g + 2 < g1) {
                        tc::tmem_wait_ld();
                        if (g + 4 < g1) {
                            tc::tmem_ld8(tXhi + lane_off + 8 * (g + 4), ah);
                            tc::tmem_ld8(tXlo + lane_off + 8 * (g + 4), al);
                        }
                        stage_group(g + 2, bh, bl);
                    }
                }
            }
            publish();
            TC_MARK(TCP_WSTAGE)
            if (issuer) {
                const uint32_t w0 = tc::smem_u32(Wt), lbo = (uint32_t)pl.lbo_w, step = 2 * lbo / 16;
                uint64_t tb[4];                                                 // Wr_hi, Wr_lo, Wi_hi, Wi_lo
#pragma unroll
                for (int t = 0; t < 4; ++t) tb[t] = tc::smem_desc(w0 + (uint32_t)t * tsz * 4, lbo, 128);
                const uint32_t id = tc::idesc_tf32(128, dp), idn = tc::idesc_tf32(128, dp, 1);
                const int nks = dp / 8;
                const uint32_t gr = tbase, gi = tbase + dp;
                // G_r (columns [0,dp)) = Wr Wr^T + Wi Wi^T ; G_i (columns [dp,2dp)) = Wi Wr^T - Wr Wi^T ; 3 split terms each
                tc::mma_chain_ss(gr, tb[0], step, tb[0], step, nks, id, 0u);
                tc::mma_chain_ss(gr, tb[0], step, tb[1], step, nks, id, 1u);
                tc::mma_chain_ss(gr, tb[1], step, tb[0], step, nks, id, 1u);
                tc::mma_chain_ss(gr, tb[2], step, tb[2], step, nks, id, 1u);
                tc::mma_chain_ss(gr, tb[2], step, tb[3], step, nks, id, 1u);
                tc::mma_chain_ss(gr, tb[3], step, tb[2], step, nks, id, 1u);
                tc::mma_chain_ss(gi, tb[2], step, tb[0], step, nks, id, 0u);
                tc::mma_chain_ss(gi, tb[2], step, tb[1], step, nks, id, 1u);
                tc::mma_chain_ss(gi, tb[3], step, tb[0], step, nks, id, 1u);
                tc::mma_chain_ss(gi, tb[0], step, tb[2], step, nks, idn, 1u);
                tc::mma_chain_ss(gi, tb[0], step, tb[3], step, nks, idn, 1u);
                tc::mma_chain_ss(gi, tb[1], step, tb[2], step, nks, idn, 1u);
                tc::mma_commit(bar_mma);
            }
            wait_mma();
        }
        __syncthreads();                 // every thread is past the MMA wait: tile area and Z box are free
        TC_MARK(TCP_REBUILD)
        if (has_next && tid == 0) load_z(sig + gridDim.x);

        // ================= epilogue: lower triangle of G -> residual, packed store =================
        float rsq = 0.f;
        {
            const int ngr = dp / 8;
            // warp-uniform trip count (the warp's last row is 32q+31); per-lane predicates only guard the stores
            for (int g = hh; worker && g < ngr && 8 * g <= 32 * q + 31; g += 2) {
                uint32_t gr[8], gi[8];
                tc::tmem_ld8(tbase + lane_off + 8 * g, gr);
                tc::tmem_ld8(tbase + lane_off + dp + 8 * g, gi);
                tc::tmem_wait_ld();
                if (row < d) {
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int b = 8 * g + jj;
                        if (b > row) continue;
                        float2 gv = make_float2(__uint_as_float(gr[jj]), __uint_as_float(gi[jj]));
                        if (b == row) gv.y = 0.f;
                        ADMM_ASSERT(pk(row, b) < npk);
                        Gs[pk(row, b)] = gv;
                        if (a.with_c > 0) {
                            float2 c = make_float2(0.f, 0.f);
                            if (b == row) c.x = (row < n) ? hs[row] : c1z;
                            else if (row == n) c = cconj(phis[b]);
                            const float rx = gv.x - c.x, ry = gv.y - c.y;
                            rsq += (b == row ? 1.f : 2.f) * (rx * rx + ry * ry);
                        }
                    }
                }
            }
        }
        tc::tmem_wait_ld();
        tc::tc_fence_before_sync();
        {
            float v[1] = {rsq};
            block_sum<1>(v, red);                        // two __syncthreads inside: Gs complete afterwards
            if (tid == 0 && a.with_c > 0) a.r_out[sig] = sqrtf(v[0]);
        }
        tc::tc_fence_after_sync();
        {
            float2* GV = a.GV + (size_t)sig * npk;
            for (int idx = tid; worker && idx < npk; idx += TC_NT) GV[idx] = Gs[idx];
        }
        __syncthreads();                 // Gs (tile area) is reused by the next signal's Gram matrices
        TC_MARK(TCP_EPILOGUE)
        if (profiling) pacc[TCP_SIGNALS] += 1;
    }
    if (profiling) {
#pragma unroll
        for (int i = 0; i < TCP_NPHASE; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.prof) + i, (unsigned long long)pacc[i]);
    }
#undef TC_MARK
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

}  // namespace admmnet
