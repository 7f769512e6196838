// Unit probe of the tcgen05 / TMEM / TMA layer (tc.cuh): one 128 x N x K tf32 GEMM tile, D = D0 + (+-A) * B^T, with the
// operands exactly as given (the caller passes tf32-representable values, so the result is exact up to fp32
// accumulation).  Every staging variant the tail kernel relies on has a flag; tests/test_gpu_parity.py runs them all
// against numpy.  This is a debug tap of the C ABI (admmnet_tc_gemm_probe), not a product entry point.
#include "tc.cuh"

namespace admmnet {

enum TcProbeFlags : int {
    TCP_A_TMEM = 1,      // A operand from tensor memory (written with tcgen05.st) instead of shared memory
    TCP_A_NEG = 2,       // negate A through the instruction descriptor
    TCP_A_MN = 4,        // A staged MN-major (shared memory only)
    TCP_B_MN = 8,        // B staged MN-major
    TCP_TMA = 16,        // A and B brought in by TMA (cp.async.bulk.tensor.2d) into a row-major staging area first
    TCP_SPLIT3 = 32,     // 3xTF32: A, B are arbitrary fp32; hi/lo split on the fly, three MMAs per K step
    TCP_DCOL8 = 64,      // accumulator at TMEM column 8 instead of 0 (windows of the tail kernel start at multiples of 8)
};

struct TcProbeArgs {
    const float* A;      // [128][K]
    const float* B;      // [N][K]
    const float* D0;     // [128][N] or null
    float* out;          // [128][N]
    int N, K, flags;
};

// shared memory: [0] mbarriers + tmem slot | canonical A hi, A lo | canonical B hi, B lo | row-major staging (TMA)
__global__ void __launch_bounds__(128, 1)
k_tc_probe(TcProbeArgs a, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem);
    uint64_t* bar_tma = bar_mma + 1;
    uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 16);
    const int N = a.N, K = a.K;
    float* Ahi = reinterpret_cast<float*>(smem + 1024);
    float* Alo = Ahi + 128 * K;
    float* Bhi = Alo + 128 * K;
    float* Blo = Bhi + N * K;
    float* stA = Blo + N * K;          // [128][K] row-major (TMA destination)
    float* stB = stA + 128 * K;        // [N][K]
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool a_tmem = a.flags & TCP_A_TMEM, a_mn = a.flags & TCP_A_MN, b_mn = a.flags & TCP_B_MN;
    const bool use_tma = a.flags & TCP_TMA, split3 = a.flags & TCP_SPLIT3;

    if (tid == 0) {
        tc::mbar_init(bar_mma, 1);
        tc::mbar_init(bar_tma, 1);
        tc::mbar_fence_init();
    }
    if (warp == 0) tc::tmem_alloc(tslot, 512);
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();
    const uint32_t tbase = *tslot;
    const uint32_t tD = tbase + ((a.flags & TCP_DCOL8) ? 8 : 0);   // columns [0, N) or [8, N+8): accumulator
    const uint32_t tAhi = tbase + 256;         // columns [256, 256+K): A hi in TMEM
    const uint32_t tAlo = tbase + 256 + 64;    // A lo
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

    // ---- bring A and B in
    const float* Asrc = a.A;
    const float* Bsrc = a.B;
    if (use_tma) {
        if (tid == 0) {
            tc::mbar_expect_tx(bar_tma, (uint32_t)((128 + N) * K * sizeof(float)));
            tc::tma_load_2d(stA, &tmA, 0, 0, bar_tma);
            tc::tma_load_2d(stB, &tmB, 0, 0, bar_tma);
        }
        tc::mbar_wait(bar_tma, 0);
        Asrc = stA;
        Bsrc = stB;
    }
    // canonical no-swizzle layouts (tc.cuh): K-major  byte(r,k) = (k/4)*LBO + r*16 + (k%4)*4, LBO = rows*16
    //                                        MN-major byte(r,k) = (k/8)*LBO + (r/4)*128 + (k%8)*16 + (r%4)*4, LBO = rows*32
    auto put = [&](float* hi, float* lo, int rows, bool mn, int r, int k, float v) {
        const int off = mn ? (k / 8) * rows * 8 + (r / 4) * 32 + (k % 8) * 4 + (r % 4)
                           : (k / 4) * rows * 4 + r * 4 + (k % 4);
        float h = v, l = 0.f;
        if (split3) tc::split_tf32(v, h, l);
        hi[off] = h;
        lo[off] = l;
    };
    for (int idx = tid; idx < 128 * K; idx += 128) put(Ahi, Alo, 128, a_mn, idx / K, idx % K, Asrc[idx]);
    for (int idx = tid; idx < N * K; idx += 128) put(Bhi, Blo, N, b_mn, idx / K, idx % K, Bsrc[idx]);
    if (a_tmem) {
        // row m = tid of A into lane m, columns [0,K) of the two TMEM A regions
        for (int k0 = 0; k0 < K; k0 += 8) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float hv = Asrc[tid * K + k0 + j], lv = 0.f;
                if (split3) tc::split_tf32(Asrc[tid * K + k0 + j], hv, lv);
                h[j] = __float_as_uint(hv);
                l[j] = __float_as_uint(lv);
            }
            tc::tmem_st8(tAhi + lane_off + k0, h);
            tc::tmem_st8(tAlo + lane_off + k0, l);
        }
    }
    // optional initial accumulator
    if (a.D0) {
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(a.D0[tid * N + c0 + j]);
            tc::tmem_st8(tD + lane_off + c0, v);
        }
    }
    tc::tmem_wait_st();
    tc::fence_async_smem();
    tc::tc_fence_before_sync();
    __syncthreads();
    tc::tc_fence_after_sync();

    // ---- one thread issues the MMAs
    if (tid == 0) {
        const uint32_t idesc = tc::idesc_tf32(128, N, (a.flags & TCP_A_NEG) ? 1 : 0, 0, a_mn ? 1 : 0, b_mn ? 1 : 0);
        const uint32_t a_lbo = a_mn ? 128 * 32 : 128 * 16, b_lbo = b_mn ? N * 32 : N * 16;
        const uint32_t a_step = a_mn ? a_lbo : 2 * a_lbo, b_step = b_mn ? b_lbo : 2 * b_lbo;   // bytes per K = 8
        uint32_t acc = a.D0 ? 1u : 0u;
        const int nterm = split3 ? 3 : 1;
        for (int term = 0; term < nterm; ++term) {
            // term 0: hi*hi, 1: hi*lo, 2: lo*hi
            const float* As = term == 2 ? Alo : Ahi;
            const float* Bs = term == 1 ? Blo : Bhi;
            const uint32_t At = term == 2 ? tAlo : tAhi;
            for (int ks = 0; ks < K / 8; ++ks) {
                const uint64_t bd = tc::smem_desc(tc::smem_u32(Bs) + ks * b_step, b_lbo, 128);
                if (a_tmem) {
                    tc::mma_tf32_ts(tD, At + ks * 8, bd, idesc, acc);
                } else {
                    const uint64_t ad = tc::smem_desc(tc::smem_u32(As) + ks * a_step, a_lbo, 128);
                    tc::mma_tf32_ss(tD, ad, bd, idesc, acc);
                }
                acc = 1u;
            }
        }
        tc::mma_commit(bar_mma);
    }
    tc::mbar_wait(bar_mma, 0);
    tc::tc_fence_after_sync();

    // ---- accumulator -> global
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        tc::tmem_ld8(tD + lane_off + c0, v);
        tc::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 8; ++j) a.out[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    tc::tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// 2D fp32 tensor [rows][cols] with row pitch `pitch_bytes` (multiple of 16), box = [box_rows][box_cols], no swizzle
inline bool make_tmap_2d_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_bytes,
                             uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace admmnet
