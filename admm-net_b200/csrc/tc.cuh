// Blackwell (sm_100a) tensor-core / tensor-memory / TMA primitives used by the tail kernel: thin inline-PTX wrappers
// around tcgen05.{alloc,dealloc,mma,commit,ld,st,fence,wait}, mbarrier and cp.async.bulk.tensor, plus the shared-memory
// matrix descriptor and instruction descriptor of kind::tf32 UMMA.
//
// Layout facts used throughout (no swizzle, "interleave" canonical layouts; 16-byte units):
//   K-major operand (rows = M or N index, K contiguous in 16-byte chunks of 4 tf32):
//       byte(r, k) = (r % 8) * 16 + (r / 8) * SBO + (k / 4) * LBO + (k % 4) * 4
//     a core matrix is 8 rows x 16 bytes = 128 contiguous bytes; one tf32 MMA consumes K = 8 = two chunks (LBO apart).
//   TMEM: 128 lanes x 512 columns of 32 bit; address = (lane << 16) | column.  An accumulator (or a TMEM A operand)
//     of an M = 128 MMA has row m in lane m.  A warp can only touch lanes 32*(warp_id % 4) .. +31.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace admmnet {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded spin: a wrong descriptor or a lost arrival must not hang the GPU — after ~2^28 polls (seconds) the
// kernel traps and the launch reports an error instead.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
        if (spins > (1u << 28)) __trap();
    }
}

// ------------------------------------------------------------------------------------------ proxies / fences
// generic-proxy writes to shared memory -> visible to the async proxy (UMMA operand reads, TMA stores)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ TMEM allocation
// One full warp allocates `cols` (power of two >= 32) columns and writes the base address to *slot (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// ------------------------------------------------------------------------------------------ TMEM <-> registers
// 32x32b shape: thread t of the warp owns lane (lane_base + t); .xN moves N consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ descriptors
// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor bit layout): start address, leading and
// stride byte offsets in 16-byte units, descriptor version 1 (Blackwell) in bits [46,48), layout type 0 in [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor of kind::tf32 (cute::UMMA::InstrDescriptor): D = fp32, A = B = tf32, M x N tile,
// optional negation of A / B, operand major-ness (0 = K-major, 1 = MN-major).
__host__ __device__ __forceinline__ uint32_t idesc_tf32(int M, int N, int a_neg = 0, int b_neg = 0, int a_mn = 0,
                                                        int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_neg << 13) | ((uint32_t)b_neg << 14) |
           ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ MMA issue (ONE thread)
// D[tmem] (+)= A[smem] * B[smem]^T ; accumulate = 0 overwrites D.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Chains of `n` MMAs over consecutive K = 8 steps: the A operand advances by 8 TMEM columns (or by a_step16 16-byte
// units in shared memory), B by b_step16.  Descriptors are advanced with one 64-bit add (only the 14-bit address
// field changes: shared memory is < 256 KB), so the single issuing thread spends ~6 instructions per MMA.
__device__ __forceinline__ void mma_chain_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t b_step16,
                                             int n, uint32_t idesc, uint32_t first_accumulate) {
    mma_tf32_ts(d_tmem, a_tmem, b_desc, idesc, first_accumulate);
#pragma unroll 4
    for (int ks = 1; ks < n; ++ks) {
        a_tmem += 8;
        b_desc += b_step16;
        mma_tf32_ts(d_tmem, a_tmem, b_desc, idesc, 1u);
    }
}
__device__ __forceinline__ void mma_chain_ss(uint32_t d_tmem, uint64_t a_desc, uint32_t a_step16, uint64_t b_desc,
                                             uint32_t b_step16, int n, uint32_t idesc, uint32_t first_accumulate) {
    mma_tf32_ss(d_tmem, a_desc, b_desc, idesc, first_accumulate);
#pragma unroll 4
    for (int ks = 1; ks < n; ++ks) {
        a_desc += a_step16;
        b_desc += b_step16;
        mma_tf32_ss(d_tmem, a_desc, b_desc, idesc, 1u);
    }
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------ TMA (tiled, 2D)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}

// ------------------------------------------------------------------------------------------ 3xTF32 split
// hi = x rounded to nearest tf32 (cvt.rna: 13 low mantissa bits zero, so exact whatever the tensor core does with its
// input bits), lo = x - hi (exact in fp32, |lo| <= 2^-11 |x|; the tensor core uses its leading 10 bits).
// a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b : relative error ~2^-21 per product with fp32 accumulation — measured
// against fp64 on the GEMM probe and through the whole 10-layer forward.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    hi = __uint_as_float(r);
    lo = x - hi;
}

}  // namespace tc
}  // namespace admmnet
