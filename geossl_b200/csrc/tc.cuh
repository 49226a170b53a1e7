// sm_100a tensor-core plumbing written directly in PTX: tcgen05.mma (kind::f16) with shared-memory
// operand descriptors, TMEM allocation / loads, mbarriers, and the 128-byte swizzled K-major operand
// layout that st.shared producers fill by hand (operands here are computed on chip, so there is no TMA
// load in front of the MMA -- the generic->async proxy fence below is what makes the stores visible).
//
// fp32-grade accuracy on 16-bit tensor cores: every fp32 operand x is split as x = hi + lo with
// hi = rn16(x), lo = rn16(x - hi) and a product A.B is issued as three MMAs
//     A_hi.B_hi + A_lo.B_hi + A_hi.B_lo            (the lo.lo term is below the kept precision)
// accumulating in fp32 in TMEM.  With fp16 parts (11+11 significant bits) the result is fp32-grade
// (~2^-22 per product) for operands of bounded range (rbf values, activations, weights); with bf16 parts
// (8+8 bits, ~2^-17) the dynamic range is fp32's, which is what gradient operands need.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace geossl {
namespace tc {

// ------------------------------------------------------------------------------------------ addresses
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
// One arrive per warp: every lane's prior writes / fences are ordered before it by the __syncwarp.
__device__ __forceinline__ void warp_arrive(uint32_t bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
constexpr uint32_t kWaitHintNs = 20000;
// Spin on the phase parity.  Bounded: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"      // suspend-time hint: sleep in hardware,
            "selp.u32 %0, 1, 0, p;\n\t}"                                        // not in an issue-slot-eating spin loop
            : "=r"(done) : "r"(bar), "r"(parity), "r"(kWaitHintNs) : "memory");
        if (spin > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------ bulk async copy (TMA, 1-D)
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (bytes % 16 == 0, 16-byte aligned), completion counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_dst), "l"(gmem_src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------ tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {   // whole warp, cols = pow2 >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// Warp-uniform leader election.  The MMA warp runs its loop with all 32 lanes (so every descriptor / address is
// provably warp-uniform and lands in uniform registers) and only the elected lane issues tcgen05.mma / commit.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// all previously issued MMAs of this thread complete -> one arrive on the mbarrier
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// D[tmem] (+)= A[smem desc] . B[smem desc]^T   (M x N x 16, 16-bit operands, fp32 accumulate)
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// D[tmem] (+)= A[tmem] . B[smem desc]^T : A (M = 128 lanes) lives in tensor memory, K-major, two 16-bit values per
// 32-bit column (column c of the operand holds k = 2c in its low half and k = 2c+1 in its high half).
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 8 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid) -> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ------------------------------------------------------------------------------------------ descriptors
// Shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 bytes (64 16-bit values of
// K), 8-row groups 1024 bytes apart (SBO), tile base 1024-byte aligned.  Advancing K by 16 elements is
// +32 bytes on the start address (the hardware applies the XOR swizzle to the final address bits).
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)      // start address  [0,14)
           | ((uint64_t)1 << 16)                        // leading byte offset (unused for swizzled K-major)
           | ((uint64_t)(1024 >> 4) << 32)              // stride byte offset   [32,46)
           | ((uint64_t)1 << 46)                        // descriptor version (Blackwell)
           | ((uint64_t)2 << 61);                       // SWIZZLE_128B
}
// MN-major operand over the SAME memory image ([k-row][64 MN values = 128 bytes], 8-row groups at 1024 B):
// LBO = byte distance between 64-wide MN blocks, SBO = 1024.  Advancing K by 16 rows is +2048 bytes.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t mn_block_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4)
           | ((uint64_t)((mn_block_bytes >> 4) & 0x3FFFu) << 16)
           | ((uint64_t)(1024 >> 4) << 32)
           | ((uint64_t)1 << 46)
           | ((uint64_t)2 << 61);
}
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;
// Instruction descriptor for kind::f16: fp32 accumulate, both operands `fmt`, majors (0 = K, 1 = MN).
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t fmt, int M, int N, uint32_t a_mn_major = 0, uint32_t b_mn_major = 0) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------------------------------ operand tiles
// Byte offset of element (row, k) inside one [rows x 64] 16-bit K-major SW128 block (rows*128 bytes).
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t k) {
    return row * 128u + ((((k >> 3) ^ (row & 7u)) << 4) | ((k & 7u) << 1));
}

template <bool FP16>
struct Split;   // x = hi + lo in 16-bit parts; packs two values per 32-bit word (low half = first value)
template <>
struct Split<true> {
    static constexpr uint32_t kFmt = kFmtF16;
    __device__ static __forceinline__ void pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
        const __half2 h = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    }
};
template <>
struct Split<false> {
    static constexpr uint32_t kFmt = kFmtBF16;
    __device__ static __forceinline__ void pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
        const float2 hf = __bfloat1622float2(h);
        const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        lo = *reinterpret_cast<const uint32_t*>(&l);
    }
};

// One value -> the hi and lo images (2-byte scattered stores; used by the transposed epilogues).
template <bool FP16>
__device__ __forceinline__ void store_split1(uint8_t* hi_block, uint8_t* lo_block, uint32_t row, uint32_t k, float x) {
    const uint32_t off = sw128_offset(row, k);
    if constexpr (FP16) {
        const __half h = __float2half_rn(x);
        *reinterpret_cast<__half*>(hi_block + off) = h;
        *reinterpret_cast<__half*>(lo_block + off) = __float2half_rn(x - __half2float(h));
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(x);
        *reinterpret_cast<__nv_bfloat16*>(hi_block + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(lo_block + off) = __float2bfloat16_rn(x - __bfloat162float(h));
    }
}

// Transposed epilogues (TMEM lane = K index of the operand tile being produced, registers = 16 consecutive rows): each lane
// holds v[j] = element (row0 + j, k).  Writing them one by one is a 2-byte store per element with two lanes landing in
// every 32-bit shared-memory word -- ncu counts one bank conflict per store instruction (2.7 M per launch of the filter
// backward kernel, L1TEX the busiest unit).  Here lanes 2m / 2m+1 (k even / odd) swap half of their rows first, so that
// every lane owns BOTH k values of 8 rows and stores packed 32-bit words: half the store instructions, no conflicts.
// The even lane keeps rows {0-3, 8-11}, the odd lane rows {4-7, 12-15}: rows written by one instruction differ by 4,
// which the 128-byte swizzle maps to disjoint bank groups.  `k` must be even on even lanes and k+1 on their partners.
template <bool FP16>
__device__ __forceinline__ void store_split16_paired(uint8_t* hi_block, uint8_t* lo_block, uint32_t row0, uint32_t k,
                                                     const float (&v)[16], int lane) {
    const bool odd = lane & 1;
    const uint32_t kpair = k & ~1u;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int r = (t & 3) + ((t >> 2) << 3);                    // 0-3, 8-11
        const float keep = odd ? v[r + 4] : v[r], send = odd ? v[r] : v[r + 4];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        uint32_t h, l;
        Split<FP16>::pair(odd ? recv : keep, odd ? keep : recv, h, l);
        const uint32_t off = sw128_offset(row0 + r + (odd ? 4 : 0), kpair);
        *reinterpret_cast<uint32_t*>(hi_block + off) = h;
        *reinterpret_cast<uint32_t*>(lo_block + off) = l;
    }
}

// Store 8 consecutive K values (k0 % 8 == 0) of `row` into the hi and lo images of a SW128 block.
template <bool FP16>
__device__ __forceinline__ void store_chunk8(uint8_t* hi_block, uint8_t* lo_block, uint32_t row, uint32_t k0, const float* x) {
    uint4 h, l;
    Split<FP16>::pair(x[0], x[1], h.x, l.x);
    Split<FP16>::pair(x[2], x[3], h.y, l.y);
    Split<FP16>::pair(x[4], x[5], h.z, l.z);
    Split<FP16>::pair(x[6], x[7], h.w, l.w);
    const uint32_t off = sw128_offset(row, k0);
    *reinterpret_cast<uint4*>(hi_block + off) = h;
    *reinterpret_cast<uint4*>(lo_block + off) = l;
}

// 16 consecutive K values of this thread's row -> the hi and lo images of a TMEM-resident A operand (8 columns each).
template <bool FP16>
__device__ __forceinline__ void tmem_store_split16(uint32_t hi_taddr, uint32_t lo_taddr, const float* x) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) Split<FP16>::pair(x[2 * j], x[2 * j + 1], h[j], l[j]);
    tmem_st8(hi_taddr, h);
    tmem_st8(lo_taddr, l);
}
// Three-product group with the A operand in tensor memory.
__device__ __forceinline__ void mma3_ts(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                        uint32_t idesc, uint32_t accumulate) {
    mma_f16_ts(d_tmem, a_hi, b_hi, idesc, accumulate);
    mma_f16_ts(d_tmem, a_lo, b_hi, idesc, 1u);
    mma_f16_ts(d_tmem, a_hi, b_lo, idesc, 1u);
}

// The three-product MMA group for one 16-wide K step.
__device__ __forceinline__ void mma3(uint32_t d_tmem, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                     uint32_t idesc, uint32_t accumulate) {
    mma_f16_ss(d_tmem, a_hi, b_hi, idesc, accumulate);
    mma_f16_ss(d_tmem, a_lo, b_hi, idesc, 1u);
    mma_f16_ss(d_tmem, a_hi, b_lo, idesc, 1u);
}

// Fast shifted softplus: max(x,0) + log1p(exp(-|x|)) - log 2 on the MUFU ex2/lg2 units.
// Absolute error <= ~4e-7 (lg2.approx on [1,2]); torch's threshold-20 branch is reproduced.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (above torch's threshold of 20 the formula already returns x exactly in fp32: log1p(exp(-20)) < ulp(20)/2)
__device__ __forceinline__ float ssp_fast(float x) {
    const float t = ex2_approx(-fabsf(x) * 1.4426950408889634f);
    return fmaf(lg2_approx(1.0f + t), 0.6931471805599453f, fmaxf(x, 0.f)) - kLog2;
}
// 1 / x on the MUFU unit (1 ulp); __frcp_rn compiles to a range check + slow-path call around the same instruction
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// sigmoid(x) = d/dx softplus(x)
__device__ __forceinline__ float sigmoid_fast(float x) {
    return rcp_approx(1.0f + ex2_approx(-x * 1.4426950408889634f));
}
// 1024-byte aligned view of the dynamic shared memory that keeps the shared address space visible to the compiler
__device__ __forceinline__ uint8_t* align1024(uint8_t* raw) { return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u); }

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc
}  // namespace geossl
