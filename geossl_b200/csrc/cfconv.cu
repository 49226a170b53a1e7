// Continuous-filter convolution aggregate and its two adjoints over the destination-sorted CSR.
//
// Replaces PyG MessagePassing.propagate + CFConv.message (Geom3D/models/schnet.py:190,194-195):
//   x_j = x[edge_index[0]] (materialised (E,F)), msg = x_j * W, torch_scatter atomicAdd by edge_index[1].
// Here: one warp per destination row, filter rows streamed with 128-bit loads that bypass L1 (through the
// `filt_row` map when the two directions of an atom pair share one row, geossl_pair_index), x rows gathered
// through L1/L2 (x is (N,F) = a few MB, L2 resident), fp32 accumulation in registers in edge order, one
// coalesced store per row: atomic free and run-to-run deterministic.
//
// HBM roofline: bytes = 4F*U (every filter row once) + 2*4F*N (x, out) + 8E (src, filt_row) + 4(N+1) (rowptr);
// U = E without pair sharing (551 B/edge), U = E/2 for untruncated graphs (300 B/edge).
#include "common.cuh"

namespace geossl {

__device__ __forceinline__ void fma4(float4& a, const float4& x, const float4& w) {
    a.x = fmaf(x.x, w.x, a.x); a.y = fmaf(x.y, w.y, a.y); a.z = fmaf(x.z, w.z, a.z); a.w = fmaf(x.w, w.w, a.w);
}

// LPR lanes cover one F-wide row with float4 each; a warp walks EPW = 32/LPR edges per step.
// One kernel serves both adjoint directions (row = target over the CSR, or row = source over the transposed view):
//   out[row] = sum_{k in [ptr[row], ptr[row+1])} filt[frow(k)] * v[idx_b[k]],   frow(k) = filt_row[e(k)] or e(k),
//   e(k) = k (CSR) or idx_a[k] (transposed view: t_eid).
// The edge ids / gather indices of up to 32 edges are fetched with ONE coalesced load per lane and broadcast by shuffle,
// so the 128-bit row loads of an unrolled step do not wait on per-edge index loads (the kernel is latency bound).
//
// Deep edition: the kernel is latency bound (DRAM 33 %, L2 25 % busy; 72 % of issue slots have no eligible warp), so what
// matters is how many filter-row loads are in flight per SM.  Eight filter rows are requested per step, but the gather
// operand (an L1/L2 hit) is fetched in two halves so that the register budget still allows four CTAs per SM.
template <int F, bool TRANSPOSED>
__global__ void __launch_bounds__(256, 4)
cfconv_gather_deep_kernel(const float* __restrict__ filt, const int32_t* __restrict__ filt_row, const float* __restrict__ v,
                          const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx_a, const int32_t* __restrict__ idx_b,
                          int n_atoms, float* __restrict__ out) {
    constexpr int LPR = F / 4, EPW = 32 / LPR, U = 8;
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wid >= n_atoms) return;
    // Rows are walked LAST to first: a row's own (contiguous) block of filter rows is touched first and the rows it shares
    // were touched just before; in the forward the tail of the freshly written filter tensor is also what L2 still holds.
    const int row = n_atoms - 1 - wid;
    const int f = (lane % LPR) * 4, sub = lane / LPR;
    const int b = __ldg(ptr + row), end = __ldg(ptr + row + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = b; base < end; base += 32) {
        const int mine = base + lane;
        int my_a = 0, my_b = 0;
        if (mine < end) {
            const int e = TRANSPOSED ? __ldg(idx_a + mine) : mine;
            my_a = filt_row ? __ldg(filt_row + e) : e;
            my_b = __ldg(idx_b + mine);
        }
        const int cnt = min(32, end - base);
        const int rb_any = __shfl_sync(0xffffffffu, my_b, 0);         // a valid row for the padded slots (multiplied by zero)
        for (int k0 = 0; k0 < cnt; k0 += EPW * U) {
            float4 w[U];
            int rb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = k0 + sub + u * EPW;                     // lanes of different sub-groups diverge on k < cnt:
                const int ra = __shfl_sync(0xffffffffu, my_a, k & 31);   // every shuffle stays outside the branch
                const int rbk = __shfl_sync(0xffffffffu, my_b, k & 31);
                if (k < cnt) { w[u] = ld_stream4(filt + (int64_t)ra * F + f); rb[u] = rbk; }
                else { w[u] = make_float4(0.f, 0.f, 0.f, 0.f); rb[u] = rb_any; }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 xv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) xv[u] = ldg4(v + (int64_t)rb[h * 4 + u] * F + f);
#pragma unroll
                for (int u = 0; u < 4; ++u) fma4(acc, xv[u], w[h * 4 + u]);
            }
        }
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    if (sub == 0) *reinterpret_cast<float4*>(out + (int64_t)row * F + f) = acc;
}

// Async edition (F = 128): the filter rows of a step go global -> shared with cp.async (no registers held while they are
// in flight), double buffered, so the NEXT step's eight rows are already requested while the current step is consumed and
// filter-row loads stay in flight continuously; each lane reads back exactly the 16 bytes it copied (no cross-lane sync).
__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool TRANSPOSED>
__global__ void __launch_bounds__(256, 3)
cfconv_gather_async_kernel(const float* __restrict__ filt, const int32_t* __restrict__ filt_row, const float* __restrict__ v,
                           const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx_a, const int32_t* __restrict__ idx_b,
                           int n_atoms, float* __restrict__ out) {
    constexpr int F = 128, U = 8;
    extern __shared__ float4 sW_raw[];                                // 64 KB: [warp][stage][row of the step][lane]
    float4 (*sW)[2][U][32] = reinterpret_cast<float4 (*)[2][U][32]>(sW_raw);
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int wid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (wid >= n_atoms) return;
    const int row = n_atoms - 1 - wid;
    const int f = lane * 4;
    const int b = __ldg(ptr + row), end = __ldg(ptr + row + 1);
    const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(&sW[w][0][0][lane]);
    constexpr uint32_t kRowB = 32 * 16, kStageB = U * kRowB;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = b; base < end; base += 32) {
        const int mine = base + lane;
        int my_a = 0, my_b = 0;
        if (mine < end) {
            const int e = TRANSPOSED ? __ldg(idx_a + mine) : mine;
            my_a = filt_row ? __ldg(filt_row + e) : e;
            my_b = __ldg(idx_b + mine);
        }
        const int cnt = min(32, end - base);
        const int n_steps = (cnt + U - 1) / U;
        auto issue = [&](int step) {                                  // rows step*U .. +U-1 of this chunk -> stage step & 1
            const uint32_t dst = s0 + (step & 1) * kStageB;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = step * U + u;
                const int ra = __shfl_sync(0xffffffffu, my_a, k & 31);
                if (k < cnt) cp_async16(dst + u * kRowB, filt + (int64_t)ra * F + f);
            }
            cp_async_commit();
        };
        issue(0);
        for (int step = 0; step < n_steps; ++step) {
            if (step + 1 < n_steps) issue(step + 1);
            else cp_async_commit();                                   // empty group keeps the wait count uniform
            int rb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int k = step * U + u;
                rb[u] = __shfl_sync(0xffffffffu, my_b, (k < cnt ? k : 0) & 31);
            }
            float4 xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = ldg4(v + (int64_t)rb[u] * F + f);
            cp_async_wait<1>();                                       // this step's rows have landed (next step's may not)
            const float4* st = &sW[w][step & 1][0][lane];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (step * U + u < cnt) fma4(acc, xv[u], st[u * 32]);
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = ldg4(v + (int64_t)rb[4 + u] * F + f);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (step * U + 4 + u < cnt) fma4(acc, xv[u], st[(4 + u) * 32]);
        }
        cp_async_wait<0>();
    }
    *reinterpret_cast<float4*>(out + (int64_t)row * F + f) = acc;
}

// Negative results of round 2, kept out of the tree (numbers: profiles/r02_v5_launches.txt, profiles/r02_v6_cfconv_ab.txt):
//  * one cp.async.bulk (TMA) copy per 512-byte filter row instead of 32 x cp.async, to take the copies off the L1TEX pipe
//    (ncu: L1/TEX 61 % busy vs DRAM 44 %): 55 us instead of 35 us -- the bulk-copy engine serialises on ~3000 small copies
//    per SM;
//  * a persistent grid (3 CTAs per SM) with dynamic row hand-out (one atomic per row) and next-row index prefetch, to remove
//    the 15 % wave tail: 41.6 us instead of 37.5 us (61.7 vs 44.5 us with one filter row per edge) -- the dependent
//    atomic -> row pointer -> index chain costs more than the tail it removes.
// With one filter row per directed edge the kernel below streams 5.4 TB/s (83 % of the HBM copy peak); with one row per atom
// pair half of the row reads are L2 hits, DRAM traffic halves but the L2 -> SM path (the same 224 MB) now sets the time.

template <int F>
__global__ void __launch_bounds__(256)
cfconv_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ g, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ src, int n_atoms, float* __restrict__ dfilt) {
    constexpr int LPR = F / 4, EPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n_atoms) return;
    const int f = (lane % LPR) * 4, sub = lane / LPR;
    const int b = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
    const float4 gi = ldg4(g + (int64_t)row * F + f);
    for (int e = b + sub; e < end; e += EPW) {
        const int j = __ldg(src + e);
        const float4 xv = ldg4(x + (int64_t)j * F + f);
        st_stream4(dfilt + (int64_t)e * F + f, make_float4(xv.x * gi.x, xv.y * gi.y, xv.z * gi.z, xv.w * gi.w));
    }
}

template <bool TRANSPOSED>
static void launch_async(int blocks, cudaStream_t st, const float* filt, const int32_t* filt_row, const float* v, const int32_t* ptr,
                         const int32_t* idx_a, const int32_t* idx_b, int n, float* out) {
    constexpr int kSmem = 8 * 2 * 8 * 32 * 16;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaFuncSetAttribute(cfconv_gather_async_kernel<TRANSPOSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        configured.set();
    }
    launch_pdl(cfconv_gather_async_kernel<TRANSPOSED>, dim3(blocks), dim3(256), (size_t)kSmem, st, filt, filt_row, v, ptr, idx_a, idx_b, n, out);
}

// F = 128 takes the cp.async edition, narrower models the deep register edition; both accumulate in edge order and agree
// bit for bit.  (Round-1 A/B of four editions: profiles/r01_v26_tune_cfconv.txt, r01_v32_tune_cfconv.txt -- the kernel is
// latency bound, what pays is filter rows in flight, not less L2 traffic for the gather operand.)
template <int F>
int launch_fwd(const float* x, const float* filt, const int32_t* filt_row, const int32_t* rowptr, const int32_t* src, int64_t n,
               float* out, cudaStream_t st) {
    const int threads = 256;
    const int blocks = (int)((n * 32 + threads - 1) / threads);
    if constexpr (F == 128) launch_async<false>(blocks, st, filt, filt_row, x, rowptr, nullptr, src, (int)n, out);
    else launch_pdl(cfconv_gather_deep_kernel<F, false>, dim3(blocks), dim3(threads), 0, st, filt, filt_row, x, rowptr, nullptr, src, (int)n, out);
    return 0;
}
template <int F>
int launch_bwd_x(const float* filt, const int32_t* filt_row, const float* g, const int32_t* tr, const int32_t* te, const int32_t* tt,
                 int64_t n, float* dx, cudaStream_t st) {
    const int threads = 256;
    const int blocks = (int)((n * 32 + threads - 1) / threads);
    if constexpr (F == 128) launch_async<true>(blocks, st, filt, filt_row, g, tr, te, tt, (int)n, dx);
    else launch_pdl(cfconv_gather_deep_kernel<F, true>, dim3(blocks), dim3(threads), 0, st, filt, filt_row, g, tr, te, tt, (int)n, dx);
    return 0;
}
template <int F>
int launch_bwd_w(const float* x, const float* g, const int32_t* rowptr, const int32_t* src, int64_t n, float* dfilt, cudaStream_t st) {
    const int threads = 256;
    const int blocks = (int)((n * 32 + threads - 1) / threads);
    cfconv_bwd_w_kernel<F><<<blocks, threads, 0, st>>>(x, g, rowptr, src, (int)n, dfilt);
    return 0;
}

}  // namespace geossl

using namespace geossl;

#define DISPATCH_F(F, CALL)                                          \
    switch (F) {                                                     \
        case 32: { constexpr int kF = 32; CALL; break; }             \
        case 64: { constexpr int kF = 64; CALL; break; }             \
        case 128: { constexpr int kF = 128; CALL; break; }           \
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL; \
    }

extern "C" {

int geossl_cfconv_fwd(const float* x, const float* filt, const int32_t* filt_row, const int32_t* rowptr, const int32_t* src,
                      int64_t n_atoms, int F, float* out, void* stream) {
    if (n_atoms == 0) return 0;
    GEOSSL_REQUIRE(x && filt && rowptr && src && out && n_atoms > 0, "null pointer");
    DISPATCH_F(F, launch_fwd<kF>(x, filt, filt_row, rowptr, src, n_atoms, out, as_stream(stream)));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_cfconv_bwd_x(const float* filt, const int32_t* filt_row, const float* grad_out, const int32_t* t_rowptr, const int32_t* t_eid,
                        const int32_t* t_tgt, int64_t n_atoms, int F, float* grad_x, void* stream) {
    if (n_atoms == 0) return 0;
    GEOSSL_REQUIRE(filt && grad_out && t_rowptr && t_eid && t_tgt && grad_x && n_atoms > 0, "null pointer");
    DISPATCH_F(F, launch_bwd_x<kF>(filt, filt_row, grad_out, t_rowptr, t_eid, t_tgt, n_atoms, grad_x, as_stream(stream)));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_cfconv_bwd_w(const float* x, const float* grad_out, const int32_t* rowptr, const int32_t* src,
                        int64_t n_atoms, int F, float* grad_filt, void* stream) {
    if (n_atoms == 0) return 0;
    GEOSSL_REQUIRE(x && grad_out && rowptr && src && grad_filt && n_atoms > 0, "null pointer");
    DISPATCH_F(F, launch_bwd_w<kF>(x, grad_out, rowptr, src, n_atoms, grad_filt, as_stream(stream)));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
