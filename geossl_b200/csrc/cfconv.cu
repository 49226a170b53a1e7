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
#include <type_traits>

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

// Pair-centric edition for small graphs (molecules): with one filter row per atom PAIR, the row-gather kernels above
// still move every filter row L2 -> SM twice (once per direction), and that path, not DRAM, sets their time.  Here one
// CTA owns one graph, keeps the graph's operand rows v and its accumulators in shared memory, and streams the graph's
// contiguous block of filter rows ONCE: for pair u = (s, t) it adds W_u * v[s] to row t and W_u * v[t] to row s.
// G warps split the block into G contiguous chunks; a warp owns a private accumulator copy (lane = 4 features), so there
// are no atomics and the summation order is fixed: results are run-to-run identical (they differ from the row-gather
// editions by fp32 summation order only).  The pairs of a chunk are sorted by their owner row t, whose running sum stays
// in registers and is flushed on a row change; 2 x UNR filter rows per warp are in flight in registers (ping-pong).
// Orphan pairs (reverse direction cut by the neighbour limit, pair_atoms = (s, ~t)) contribute to one side only:
// forward to row t, transposed (d/dx) to row s.  That is folded into the pair's shared-memory ROW numbers once per 32
// pairs (each lane decodes one pair, the packed word is broadcast by shuffle): the missing side reads v from an all-zero
// row / accumulates into a dump row, so the inner loop has no selects.  A graph with more than n_max atoms (host-side
// bound wrong, or the edge-less padding graph of data.pad_batch) is still summed correctly by one warp accumulating
// straight into `out` (pair_chunk_global).
// fire-and-forget L2 prefetch of a contiguous block (no destination register, so no scoreboard): bytes % 16 == 0
__device__ __forceinline__ void prefetch_l2_bulk(const void* gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }

template <bool TRANSPOSED>
__device__ __noinline__ void pair_chunk_global(const float* __restrict__ filt, const int2* __restrict__ pair_atoms, int lo, int hi,
                                               int a0, int lane, float4* acc, const float4* vv) {
    // acc / vv: this lane's float4 column of row 0 of the graph in global memory; rows are 32 float4 apart
    for (int u = lo; u < hi; ++u) {
        const int2 st = __ldg(pair_atoms + u);
        const bool both = st.y >= 0;
        const int t = (both ? st.y : ~st.y) - a0, s = st.x - a0;
        const float4 w = ld_stream4(filt + (int64_t)u * 128 + lane * 4);
        if (!TRANSPOSED || both) { float4 a = acc[t * 32]; fma4(a, vv[s * 32], w); acc[t * 32] = a; }
        if (TRANSPOSED || both) { float4 a = acc[s * 32]; fma4(a, vv[t * 32], w); acc[s * 32] = a; }
    }
}

// One warp's state while it walks its chunk of a graph's pair block.  Pairs are consumed four at a time: when all four
// belong to the running row t (the common case, rows hold ~n/2 pairs) their accumulator rows are distinct, so the four
// shared-memory read-modify-writes are issued together instead of as one dependent chain; a group that crosses a row
// boundary takes the pair-by-pair path.
struct PairWalk {
    uint32_t cur_t = 0xffffffffu;
    float4 acc_t = make_float4(0.f, 0.f, 0.f, 0.f), v_t = make_float4(0.f, 0.f, 0.f, 0.f);

    __device__ __forceinline__ void flush(float4* acc) {
        if (cur_t != 0xffffffffu) { float4 a = acc[cur_t * 32]; add4(a, acc_t); acc[cur_t * 32] = a; }
    }
    __device__ __forceinline__ void one(const float4& wk, uint32_t m, float4* acc, const float4* vv) {
        const uint32_t t = m & 0xffu;
        if (t != cur_t) {
            flush(acc);
            cur_t = t;
            acc_t = make_float4(0.f, 0.f, 0.f, 0.f);
            v_t = vv[t * 32];
        }
        fma4(acc_t, vv[((m >> 8) & 0xffu) * 32], wk);
        float4* ap = acc + (m >> 16) * 32;
        float4 as = *ap;
        fma4(as, v_t, wk);
        *ap = as;
    }
    __device__ __forceinline__ void change_row(uint32_t t, float4* acc, const float4* vv) {
        flush(acc);
        cur_t = t;
        acc_t = make_float4(0.f, 0.f, 0.f, 0.f);
        v_t = vv[t * 32];
    }
    __device__ __forceinline__ void four(const float4& w0, const float4& w1, const float4& w2, const float4& w3, uint4 m,
                                         float4* acc, const float4* vv) {
        const uint32_t r0 = m.x >> 16, r1 = m.y >> 16, r2 = m.z >> 16, r3 = m.w >> 16;
        float4* p0 = acc + r0 * 32;
        float4* p1 = acc + r1 * 32;
        float4* p2 = acc + r2 * 32;
        float4* p3 = acc + r3 * 32;
        const float4 x0 = vv[((m.x >> 8) & 0xffu) * 32], x1 = vv[((m.y >> 8) & 0xffu) * 32];
        const float4 x2 = vv[((m.z >> 8) & 0xffu) * 32], x3 = vv[((m.w >> 8) & 0xffu) * 32];
        if ((((m.x ^ m.y) | (m.x ^ m.z) | (m.x ^ m.w)) & 0xffu) == 0u) {
            // one owner row: the four accumulator rows are distinct (or the dump row), their updates are independent
            if ((m.x & 0xffu) != cur_t) change_row(m.x & 0xffu, acc, vv);
            float4 a0 = *p0, a1 = *p1, a2 = *p2, a3 = *p3;
            fma4(acc_t, x0, w0); fma4(acc_t, x1, w1); fma4(acc_t, x2, w2); fma4(acc_t, x3, w3);
            fma4(a0, v_t, w0); fma4(a1, v_t, w1); fma4(a2, v_t, w2); fma4(a3, v_t, w3);
            *p0 = a0; *p1 = a1; *p2 = a2; *p3 = a3;
        } else if (r0 != r1 && r0 != r2 && r0 != r3 && r1 != r2 && r1 != r3 && r2 != r3) {
            // the group crosses a row boundary but still updates four different accumulator rows: same independent
            // updates with each pair's own v[t]; the owner-row sums follow in order (a flush is a later read-modify-write)
            const float4 t0 = vv[(m.x & 0xffu) * 32], t1 = vv[(m.y & 0xffu) * 32];
            const float4 t2 = vv[(m.z & 0xffu) * 32], t3 = vv[(m.w & 0xffu) * 32];
            float4 a0 = *p0, a1 = *p1, a2 = *p2, a3 = *p3;
            fma4(a0, t0, w0); fma4(a1, t1, w1); fma4(a2, t2, w2); fma4(a3, t3, w3);
            *p0 = a0; *p1 = a1; *p2 = a2; *p3 = a3;
            if ((m.x & 0xffu) != cur_t) change_row(m.x & 0xffu, acc, vv);
            fma4(acc_t, x0, w0);
            if ((m.y & 0xffu) != cur_t) change_row(m.y & 0xffu, acc, vv);
            fma4(acc_t, x1, w1);
            if ((m.z & 0xffu) != cur_t) change_row(m.z & 0xffu, acc, vv);
            fma4(acc_t, x2, w2);
            if ((m.w & 0xffu) != cur_t) change_row(m.w & 0xffu, acc, vv);
            fma4(acc_t, x3, w3);
        } else {
            one(w0, m.x, acc, vv); one(w1, m.y, acc, vv); one(w2, m.z, acc, vv); one(w3, m.w, acc, vv);
        }
    }
};

// Shared memory: v rows [n_max + 1][32] float4 (last row: zeros) | G accumulator copies [n_max + 1][32] (last row: dump)
// | meta_cap packed pair records (t | row of v[s] << 8 | row gaining W v[t] << 16), decoded once by the whole CTA.
// The only long-latency loads inside the loop are the filter rows, in two register buffers of UNR rows that are waited
// for as a whole (one scoreboard each): a load that shares a scoreboard with younger loads would wait for those too.
template <int G, int UNR, bool TRANSPOSED>
__global__ void __launch_bounds__(32 * G)
cfconv_pairs_kernel(const float* __restrict__ filt, const int2* __restrict__ pair_atoms, const int32_t* __restrict__ pair_rowptr,
                    const int32_t* __restrict__ graph_ptr, const float* __restrict__ v, int n_max, int meta_cap,
                    float* __restrict__ out) {
    extern __shared__ float4 sP[];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);            // provably warp-uniform
    const int a0 = __ldg(graph_ptr + blockIdx.x), n = __ldg(graph_ptr + blockIdx.x + 1) - a0;
    if (n <= 0) return;
    const int p0 = __ldg(pair_rowptr + a0), p1 = __ldg(pair_rowptr + a0 + n);
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n > n_max || p1 - p0 > meta_cap) {
        if (wid == 0) {
            float4* o = reinterpret_cast<float4*>(out + (int64_t)a0 * 128) + lane;
            for (int i = 0; i < n; ++i) o[i * 32] = z4;
            if (p1 > p0)
                pair_chunk_global<TRANSPOSED>(filt, pair_atoms, p0, p1, a0, lane, o,
                                              reinterpret_cast<const float4*>(v + (int64_t)a0 * 128) + lane);
        }
        return;
    }
    const int per = (((p1 - p0 + G - 1) / G) + 3) & ~3;                // multiple of 4: the records of a group are one 16-byte load
    const int lo = min(p0 + wid * per, p1), hi = min(lo + per, p1);
    // ---- the first two buffers of filter rows are requested before anything else
    float4 wa[UNR], wb[UNR];
    auto load = [&](float4 (&w)[UNR], int u0) {
#pragma unroll
        for (int k = 0; k < UNR; ++k)     // plain loads: inline-asm loads all land on ONE scoreboard, and a wait on it drains both buffers
            w[k] = (u0 + k < hi) ? __ldcs(reinterpret_cast<const float4*>(filt + (int64_t)(u0 + k) * 128) + lane) : z4;
    };
    constexpr int PD = 32 / UNR;                                      // prefetch distance in buffers: 16 KB per warp
    auto prefetch = [&](int u0, int count) {                          // rows [u0, u0 + count) clipped to the chunk -> L2
        const int end = min(u0 + count, hi);
        if (lane == 0 && end > u0) prefetch_l2_bulk(filt + (int64_t)u0 * 128, (uint32_t)(end - u0) * 512u);
    };
    load(wa, lo);
    load(wb, lo + UNR);
    prefetch(lo + 2 * UNR, PD * UNR);
    // ---- stage the graph's operand rows and pair records (every load is issued before the first store), clear the accumulators
    const int rows = n_max + 1;
    float4* sv = sP;
    float4* sacc = sP + rows * 32;
    uint32_t* smeta = reinterpret_cast<uint32_t*>(sacc + G * rows * 32);
    const float4* vg = reinterpret_cast<const float4*>(v + (int64_t)a0 * 128);
    constexpr int VR = 32 / G, MR = 16 / G;                           // per round: 32 rows of v, 512 pair records
    for (int i0 = 0, j0 = 0; i0 < n * 32 || j0 < p1 - p0; i0 += 32 * G * VR, j0 += 32 * G * MR) {
        float4 tv[VR];
        int2 tm[MR];
#pragma unroll
        for (int j = 0; j < VR; ++j) { const int i = i0 + j * 32 * G + tid; tv[j] = (i < n * 32) ? __ldg(vg + i) : z4; }
#pragma unroll
        for (int j = 0; j < MR; ++j) { const int i = j0 + j * 32 * G + tid; tm[j] = (i < p1 - p0) ? __ldg(pair_atoms + p0 + i) : make_int2(a0, a0); }
#pragma unroll
        for (int j = 0; j < VR; ++j) { const int i = i0 + j * 32 * G + tid; if (i < n * 32) sv[i] = tv[j]; }
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            const int i = j0 + j * 32 * G + tid;
            if (i < p1 - p0) {
                const bool both = tm[j].y >= 0;
                const uint32_t t = (uint32_t)((both ? tm[j].y : ~tm[j].y) - a0), s_ = (uint32_t)(tm[j].x - a0);
                const uint32_t v_row = (TRANSPOSED && !both) ? (uint32_t)n_max : s_;     // missing side: zero operand row ...
                const uint32_t a_row = (!TRANSPOSED && !both) ? (uint32_t)n_max : s_;    // ... or the dump accumulator row
                smeta[i] = t | (v_row << 8) | (a_row << 16);
            }
        }
    }
    for (int i = tid; i < n * 32; i += 32 * G) {
#pragma unroll
        for (int c = 0; c < G; ++c) sacc[c * rows * 32 + i] = z4;
    }
    if (tid < 32) sv[n_max * 32 + tid] = z4;
    __syncthreads();
    float4* acc = sacc + wid * rows * 32 + lane;
    const float4* vv = sv + lane;
    const uint32_t* sm = smeta - p0;
    PairWalk walk;
    auto consume = [&](const float4 (&w)[UNR], int u0) {
        if (u0 + UNR <= hi) {
#pragma unroll
            for (int k = 0; k < UNR; k += 4)
                walk.four(w[k], w[k + 1], w[k + 2], w[k + 3], *reinterpret_cast<const uint4*>(sm + u0 + k), acc, vv);
        } else {
#pragma unroll
            for (int k = 0; k < UNR; ++k)
                if (u0 + k < hi) walk.one(w[k], sm[u0 + k], acc, vv);
        }
    };
    for (int u0 = lo; u0 < hi; u0 += 2 * UNR) {
        consume(wa, u0);
        load(wa, u0 + 2 * UNR);
        if (u0 + UNR < hi) consume(wb, u0 + UNR);
        load(wb, u0 + 3 * UNR);
        prefetch(u0 + (2 + PD) * UNR, 2 * UNR);
    }
    walk.flush(acc);
    __syncthreads();
    float4* og = reinterpret_cast<float4*>(out + (int64_t)a0 * 128);
    for (int i = tid; i < n * 32; i += 32 * G) {
        float4 s = sacc[i];
#pragma unroll
        for (int c = 1; c < G; ++c) add4(s, sacc[c * rows * 32 + i]);
        og[i] = s;
    }
}

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

// Edge product of the shared-row aggregate: dW_u = x[s] * g[t] (+ x[t] * g[s] when the reverse direction exists), one row
// per atom pair u = (s, t) -- the filter gradient of geossl_cfconv_fwd with filt_row = pair_of_edge, materialised (U,F) for
// the composed double-backward path.
template <int F>
__global__ void __launch_bounds__(256)
cfconv_pair_product_kernel(const float* __restrict__ x, const float* __restrict__ g, const int2* __restrict__ pair_atoms,
                           int64_t n_pairs, float* __restrict__ dfilt) {
    constexpr int LPR = F / 4, EPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int f = (lane % LPR) * 4, sub = lane / LPR;
    const int64_t u = warp * EPW + sub;
    if (u >= n_pairs) return;
    const int2 st = __ldg(pair_atoms + u);
    const bool both = st.y >= 0;
    const int s = st.x, t = both ? st.y : ~st.y;
    const float4 xs = ldg4(x + (int64_t)s * F + f), gt = ldg4(g + (int64_t)t * F + f);
    float4 r = make_float4(xs.x * gt.x, xs.y * gt.y, xs.z * gt.z, xs.w * gt.w);
    if (both) {
        const float4 xt = ldg4(x + (int64_t)t * F + f), gs = ldg4(g + (int64_t)s * F + f);
        fma4(r, xt, gs);
    }
    st_stream4(dfilt + u * F + f, r);
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
inline int pairs_meta_cap(int n_max) {                              // pair records per graph: <= 32 neighbours per row, <= n(n-1)
    const int cap = n_max * 32 < n_max * (n_max - 1) ? n_max * 32 : n_max * (n_max - 1);
    return (cap + 7) & ~3;
}
inline size_t pairs_smem(int groups, int n_max) { return (size_t)(1 + groups) * (n_max + 1) * 512 + 4 * (size_t)pairs_meta_cap(n_max); }

template <int G, int UNR>
int launch_pairs(const float* filt, const int2* pair_atoms, const int32_t* pair_rowptr, const int32_t* graph_ptr, int n_graphs,
                 const float* v, int n_max, bool transposed, float* out, cudaStream_t st) {
    const size_t smem = pairs_smem(G, n_max);
    const int meta_cap = pairs_meta_cap(n_max);
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaFuncSetAttribute(cfconv_pairs_kernel<G, UNR, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(cfconv_pairs_kernel<G, UNR, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured.set();
    }
    if (transposed)
        launch_pdl(cfconv_pairs_kernel<G, UNR, true>, dim3(n_graphs), dim3(32 * G), smem, st, filt, pair_atoms, pair_rowptr, graph_ptr, v, n_max, meta_cap, out);
    else
        launch_pdl(cfconv_pairs_kernel<G, UNR, false>, dim3(n_graphs), dim3(32 * G), smem, st, filt, pair_atoms, pair_rowptr, graph_ptr, v, n_max, meta_cap, out);
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

int geossl_cfconv_pair_product(const float* x, const float* grad_out, const int32_t* pair_atoms, int64_t n_pairs, int F,
                               float* grad_filt, void* stream) {
    if (n_pairs == 0) return 0;
    GEOSSL_REQUIRE(x && grad_out && pair_atoms && grad_filt && n_pairs > 0, "null pointer");
    GEOSSL_REQUIRE((reinterpret_cast<uintptr_t>(pair_atoms) & 7) == 0, "pair_atoms must be 8-byte aligned");
    const int2* pa = reinterpret_cast<const int2*>(pair_atoms);
    cudaStream_t st = as_stream(stream);
    switch (F) {
        case 32: cfconv_pair_product_kernel<32><<<(int)((n_pairs * 8 + 255) / 256), 256, 0, st>>>(x, grad_out, pa, n_pairs, grad_filt); break;
        case 64: cfconv_pair_product_kernel<64><<<(int)((n_pairs * 16 + 255) / 256), 256, 0, st>>>(x, grad_out, pa, n_pairs, grad_filt); break;
        case 128: cfconv_pair_product_kernel<128><<<(int)((n_pairs * 32 + 255) / 256), 256, 0, st>>>(x, grad_out, pa, n_pairs, grad_filt); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_cfconv_pairs_max_atoms(int groups) {
    if (groups < 1) groups = 2;
    int n = 254;                                                     // row numbers travel as bytes
    while (n > 1 && pairs_smem(groups, n) > 226 * 1024) --n;
    return n;
}

int geossl_cfconv_pairs(const float* v, const float* filt, const int32_t* pair_atoms, const int32_t* pair_rowptr,
                        const int32_t* graph_ptr, int64_t n_graphs, int max_graph_atoms, int transposed, int tuning,
                        float* out, void* stream) {
    if (n_graphs == 0) return 0;
    GEOSSL_REQUIRE(v && filt && pair_atoms && pair_rowptr && graph_ptr && out && n_graphs > 0, "null pointer");
    GEOSSL_REQUIRE((reinterpret_cast<uintptr_t>(pair_atoms) & 7) == 0, "pair_atoms must be 8-byte aligned");
    if (tuning <= 0) tuning = 208;
    const int groups = tuning / 100, unroll = tuning % 100;
    GEOSSL_REQUIRE(max_graph_atoms >= 1 && max_graph_atoms <= geossl_cfconv_pairs_max_atoms(groups),
                   "max_graph_atoms exceeds the shared-memory window (use geossl_cfconv_fwd / geossl_cfconv_bwd_x)");
    const int2* pa = reinterpret_cast<const int2*>(pair_atoms);
    cudaStream_t st = as_stream(stream);
#define PAIRS_CASE(G_, U_) case G_ * 100 + U_: launch_pairs<G_, U_>(filt, pa, pair_rowptr, graph_ptr, (int)n_graphs, v, max_graph_atoms, transposed != 0, out, st); break;
    switch (groups * 100 + unroll) {
        PAIRS_CASE(1, 16) PAIRS_CASE(2, 8) PAIRS_CASE(2, 16) PAIRS_CASE(3, 8) PAIRS_CASE(3, 16) PAIRS_CASE(4, 8) PAIRS_CASE(4, 16)
        default: set_error("%s: unsupported tuning %d (groups*100 + unroll)", __func__, tuning); return GEOSSL_EINVAL;
    }
#undef PAIRS_CASE
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
