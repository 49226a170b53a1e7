// SchNet filter network on the sm_100a tensor cores (tcgen05 + TMEM), forward.
//
//   d_e -> rbf (64-wide, zero padded) --MMA1--> a = rbf W1^T  --epilogue--> s = ssp(a + b1)
//       --MMA2--> u = s W2^T --epilogue--> W_e = (u + b2) * cutoff(d_e)  -> filt (E,128) in HBM
//
// Replaces GaussianSmearing.forward (Geom3D/models/schnet.py:205-207), InteractionBlock.mlp
// (schnet.py:141-145) and the cutoff product of CFConv.forward (schnet.py:186-187).  F = 128, G <= 64.
//
// One persistent CTA per SM, 128-edge tiles, 25 warps:
//   warps 0-15  epilogue (quadrant = warp & 3, 32 accumulator columns per warp = warp >> 2):
//               E1 reads D1 (lanes = edges) and writes the s tile with 16-byte row stores;
//               E2 reads D2^T (lanes = FEATURES, columns = edges -- MMA2 is issued with the operands swapped),
//               so every store instruction writes 128 contiguous bytes of one W_e row
//   warps 16-23 producer: rbf tile (hi/lo, K-major SW128) for the next tile, double buffered
//   warp 24     TMEM allocation + the single MMA-issuing thread
// Measured structure (profiles/r01_v4_trace_fwd.txt -> v5): the epilogue is MUFU bound (2 MUFU per softplus),
// the tensor pipe needs ~2.3 k cycles per tile, uncoalesced row-per-thread stores cost 4 k LSU cycles per tile.
// Operands never come from HBM (they are computed on chip), so tiles are written with st.shared in the
// swizzled layout and published to the async proxy with fence.proxy.async; W1 is split once per CTA into shared
// memory, W2 into TENSOR memory (A operand of MMA2, which frees 64 KB of shared memory for a second s buffer and halves
// MMA2's shared-memory operand traffic).  TMEM: D1 x2 | D2^T | W2 hi | W2 lo = 512 columns.
// fp32 operands are split into two 16-bit parts and every product is three MMAs (see tc.cuh).
#include "common.cuh"
#include "tc.cuh"

namespace geossl {
namespace tc {

// Optional per-phase clock64() trace of CTA 0 (set with geossl_debug_set_trace); slot = tile*16 + event.
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace(int tile, int event) {
    long long* t = g_trace;
    if (t != nullptr && blockIdx.x == 0 && tile < 32 && (threadIdx.x & 31) == 0) t[tile * 16 + event] = clock64();
}

constexpr int kF = 128;            // filters
constexpr int kTile = 128;         // edges per tile (UMMA M)
constexpr int kBlk = kTile * 128;  // bytes of one [128 rows x 64 k] 16-bit block
constexpr int kEpiWarps = 16, kEpiThreads = kEpiWarps * 32, kProdWarps = 8, kProdThreads = kProdWarps * 32;
constexpr int kThreads = kEpiThreads + kProdThreads + 32;
constexpr int kProdWarp0 = kEpiWarps, kMmaWarp = kEpiWarps + kProdWarps;

struct FwdLayout {                 // byte offsets from the 1024-aligned dynamic smem base
    static constexpr int W1_hi = 0, W1_lo = W1_hi + kBlk;                 // [128 f][64 g]
    static constexpr int PHI = W1_lo + kBlk;                              // 2 buffers x (hi, lo)
    static constexpr int S = PHI + 4 * kBlk;                              // 2 buffers x (hi: 2 k-blocks, lo: 2 k-blocks)
    static constexpr int B1 = S + 8 * kBlk;                               // 128 floats
    static constexpr int B2 = B1 + 512;
    static constexpr int OFF = B2 + 512;                                  // 64 floats
    static constexpr int BAR = OFF + 256;                                 // 16 mbarriers
    static constexpr int TMEM_PTR = BAR + 16 * 8;
    static constexpr int kBytes = TMEM_PTR + 16;
};
static_assert(FwdLayout::kBytes + 1024 <= 227 * 1024, "shared memory budget");

enum Bar { PHI_FULL = 0, PHI_EMPTY = 2, D1_FULL = 4, D1_EMPTY = 6, S_FULL = 8, S_EMPTY = 10, D2_FULL = 12, D2_EMPTY = 13 };

template <bool FP16>
__global__ void __launch_bounds__(kThreads, 1)
filter_fwd_tc_kernel(const float* __restrict__ edge_dist, const int32_t* __restrict__ n_edges_dev, int64_t capacity,
                     const float* __restrict__ offset, float coeff, float cutoff, int G,
                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                     const float* __restrict__ b2, float* __restrict__ filt) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    using L = FwdLayout;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + L::BAR;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    float* sB1 = reinterpret_cast<float*>(smem + L::B1);
    float* sB2 = reinterpret_cast<float*>(smem + L::B2);
    float* sOff = reinterpret_cast<float*>(smem + L::OFF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    pdl_wait();

    int64_t n_edges = (int64_t)(*n_edges_dev);
    if (n_edges > capacity) n_edges = capacity;
    const int64_t n_tiles = (n_edges + kTile - 1) / kTile;
    const int my_tiles = (blockIdx.x < n_tiles) ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;

    // ---- one-time setup: weights (split, swizzled), biases, barriers, TMEM
    for (int idx = tid; idx < kF * 8; idx += kThreads) {             // W1: row f, chunk c (8 g each), zero padded to 64
        const int f = idx >> 3, c = idx & 7;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int g = c * 8 + j; x[j] = (g < G) ? __ldg(w1 + f * G + g) : 0.f; }
        store_chunk8<FP16>(smem + L::W1_hi, smem + L::W1_lo, f, c * 8, x);
    }
    if (tid < kF) { sB1[tid] = __ldg(b1 + tid); sB2[tid] = __ldg(b2 + tid); }
    if (tid < 64) sOff[tid] = (tid < G) ? __ldg(offset + tid) : 1e18f;   // padded gaussians: exp(coeff * 1e36) == 0, no branch
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(PHI_FULL + b), kProdThreads / 32);
            mbar_init(bar(PHI_EMPTY + b), 1);
            mbar_init(bar(D1_FULL + b), 1);
            mbar_init(bar(D1_EMPTY + b), kEpiThreads / 32);
            mbar_init(bar(S_FULL + b), kEpiThreads / 32);
            mbar_init(bar(S_EMPTY + b), 1);
        }
        mbar_init(bar(D2_FULL), 1);
        mbar_init(bar(D2_EMPTY), kEpiThreads / 32);
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(sbase + L::TMEM_PTR, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    // tensor memory: D1 x2 | D2^T | W2 hi | W2 lo   (W2 resident: MMA2 reads its A operand from TMEM, only the s tile from smem)
    const uint32_t tD1[2] = {tmem, tmem + 128}, tD2 = tmem + 256, tW2h = tmem + 384, tW2l = tmem + 448;
    if (warp < kEpiWarps) {
        const int q = warp & 3, part = warp >> 2, o = q * 32 + lane;           // lane o owns row o of W2, 32 k per warp
        const uint32_t lane_b = (uint32_t)(q * 32) << 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k0 = part * 32 + h * 16;
            float x[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 t = ldg4(w2 + o * kF + k0 + j);
                x[j] = t.x; x[j + 1] = t.y; x[j + 2] = t.z; x[j + 3] = t.w;
            }
            tmem_store_split16<FP16>(tW2h + lane_b + k0 / 2, tW2l + lane_b + k0 / 2, x);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= kProdWarp0 && warp < kMmaWarp) {
        // ===================== producer: rbf tile of local tile i into PHI[i & 1]; thread = (row, half of the 64 columns)
        const int tp = tid - kEpiThreads;
        const int r = tp & 127, c0 = (tp >> 7) * 4;
        const float cl2 = coeff * 1.4426950408889634f;                // exp(coeff * x) = 2^(cl2 * x)
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t e = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kTile + r;
            const int b = i & 1;
            mbar_wait(bar(PHI_EMPTY + b), ((i >> 1) & 1) ^ 1);
            if (warp == kProdWarp0) trace(i, 0);
            const float d = (e < n_edges) ? __ldg(edge_dist + e) : 0.f;
            uint8_t* hi = smem + L::PHI + b * 2 * kBlk;
            uint8_t* lo = hi + kBlk;
#pragma unroll
            for (int c = c0; c < c0 + 4; ++c) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float diff = d - sOff[c * 8 + j];
                    x[j] = ex2_approx(cl2 * (diff * diff));
                }
                store_chunk8<FP16>(hi, lo, r, c * 8, x);
            }
            fence_proxy_async();
            warp_arrive(bar(PHI_FULL + b));
            if (warp == kProdWarp0) trace(i, 1);
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer (whole warp runs the loop; the elected lane issues)
        {
            const bool leader = elect_one_sync();
            const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, kTile, kF);
            const uint64_t dW1h = desc_k_sw128(sbase + L::W1_hi), dW1l = desc_k_sw128(sbase + L::W1_lo);
            auto issue_mma1 = [&](int i) {
                const int b = i & 1;
                mbar_wait(bar(PHI_FULL + b), (i >> 1) & 1);
                mbar_wait(bar(D1_EMPTY + b), ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                trace(i, 2);
                const uint64_t dPh = desc_k_sw128(sbase + L::PHI + b * 2 * kBlk), dPl = desc_k_sw128(sbase + L::PHI + b * 2 * kBlk + kBlk);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)                     // K = 64 = 4 x 16 (+32 bytes = +2 encoded)
                        mma3(tD1[b], dPh + 2 * kk, dPl + 2 * kk, dW1h + 2 * kk, dW1l + 2 * kk, idesc, kk > 0);
                    tc_commit(bar(PHI_EMPTY + b));
                    tc_commit(bar(D1_FULL + b));
                }
                __syncwarp();
                trace(i, 3);
            };
            if (my_tiles > 0) issue_mma1(0);
            for (int i = 0; i < my_tiles; ++i) {
                if (i + 1 < my_tiles) issue_mma1(i + 1);
                const int b = i & 1;
                mbar_wait(bar(S_FULL + b), (i >> 1) & 1);
                mbar_wait(bar(D2_EMPTY), (i & 1) ^ 1);
                const uint64_t dSh = desc_k_sw128(sbase + L::S + b * 4 * kBlk), dSl = desc_k_sw128(sbase + L::S + b * 4 * kBlk + 2 * kBlk);
                tc_fence_after();
                trace(i, 4);
                if (leader) {
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)                     // K = 128 = 2 blocks x 4 x 16
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t o = kb * (kBlk >> 4) + 2 * kk, ka = (kb * 4 + kk) * 8;
                            mma3_ts(tD2, tW2h + ka, tW2l + ka, dSh + o, dSl + o, idesc, (kb | kk) > 0);   // D2^T = W2 . S^T
                        }
                    tc_commit(bar(S_EMPTY + b));
                    tc_commit(bar(D2_FULL));
                }
                __syncwarp();
                trace(i, 5);
            }
        }
    } else {
        // ===================== epilogue warps: E1(i) then E2(i-1); warp = (lane quadrant q, column quarter cq)
        const int q = warp & 3, cq = warp >> 2;
        const int r = q * 32 + lane;                                  // TMEM lane: edge row in E1, FEATURE in E2
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int col0 = cq * 32;                                     // this warp's 32 accumulator columns
        const float b2f = sB2[r];
        for (int i = 0; i <= my_tiles; ++i) {
            // E2(i-1)'s edge distance is fetched before E1(i) so its global-load latency hides behind the softplus phase
            const int64_t e0 = ((int64_t)blockIdx.x + (int64_t)(i - 1) * gridDim.x) * kTile + col0;
            float dpre = 0.f;
            if (i > 0 && e0 + lane < n_edges) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dpre) : "l"(edge_dist + e0 + lane));
            if (i < my_tiles) {
                const int b = i & 1;
                mbar_wait(bar(D1_FULL + b), (i >> 1) & 1);
                tc_fence_after();
                if (warp == 0) trace(i, 6);
                float v[32];
                tmem_ld32(tD1[b] + lane_base + col0, v);
                tc_fence_before();
                warp_arrive(bar(D1_EMPTY + b));
                if (warp == 0) trace(i, 7);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ssp_fast(v[j] + sB1[col0 + j]);
                if (warp == 0) trace(i, 8);
                mbar_wait(bar(S_EMPTY + b), ((i >> 1) & 1) ^ 1);       // MMA2 of tile i-2 released this S buffer
                if (warp == 0) trace(i, 9);
                uint8_t* hi = smem + L::S + b * 4 * kBlk + (cq >> 1) * kBlk;   // k-block holds columns 64*(cq>>1)..+63
                uint8_t* lo = hi + 2 * kBlk;
#pragma unroll
                for (int c = 0; c < 4; ++c) store_chunk8<FP16>(hi, lo, r, (cq & 1) * 32 + c * 8, &v[c * 8]);
                fence_proxy_async();
                warp_arrive(bar(S_FULL + b));
                if (warp == 0) trace(i, 10);
            }
            if (i > 0) {
                const int t = i - 1;
                // lane j holds the cutoff of edge column j of this warp; broadcast by shuffle in the store loop
                const float myc = (e0 + lane < n_edges) ? 0.5f * (__cosf(dpre * (kPi / cutoff)) + 1.0f) : 0.f;
                mbar_wait(bar(D2_FULL), t & 1);
                tc_fence_after();
                if (warp == 0) trace(t, 11);
                float v[32];
                tmem_ld32(tD2 + lane_base + col0, v);                  // lane = feature r, columns = edges col0..col0+31
                tc_fence_before();
                warp_arrive(bar(D2_EMPTY));
                float* out = filt + e0 * kF + r;
                if (e0 + 32 <= n_edges) {                              // full column group: no per-edge bounds checks
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        out[j * kF] = (v[j] + b2f) * __shfl_sync(0xffffffffu, myc, j);   // 32 lanes -> 128 contiguous bytes
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float c = __shfl_sync(0xffffffffu, myc, j);
                        if (e0 + j < n_edges) out[(int64_t)j * kF] = (v[j] + b2f) * c;
                    }
                }
                if (warp == 0) trace(t, 12);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        __syncwarp();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------ self test
// mode 0: D[m][n] = sum_k A[m][k] B[n][k]        A (128,K), B (128,K) row-major fp32, K in {64,128}  (K-major operands)
// mode 1: D[m][n] = sum_k X[k][m] Y[k][n]        X (128,128), Y (128,N) row-major fp32, N in {64,128} (MN-major operands)
template <bool FP16>
__global__ void __launch_bounds__(128, 1)
tc_selftest_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B, int K, int N, float* __restrict__ D) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const uint32_t sbase = smem_u32(smem);
    // A_hi | A_lo | B_hi | B_lo, each 2 blocks of 16 KB; barrier + tmem ptr after
    uint8_t* Ah = smem; uint8_t* Al = smem + 2 * kBlk; uint8_t* Bh = smem + 4 * kBlk; uint8_t* Bl = smem + 6 * kBlk;
    const uint32_t bar = sbase + 8 * kBlk, tptr = bar + 8;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (mode != 1) {
        for (int c = 0; c < K / 8; ++c) {                            // thread = row
            store_chunk8<FP16>(Ah + (c >> 3) * kBlk, Al + (c >> 3) * kBlk, tid, (c & 7) * 8, A + tid * K + c * 8);
            store_chunk8<FP16>(Bh + (c >> 3) * kBlk, Bl + (c >> 3) * kBlk, tid, (c & 7) * 8, B + tid * K + c * 8);
        }
    } else {
        for (int c = 0; c < 16; ++c)                                 // thread = k row; X has 128 m = 2 MN blocks
            store_chunk8<FP16>(Ah + (c >> 3) * kBlk, Al + (c >> 3) * kBlk, tid, (c & 7) * 8, A + tid * 128 + c * 8);
        for (int c = 0; c < N / 8; ++c)
            store_chunk8<FP16>(Bh + (c >> 3) * kBlk, Bl + (c >> 3) * kBlk, tid, (c & 7) * 8, B + tid * N + c * 8);
    }
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tptr, 256);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 8 * kBlk + 8);
    if (mode == 3) {                                                 // A (hi | lo) into TMEM columns [128,128+K/2) | [192,192+K/2)
        const uint32_t lane_b = (uint32_t)(warp * 32) << 16;
        for (int k0 = 0; k0 < K; k0 += 16) tmem_store_split16<FP16>(tmem + lane_b + 128 + k0 / 2, tmem + lane_b + 192 + k0 / 2, A + tid * K + k0);
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 0 && elect_one_sync()) {                             // uniform operands; the elected lane issues
        if (mode == 0) {
            const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, 128);
            const uint64_t ah = desc_k_sw128(sbase), al = desc_k_sw128(sbase + 2 * kBlk);
            const uint64_t bh = desc_k_sw128(sbase + 4 * kBlk), bl = desc_k_sw128(sbase + 6 * kBlk);
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t o = (ks >> 2) * (kBlk >> 4) + 2 * (ks & 3);
                mma3(tmem, ah + o, al + o, bh + o, bl + o, idesc, ks > 0);
            }
        } else if (mode == 3) {
            const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, N);
            const uint64_t bh = desc_k_sw128(sbase + 4 * kBlk), bl = desc_k_sw128(sbase + 6 * kBlk);
            const long long t0 = clock64();
            for (int ks = 0; ks < K / 16; ++ks) {
                const uint32_t o = (ks >> 2) * (kBlk >> 4) + 2 * (ks & 3);
                mma3_ts(tmem, tmem + 128 + ks * 8, tmem + 192 + ks * 8, bh + o, bl + o, idesc, ks > 0);
            }
            (void)t0;
        } else if (mode >= 4 && mode <= 7) {
            // throughput probes with a LEAN issue loop (descriptors precomputed, no branches inside):
            // 4 = A in TMEM (K-major B from smem), 5 = A,B MN-major, 6 = A MN-major / B K-major, 7 = A K-major / B MN-major
            const uint32_t idesc = mode == 4 ? idesc_f16(Split<FP16>::kFmt, 128, N)
                                             : idesc_f16(Split<FP16>::kFmt, 128, N, mode != 7 ? 1 : 0, mode != 6 ? 1 : 0);
            uint64_t da[4], db[4];
            uint32_t ta[4];
            for (int ks = 0; ks < 4; ++ks) {
                ta[ks] = tmem + 128 + ks * 8;
                da[ks] = (mode == 7) ? desc_k_sw128(sbase) + 2 * ks : desc_mn_sw128(sbase, kBlk) + ks * 128;
                db[ks] = (mode == 4 || mode == 6) ? desc_k_sw128(sbase + 4 * kBlk) + 2 * ks : desc_mn_sw128(sbase + 4 * kBlk, kBlk) + ks * 128;
            }
            const long long t0 = clock64();
            if (mode == 4) {
                for (int rep = 0; rep < 60; ++rep) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_f16_ts(tmem, ta[ks], db[ks], idesc, 1u);
                }
            } else {
                for (int rep = 0; rep < 60; ++rep) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_f16_ss(tmem, da[ks], db[ks], idesc, 1u);
                }
            }
            const long long t1 = clock64();
            tc_commit(bar);
            mbar_wait(bar, 0);
            const long long t2 = clock64();
            reinterpret_cast<long long*>(D)[0] = t1 - t0;
            reinterpret_cast<long long*>(D)[1] = t2 - t0;
        } else if (mode == 2) {
            // throughput probe: 240 back-to-back MMAs (K = 64 reused), N in {64,128}; cycles -> D[0..1] as raw ints
            const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, N);
            const uint64_t ah = desc_k_sw128(sbase), bh = desc_k_sw128(sbase + 4 * kBlk);
            const long long t0 = clock64();
            for (int rep = 0; rep < 60; ++rep)
                for (int ks = 0; ks < 4; ++ks) mma_f16_ss(tmem, ah + 2 * ks, bh + 2 * ks, idesc, 1u);
            const long long t1 = clock64();
            tc_commit(bar);
            mbar_wait(bar, 0);
            const long long t2 = clock64();
            reinterpret_cast<long long*>(D)[0] = t1 - t0;
            reinterpret_cast<long long*>(D)[1] = t2 - t0;
        } else {
            const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, N, 1, 1);
            const uint64_t ah = desc_mn_sw128(sbase, kBlk), al = desc_mn_sw128(sbase + 2 * kBlk, kBlk);
            const uint64_t bh = desc_mn_sw128(sbase + 4 * kBlk, kBlk), bl = desc_mn_sw128(sbase + 6 * kBlk, kBlk);
            for (int ks = 0; ks < 8; ++ks) {                          // K = 128 rows = 8 x 16; +2048 bytes per step
                const uint32_t o = ks * (2048 >> 4);
                mma3(tmem, ah + o, al + o, bh + o, bl + o, idesc, ks > 0);
            }
        }
        if (mode != 2 && mode < 4) tc_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < N && mode != 2 && mode < 4; c0 += 32) {
        float v[32];
        tmem_ld32(tmem + lane_base + c0, v);
        for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = v[j];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, 256);
    }
}

}  // namespace tc
}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_debug_set_trace(long long* device_buffer) {
    GEOSSL_CUDA(cudaMemcpyToSymbol(tc::g_trace, &device_buffer, sizeof(device_buffer)));
    return 0;
}

int geossl_tc_selftest(int mode, int fp16, const float* a, const float* b, int K, int N, float* d, void* stream) {
    GEOSSL_REQUIRE(a && b && d, "null pointer");
    GEOSSL_REQUIRE((mode == 0 && (K == 64 || K == 128) && N == 128) || (mode == 1 && K == 128 && (N == 64 || N == 128)) ||
                   ((mode == 2 || (mode >= 4 && mode <= 7)) && K == 64 && (N == 64 || N == 128)) || (mode == 3 && (K == 64 || K == 128) && (N == 64 || N == 128)),
                   "unsupported shape");
    const size_t smem = 8 * tc::kBlk + 64 + 1024;
    if (fp16) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::tc_selftest_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::tc_selftest_kernel<true><<<1, 128, smem, as_stream(stream)>>>(mode, a, b, K, N, d);
    } else {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::tc_selftest_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc::tc_selftest_kernel<false><<<1, 128, smem, as_stream(stream)>>>(mode, a, b, K, N, d);
    }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_filter_fwd_tc(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                         const float* offset, float coeff, float cutoff, int G, int F,
                         const float* w1, const float* b1, const float* w2, const float* b2,
                         float* filt, int bf16_parts, void* stream) {
    if (capacity == 0) return 0;
    GEOSSL_REQUIRE(edge_dist && n_edges_dev && offset && w1 && b1 && w2 && b2 && filt && capacity > 0, "null pointer");
    GEOSSL_REQUIRE(F == 128, "the tensor-core filter kernel is built for num_filters = 128");
    GEOSSL_REQUIRE(G >= 1 && G <= 64, "num_gaussians must be in [1,64]");
    const size_t smem = tc::FwdLayout::kBytes + 1024;
    int64_t tiles = (capacity + tc::kTile - 1) / tc::kTile;
    const int grid = (int)(tiles < kNumSM ? tiles : kNumSM);
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::filter_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::filter_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    if (bf16_parts) {
        GEOSSL_CUDA(launch_pdl(tc::filter_fwd_tc_kernel<false>, dim3(grid), dim3(tc::kThreads), smem, as_stream(stream), edge_dist,
                               n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, b2, filt));
    } else {
        GEOSSL_CUDA(launch_pdl(tc::filter_fwd_tc_kernel<true>, dim3(grid), dim3(tc::kThreads), smem, as_stream(stream), edge_dist,
                               n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, b2, filt));
    }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
