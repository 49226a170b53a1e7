// SchNet filter network backward on the sm_100a tensor cores (tcgen05 + TMEM).
//
// Per 64-edge tile (everything transposed so that TMEM lanes are FEATURES and columns are EDGES):
//   MMA1   D1^T[f][e] = W1[f][:] . rbf[e][:]                      a = D1^T + b1
//   E1     s = ssp(a) -> S tile,  sigma = sigmoid(a) kept in registers
//   P      dU[e][o] = x[src_e][o] * g[tgt_e][o] * cutoff(d_e)     (never materialised in HBM)
//   MMA3   D3^T[i][e] = W2^T[i][:] . dU[e][:]                      ds
//   E3     da = ds * sigma -> dA tile
//   WG2    DW2[o][i] += sum_e dU[e][o] s[e][i]                     (MN-major views of the same tiles)
//   WG1    DW1[f][g] += sum_e da[e][f] rbf[e][g]                   column g = 63 of the rbf tile is a
//                                                                  constant 1, so DW1[:,63] = db1
// DW2 / DW1 live in TMEM for the whole kernel; db2 is summed by the dU producers; per-CTA partials are
// reduced in a fixed order by a second kernel (deterministic).  Operands are split into two bf16 parts
// (fp32 range for the gradient operands) and every product is three MMAs with fp32 accumulation.
//
// Replaces the autograd of schnet.py:141-145,186-187,190,194-195 w.r.t. the filter-network parameters.
// Warp roles as in filter_tc.cu: warps 0-15 epilogue (lane quadrant x edge quarter), 16-19 producers, 20 MMA.
#include "common.cuh"
#include "tc.cuh"

namespace geossl {
namespace tc {

__device__ long long* g_trace_bwd = nullptr;   // optional clock64() trace of CTA 0 (geossl_debug_set_trace_bwd)
__device__ __forceinline__ void trace_b(int tile, int event) {
    long long* t = g_trace_bwd;
    if (t != nullptr && blockIdx.x == 0 && tile < 32 && (threadIdx.x & 31) == 0) t[tile * 16 + event] = clock64();
}

constexpr int kBT = 64;                  // edges per tile
constexpr int kBlkW = 128 * 128;         // bytes of a [128 rows x 64 k] weight block
constexpr int kBlkT = kBT * 128;         // bytes of a [64 rows x 64 k] tile block
constexpr int kBwdEpiWarps = 16, kBwdEpiThreads = kBwdEpiWarps * 32, kBwdThreads = kBwdEpiThreads + 128 + 32;
constexpr int kBwdProdWarp0 = kBwdEpiWarps, kBwdMmaWarp = kBwdEpiWarps + 4;
constexpr bool kBwdFP16 = false;         // bf16 parts

struct BwdLayout {
    static constexpr int W1_hi = 0, W1_lo = W1_hi + kBlkW;                    // A of MMA1: rows f, k = g
    static constexpr int W2T_hi = W1_lo + kBlkW, W2T_lo = W2T_hi + 2 * kBlkW;  // A of MMA3: rows i, k = o
    static constexpr int PHI = W2T_lo + 2 * kBlkW;                             // 2 x (hi, lo)          32 KB
    static constexpr int SA = PHI + 4 * kBlkT;                                 // S / dA (hi 2 blk, lo 2 blk) 32 KB
    static constexpr int DU = SA + 4 * kBlkT;                                  // 2 x (hi 2 blk, lo 2 blk)   64 KB
    static constexpr int B1 = DU + 8 * kBlkT;
    static constexpr int OFF = B1 + 512;
    static constexpr int BAR = OFF + 256;
    static constexpr int TMEM_PTR = BAR + 24 * 8;
    static constexpr int kBytes = TMEM_PTR + 16;
};
static_assert(BwdLayout::kBytes + 1024 <= 227 * 1024, "shared memory budget");

enum BBar { PHI_FULL_ = 0, PHI_FREE_ = 2, DU_FULL_ = 4, DU_FREE_ = 6, D1_FULL_ = 8, D1_FREE_ = 10, D3_FULL_ = 12, D3_FREE_ = 14,
            S_FULL_ = 16, S_FREE_ = 17, DA_FULL_ = 18, DA_FREE_ = 19, DONE_ = 20 };

// per-CTA partial layout (identical to Partial<128> of filter_simt.cu so the same reduction applies)
struct Part {
    static constexpr int kW2 = 0, kW1 = 128 * 128, kB2 = kW1 + 64 * 128, kB1 = kB2 + 128, kFloats = kB1 + 128;
};

__device__ __forceinline__ void store_split_bf16(uint8_t* hi_block, uint8_t* lo_block, uint32_t row, uint32_t k, float x) {
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    const uint32_t off = sw128_offset(row, k);
    *reinterpret_cast<__nv_bfloat16*>(hi_block + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(lo_block + off) = l;
}

__global__ void __launch_bounds__(kBwdThreads, 1)
filter_bwd_tc_kernel(const float* __restrict__ edge_dist, const int32_t* __restrict__ n_edges_dev, int64_t capacity,
                     const float* __restrict__ offset, float coeff, float cutoff, int G,
                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                     const float* __restrict__ x, const float* __restrict__ grad_out,
                     const int32_t* __restrict__ src, const int32_t* __restrict__ edge_tgt, float* __restrict__ workspace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    using L = BwdLayout;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + L::BAR;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    float* sB1 = reinterpret_cast<float*>(smem + L::B1);
    float* sOff = reinterpret_cast<float*>(smem + L::OFF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int64_t n_edges = (int64_t)(*n_edges_dev);
    if (n_edges > capacity) n_edges = capacity;
    const int64_t n_tiles = (n_edges + kBT - 1) / kBT;
    const int my_tiles = (blockIdx.x < n_tiles) ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    float* ws = workspace + (int64_t)blockIdx.x * Part::kFloats;

    // ---- one-time setup
    for (int idx = tid; idx < 128 * 8; idx += kBwdThreads) {            // W1[f][g], zero padded to 64
        const int f = idx >> 3, c = idx & 7;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int g = c * 8 + j; v[j] = (g < G) ? __ldg(w1 + f * G + g) : 0.f; }
        store_chunk8<kBwdFP16>(smem + L::W1_hi, smem + L::W1_lo, f, c * 8, v);
    }
    for (int idx = tid; idx < 128 * 16; idx += kBwdThreads) {           // W2^T[i][o] = W2[o][i]
        const int i = idx & 127, c = idx >> 7;                           // lanes over i: coalesced reads of W2 rows
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(w2 + (c * 8 + j) * 128 + i);
        const int blk = c >> 3;
        store_chunk8<kBwdFP16>(smem + L::W2T_hi + blk * kBlkW, smem + L::W2T_lo + blk * kBlkW, i, (c & 7) * 8, v);
    }
    if (tid < 128) sB1[tid] = __ldg(b1 + tid);
    if (tid < 64) sOff[tid] = (tid < G) ? __ldg(offset + tid) : 0.f;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(PHI_FULL_ + b), 4); mbar_init(bar(PHI_FREE_ + b), 1);
            mbar_init(bar(DU_FULL_ + b), 4);  mbar_init(bar(DU_FREE_ + b), 1);
            mbar_init(bar(D1_FULL_ + b), 1);  mbar_init(bar(D1_FREE_ + b), kBwdEpiWarps);
            mbar_init(bar(D3_FULL_ + b), 1);  mbar_init(bar(D3_FREE_ + b), kBwdEpiWarps);
        }
        mbar_init(bar(S_FULL_), kBwdEpiWarps);  mbar_init(bar(S_FREE_), 1);
        mbar_init(bar(DA_FULL_), kBwdEpiWarps); mbar_init(bar(DA_FREE_), 1);
        mbar_init(bar(DONE_), 1);
        fence_barrier_init();
    }
    if (warp == kBwdMmaWarp) tmem_alloc(sbase + L::TMEM_PTR, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    const uint32_t tDW2 = tmem, tDW1 = tmem + 128;
    const uint32_t tD1[2] = {tmem + 192, tmem + 256}, tD3[2] = {tmem + 320, tmem + 384};

    if (warp >= kBwdProdWarp0 && warp < kBwdMmaWarp) {
        // ===================== producers: rbf tile and dU tile
        const int tp = tid - kBwdEpiThreads;
        const int cg = tp & 15, ro = tp >> 4;                            // dU: 8 columns 8cg.., rows ro + 8 s
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t e_base = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kBT;
            const int b = i & 1;
            {   // ---- rbf(i): thread = (row, half of the 8 chunks)
                mbar_wait(bar(PHI_FREE_ + b), ((i >> 1) & 1) ^ 1);
                if (warp == kBwdProdWarp0) trace_b(i, 0);
                const int row = tp >> 1, c0 = (tp & 1) * 4;
                const int64_t e = e_base + row;
                const bool valid = e < n_edges;
                const float d = valid ? __ldg(edge_dist + e) : 0.f;
                uint8_t* hi = smem + L::PHI + b * 2 * kBlkT;
                uint8_t* lo = hi + kBlkT;
#pragma unroll
                for (int c = c0; c < c0 + 4; ++c) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int g = c * 8 + j;
                        const float diff = d - sOff[g];
                        v[j] = (g < G) ? __expf(__fmul_rn(coeff, __fmul_rn(diff, diff))) : ((g == 63 && valid) ? 1.f : 0.f);
                    }
                    store_chunk8<kBwdFP16>(hi, lo, row, c * 8, v);
                }
                fence_proxy_async();
                warp_arrive(bar(PHI_FULL_ + b));
                if (warp == kBwdProdWarp0) trace_b(i, 1);
            }
            {   // ---- dU(i)
                mbar_wait(bar(DU_FREE_ + b), ((i >> 1) & 1) ^ 1);
                if (warp == kBwdProdWarp0) trace_b(i, 2);
                uint8_t* hi = smem + L::DU + b * 4 * kBlkT + (cg >> 3) * kBlkT;
                uint8_t* lo = hi + 2 * kBlkT;
#pragma unroll
                for (int sb = 0; sb < 2; ++sb) {
                    float4 xv[4][2], gv[4][2];
                    float cs[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t e = e_base + (sb * 4 + u) * 8 + ro;
                        if (e < n_edges) {
                            const float* xr = x + (int64_t)__ldg(src + e) * 128 + cg * 8;
                            const float* gr = grad_out + (int64_t)__ldg(edge_tgt + e) * 128 + cg * 8;
                            xv[u][0] = ldg4(xr); xv[u][1] = ldg4(xr + 4);
                            gv[u][0] = ldg4(gr); gv[u][1] = ldg4(gr + 4);
                            cs[u] = cosine_cutoff(__ldg(edge_dist + e), cutoff);
                        } else {
                            xv[u][0] = xv[u][1] = gv[u][0] = gv[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                            cs[u] = 0.f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int row = (sb * 4 + u) * 8 + ro;
                        float v[8] = {xv[u][0].x * gv[u][0].x * cs[u], xv[u][0].y * gv[u][0].y * cs[u],
                                      xv[u][0].z * gv[u][0].z * cs[u], xv[u][0].w * gv[u][0].w * cs[u],
                                      xv[u][1].x * gv[u][1].x * cs[u], xv[u][1].y * gv[u][1].y * cs[u],
                                      xv[u][1].z * gv[u][1].z * cs[u], xv[u][1].w * gv[u][1].w * cs[u]};
#pragma unroll
                        for (int k = 0; k < 8; ++k) acc[k] += v[k];
                        store_chunk8<kBwdFP16>(hi, lo, row, (cg & 7) * 8, v);
                    }
                }
                fence_proxy_async();
                warp_arrive(bar(DU_FULL_ + b));
                if (warp == kBwdProdWarp0) trace_b(i, 3);
            }
        }
        // db2[o] = sum over the 8 row groups; scratch = dU buffer 0 once every MMA has retired
        mbar_wait(bar(DONE_), 0);
        float* red = reinterpret_cast<float*>(smem + L::DU);
#pragma unroll
        for (int k = 0; k < 8; ++k) red[ro * 128 + cg * 8 + k] = acc[k];
        asm volatile("bar.sync 1, 128;" ::: "memory");
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 8; ++r) s += red[r * 128 + tp];
        ws[Part::kB2 + tp] = s;
    } else if (warp == kBwdMmaWarp) {
        // ===================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t fmt = Split<kBwdFP16>::kFmt;
            const uint32_t id_t = idesc_f16(fmt, 128, kBT);                 // D^T tiles: M = features, N = 64 edges
            const uint32_t id_w2 = idesc_f16(fmt, 128, 128, 1, 1), id_w1 = idesc_f16(fmt, 128, 64, 1, 1);
            const uint64_t dW1h = desc_k_sw128(sbase + L::W1_hi), dW1l = desc_k_sw128(sbase + L::W1_lo);
            const uint64_t dW2h = desc_k_sw128(sbase + L::W2T_hi), dW2l = desc_k_sw128(sbase + L::W2T_lo);
            const uint64_t dSAh = desc_mn_sw128(sbase + L::SA, kBlkT), dSAl = desc_mn_sw128(sbase + L::SA + 2 * kBlkT, kBlkT);
            auto mma1 = [&](int i) {
                const int b = i & 1;
                mbar_wait(bar(PHI_FULL_ + b), (i >> 1) & 1);
                mbar_wait(bar(D1_FREE_ + b), ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint64_t ph = desc_k_sw128(sbase + L::PHI + b * 2 * kBlkT), pl = desc_k_sw128(sbase + L::PHI + b * 2 * kBlkT + kBlkT);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) mma3(tD1[b], dW1h + 2 * kk, dW1l + 2 * kk, ph + 2 * kk, pl + 2 * kk, id_t, kk > 0);
                tc_commit(bar(D1_FULL_ + b));
            };
            if (my_tiles > 0) mma1(0);
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                const uint32_t du = sbase + L::DU + b * 4 * kBlkT;
                // ---- MMA3(i): ds^T = W2^T . dU^T
                mbar_wait(bar(DU_FULL_ + b), (i >> 1) & 1);
                mbar_wait(bar(D3_FREE_ + b), ((i >> 1) & 1) ^ 1);
                tc_fence_after();
                trace_b(i, 4);
                {
                    const uint64_t uh = desc_k_sw128(du), ul = desc_k_sw128(du + 2 * kBlkT);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const uint32_t ow = kb * (kBlkW >> 4) + 2 * kk, ot = kb * (kBlkT >> 4) + 2 * kk;
                            mma3(tD3[b], dW2h + ow, dW2l + ow, uh + ot, ul + ot, id_t, (kb | kk) > 0);
                        }
                }
                tc_commit(bar(D3_FULL_ + b));
                // ---- WG2(i): DW2 += dU^T s
                mbar_wait(bar(S_FULL_), i & 1);
                tc_fence_after();
                trace_b(i, 5);
                {
                    const uint64_t uh = desc_mn_sw128(du, kBlkT), ul = desc_mn_sw128(du + 2 * kBlkT, kBlkT);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t o = ks * (2048 >> 4);
                        mma3(tDW2, uh + o, ul + o, dSAh + o, dSAl + o, id_w2, (i | ks) > 0);
                    }
                }
                tc_commit(bar(S_FREE_));
                tc_commit(bar(DU_FREE_ + b));
                // ---- MMA1(i+1)
                trace_b(i, 6);
                if (i + 1 < my_tiles) mma1(i + 1);
                // ---- WG1(i): DW1 += dA^T rbf
                mbar_wait(bar(DA_FULL_), i & 1);
                tc_fence_after();
                trace_b(i, 7);
                {
                    const uint32_t ph_addr = sbase + L::PHI + b * 2 * kBlkT;
                    const uint64_t ph = desc_mn_sw128(ph_addr, kBlkT), pl = desc_mn_sw128(ph_addr + kBlkT, kBlkT);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t o = ks * (2048 >> 4);
                        mma3(tDW1, dSAh + o, dSAl + o, ph + o, pl + o, id_w1, (i | ks) > 0);
                    }
                }
                tc_commit(bar(DA_FREE_));
                tc_commit(bar(PHI_FREE_ + b));
                trace_b(i, 8);
            }
            tc_commit(bar(DONE_));
        }
    } else {
        // ===================== epilogue warps (TMEM lane = feature, columns = edges): warp = (quadrant q, edge quarter eq)
        const int q = warp & 3, eq = warp >> 2;
        const int f = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float b1f = sB1[f];
        uint8_t* sa_hi = smem + L::SA + (f >> 6) * kBlkT;
        uint8_t* sa_lo = sa_hi + 2 * kBlkT;
        const uint32_t kcol = f & 63;
        float sig[16];
        auto e1 = [&](int i) {
            const int b = i & 1;
            mbar_wait(bar(D1_FULL_ + b), (i >> 1) & 1);
            tc_fence_after();
            if (warp == 0) trace_b(i, 9);
            float a[16];
            tmem_ld16(tD1[b] + lane_base + eq * 16, a);
            tc_fence_before();
            warp_arrive(bar(D1_FREE_ + b));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float av = a[j] + b1f;
                a[j] = ssp_fast(av);
                sig[j] = av > 20.f ? 1.f : sigmoid_fast(av);
            }
            if (i > 0) mbar_wait(bar(DA_FREE_), (i - 1) & 1);            // WG1(i-1) has consumed dA in the shared S/dA buffer
#pragma unroll
            for (int j = 0; j < 16; ++j) store_split_bf16(sa_hi, sa_lo, eq * 16 + j, kcol, a[j]);
            fence_proxy_async();
            warp_arrive(bar(S_FULL_));
            if (warp == 0) trace_b(i, 10);
        };
        if (my_tiles > 0) e1(0);
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1;
            mbar_wait(bar(D3_FULL_ + b), (i >> 1) & 1);
            tc_fence_after();
            if (warp == 0) trace_b(i, 11);
            float ds[16];
            tmem_ld16(tD3[b] + lane_base + eq * 16, ds);
            tc_fence_before();
            warp_arrive(bar(D3_FREE_ + b));
            mbar_wait(bar(S_FREE_), i & 1);                              // WG2(i) has consumed S
            if (warp == 0) trace_b(i, 12);
#pragma unroll
            for (int j = 0; j < 16; ++j) store_split_bf16(sa_hi, sa_lo, eq * 16 + j, kcol, ds[j] * sig[j]);
            fence_proxy_async();
            warp_arrive(bar(DA_FULL_));
            if (warp == 0) trace_b(i, 13);
            if (i + 1 < my_tiles) e1(i + 1);
        }
        // ---- accumulators -> per-CTA partials
        mbar_wait(bar(DONE_), 0);
        tc_fence_after();
        if (my_tiles > 0) {
            {
                float v[32];
                tmem_ld32(tDW2 + lane_base + eq * 32, v);
                float* dst = ws + Part::kW2 + f * 128 + eq * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            float v[16];
            tmem_ld16(tDW1 + lane_base + eq * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int g = eq * 16 + j;
                if (g < G) ws[Part::kW1 + g * 128 + f] = v[j];
                if (g == 63) ws[Part::kB1 + f] = v[j];
            }
        } else {
            for (int c = eq * 32; c < eq * 32 + 32; ++c) ws[Part::kW2 + f * 128 + c] = 0.f;
            for (int g = eq * 16; g < eq * 16 + 16; ++g) ws[Part::kW1 + g * 128 + f] = 0.f;
            if (eq == 3) ws[Part::kB1 + f] = 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kBwdMmaWarp) {
        __syncwarp();
        tmem_dealloc(tmem, 512);
    }
}

__global__ void filter_bwd_tc_reduce_kernel(const float* __restrict__ workspace, int n_parts, int G,
                                            float* __restrict__ gw1, float* __restrict__ gb1,
                                            float* __restrict__ gw2, float* __restrict__ gb2) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Part::kFloats) return;
    const bool w1_slot = idx >= Part::kW1 && idx < Part::kB2;
    const int g = w1_slot ? (idx - Part::kW1) / 128 : 0;
    if (w1_slot && g >= G) return;                                      // padded gaussians are never written
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) s += workspace[(int64_t)p * Part::kFloats + idx];
    if (idx < Part::kW1) gw2[idx] = s;
    else if (w1_slot) gw1[((idx - Part::kW1) % 128) * G + g] = s;
    else if (idx < Part::kB1) gb2[idx - Part::kB2] = s;
    else gb1[idx - Part::kB1] = s;
}

}  // namespace tc
}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_debug_set_trace_bwd(long long* device_buffer) {
    GEOSSL_CUDA(cudaMemcpyToSymbol(tc::g_trace_bwd, &device_buffer, sizeof(device_buffer)));
    return 0;
}

int64_t geossl_filter_bwd_tc_workspace(void) { return (int64_t)kNumSM * tc::Part::kFloats; }

int geossl_filter_bwd_tc(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                         const float* offset, float coeff, float cutoff, int G, int F,
                         const float* w1, const float* b1, const float* w2,
                         const float* x, const float* grad_out, const int32_t* src, const int32_t* edge_tgt,
                         float* workspace, float* gw1, float* gb1, float* gw2, float* gb2, void* stream) {
    GEOSSL_REQUIRE(edge_dist && n_edges_dev && offset && w1 && b1 && w2 && x && grad_out && src && edge_tgt && workspace &&
                   gw1 && gb1 && gw2 && gb2, "null pointer");
    GEOSSL_REQUIRE(F == 128, "the tensor-core filter kernel is built for num_filters = 128");
    GEOSSL_REQUIRE(G >= 1 && G <= 63, "num_gaussians must be in [1,63] (column 63 of the rbf tile carries the bias sum)");
    const size_t smem = tc::BwdLayout::kBytes + 1024;
    static bool configured = false;
    if (!configured) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::filter_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    tc::filter_bwd_tc_kernel<<<kNumSM, tc::kBwdThreads, smem, as_stream(stream)>>>(edge_dist, n_edges_dev, capacity, offset, coeff,
                                                                                  cutoff, G, w1, b1, w2, x, grad_out, src, edge_tgt,
                                                                                  workspace);
    GEOSSL_LAUNCH_CHECK();
    tc::filter_bwd_tc_reduce_kernel<<<(tc::Part::kFloats + 255) / 256, 256, 0, as_stream(stream)>>>(workspace, kNumSM, G, gw1, gb1,
                                                                                                   gw2, gb2);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
