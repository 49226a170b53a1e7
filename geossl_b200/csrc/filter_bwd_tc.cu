// SchNet filter network backward on the sm_100a tensor cores (tcgen05 + TMEM).
//
// Per 64-edge tile (everything transposed so that TMEM lanes are FEATURES and columns are EDGES):
//   MMA1   D1^T[f][e] = W1[f][:] . rbf[e][:]                      a = D1^T + b1        (A = W1 resident in TMEM)
//   E1     s = ssp(a) -> S tile,  sigma = sigmoid(a) kept in registers
//   P-dU   dU[e][o] = x[src_e][o] * g[tgt_e][o] * cutoff(d_e)     (never materialised in HBM)
//   MMA3   D3^T[i][e] = W2^T[i][:] . dU[e][:]                      ds                  (A = W2^T resident in TMEM)
//   E3     da = ds * sigma -> dA tile
//   WG2    DW2[o][i] += sum_e dU[e][o] s[e][i]                     (MN-major views of the same smem tiles)
//   WG1    DW1[f][g] += sum_e da[e][f] rbf[e][g]                   column g = 63 of the rbf tile is a
//                                                                  constant 1, so DW1[:,63] = db1
// Tensor memory (512 columns): DW2 128 | DW1 64 | W1 hi,lo 32+32 | W2^T hi,lo 64+64 | D1^T 64 | D3^T 64.
// Keeping the two weight matrices in TMEM halves the shared-memory operand traffic of MMA1/MMA3 (SS-mode MMAs
// are shared-memory-bandwidth bound: profiles/r01_mma_probe.txt) and frees 96 KB for double-buffered tiles.
// DW2 / DW1 accumulate in TMEM for the whole kernel; db2 is summed by the dU producers; per-CTA partials are
// reduced in a fixed order by a second kernel (deterministic).  Operands are split into two bf16 parts
// (fp32 range for the gradient operands) and every product is three MMAs with fp32 accumulation.
//
// Replaces the autograd of schnet.py:141-145,186-187,190,194-195 w.r.t. the filter-network parameters.
// Warps: 0-15 epilogue (lane quadrant x edge quarter), 16-19 rbf producers, 20-27 dU producers, 28 MMA issuer.
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
constexpr int kBlkT = kBT * 128;         // bytes of a [64 rows x 64 k] tile block
constexpr int kBwdEpiWarps = 16, kBwdPhiWarps = 4, kBwdDuWarps = 8;
constexpr int kBwdPhiWarp0 = kBwdEpiWarps, kBwdDuWarp0 = kBwdPhiWarp0 + kBwdPhiWarps, kBwdMmaWarp = kBwdDuWarp0 + kBwdDuWarps;
constexpr int kBwdThreads = (kBwdMmaWarp + 1) * 32;
constexpr bool kBwdFP16 = false;         // bf16 parts

struct BwdLayout {
    static constexpr int PHI = 0;                          // 2 x (hi, lo)                     32 KB
    static constexpr int S = PHI + 4 * kBlkT;              // 2 x (hi 2 blk, lo 2 blk)         64 KB
    static constexpr int DA = S + 8 * kBlkT;               // (hi 2 blk, lo 2 blk)             32 KB
    static constexpr int DU = DA + 4 * kBlkT;              // 2 x (hi 2 blk, lo 2 blk)         64 KB
    static constexpr int OFF = DU + 8 * kBlkT;             // 64 floats
    static constexpr int BAR = OFF + 256;
    static constexpr int TMEM_PTR = BAR + 24 * 8;
    static constexpr int kBytes = TMEM_PTR + 16;
};
static_assert(BwdLayout::kBytes + 1024 <= 227 * 1024, "shared memory budget");

enum BBar { PHI_FULL_ = 0, PHI_FREE_ = 2, DU_FULL_ = 4, DU_FREE_ = 6, S_FULL_ = 8, S_FREE_ = 10, D1_FULL_ = 12, D1_FREE_ = 13,
            D3_FULL_ = 14, D3_FREE_ = 15, DA_FULL_ = 16, DA_FREE_ = 17, DONE_ = 18 };

// per-CTA partial layout (identical to Partial<128> of filter_simt.cu)
struct Part {
    static constexpr int kW2 = 0, kW1 = 128 * 128, kB2 = kW1 + 64 * 128, kB1 = kB2 + 128, kFloats = kB1 + 128;
};

// PAIRS: the tile rows are undirected pairs (geossl_pair_index): edge_dist / n_edges_dev describe the pairs, and
// dU_u = C(d_u) * (x[s] * g[t] + x[t] * g[s]) sums both directions of pair u = (s -> t) (second term only if the reverse
// edge exists: pair_atoms[u] = (s, t) or (s, ~t)).  Everything downstream of dU is linear in it, so the parameter gradients are unchanged.
template <bool PAIRS>
__global__ void __launch_bounds__(kBwdThreads, 1)
filter_bwd_tc_kernel(const float* __restrict__ edge_dist, const int32_t* __restrict__ n_edges_dev, int64_t capacity,
                     const float* __restrict__ offset, float coeff, float cutoff, int G,
                     const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                     const float* __restrict__ x, const float* __restrict__ grad_out,
                     const int32_t* __restrict__ src, const int32_t* __restrict__ edge_tgt,
                     const int2* __restrict__ pair_atoms, float* __restrict__ workspace) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    using L = BwdLayout;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + L::BAR;
    auto bar = [&](int i) { return bar0 + 8u * i; };
    float* sOff = reinterpret_cast<float*>(smem + L::OFF);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    pdl_wait();

    int64_t n_edges = (int64_t)(*n_edges_dev);
    if (n_edges > capacity) n_edges = capacity;
    const int64_t n_tiles = (n_edges + kBT - 1) / kBT;
    const int my_tiles = (blockIdx.x < n_tiles) ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;
    float* ws = workspace + (int64_t)blockIdx.x * Part::kFloats;

    // ---- one-time setup: barriers, TMEM, weights into TMEM
    if (tid < 64) sOff[tid] = (tid < G) ? __ldg(offset + tid) : 1e18f;    // padded gaussians evaluate to exactly 0
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar(PHI_FULL_ + b), kBwdPhiWarps); mbar_init(bar(PHI_FREE_ + b), 1);
            mbar_init(bar(DU_FULL_ + b), kBwdDuWarps);   mbar_init(bar(DU_FREE_ + b), 1);
            mbar_init(bar(S_FULL_ + b), kBwdEpiWarps);   mbar_init(bar(S_FREE_ + b), 1);
        }
        mbar_init(bar(D1_FULL_), 1);  mbar_init(bar(D1_FREE_), kBwdEpiWarps);
        mbar_init(bar(D3_FULL_), 1);  mbar_init(bar(D3_FREE_), kBwdEpiWarps);
        mbar_init(bar(DA_FULL_), kBwdEpiWarps); mbar_init(bar(DA_FREE_), 1);
        mbar_init(bar(DONE_), 1);
        fence_barrier_init();
    }
    if (warp == kBwdMmaWarp) tmem_alloc(sbase + L::TMEM_PTR, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    const uint32_t tDW2 = tmem, tDW1 = tmem + 128, tW1h = tmem + 192, tW1l = tmem + 224, tW2h = tmem + 256, tW2l = tmem + 320;
    const uint32_t tD1 = tmem + 384, tD3 = tmem + 448;
    if (warp < kBwdEpiWarps) {
        // weights -> tensor memory: lane f owns row f of W1 (k = g, zero padded) and of W2^T (k = o)
        const int q = warp & 3, part = warp >> 2, f = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) { const int g = part * 16 + j; v[j] = (g < G) ? __ldg(w1 + f * G + g) : 0.f; }
            tmem_store_split16<kBwdFP16>(tW1h + lane_base + part * 8, tW1l + lane_base + part * 8, v);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int o0 = (part * 2 + h) * 16;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __ldg(w2 + (o0 + j) * 128 + f);          // W2^T[f][o] = W2[o][f]
            tmem_store_split16<kBwdFP16>(tW2h + lane_base + o0 / 2, tW2l + lane_base + o0 / 2, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    if (warp >= kBwdPhiWarp0 && warp < kBwdDuWarp0) {
        // ===================== rbf producers: thread = (row, half of the 8 chunks)
        const int tp = tid - kBwdPhiWarp0 * 32;
        const int row = tp >> 1, c0 = (tp & 1) * 4;
        const float cl2 = coeff * 1.4426950408889634f;
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t e = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kBT + row;
            const int b = i & 1;
            const bool valid = e < n_edges;
            const float d = valid ? __ldg(edge_dist + e) : 0.f;
            mbar_wait(bar(PHI_FREE_ + b), ((i >> 1) & 1) ^ 1);
            if (warp == kBwdPhiWarp0) trace_b(i, 0);
            uint8_t* hi = smem + L::PHI + b * 2 * kBlkT;
            uint8_t* lo = hi + kBlkT;
#pragma unroll
            for (int c = c0; c < c0 + 4; ++c) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float diff = d - sOff[c * 8 + j];
                    v[j] = ex2_approx(cl2 * (diff * diff));
                }
                if (c == 7) v[7] = valid ? 1.f : 0.f;                    // ones column: DW1[:,63] accumulates db1
                store_chunk8<kBwdFP16>(hi, lo, row, c * 8, v);
            }
            fence_proxy_async();
            warp_arrive(bar(PHI_FULL_ + b));
            if (warp == kBwdPhiWarp0) trace_b(i, 1);
        }
    } else if (warp >= kBwdDuWarp0 && warp < kBwdMmaWarp) {
        // ===================== dU producers: thread = (8 columns 8cg.., rows ro + 16 s)
        const int tq = tid - kBwdDuWarp0 * 32;
        const int cg = tq & 15, ro = tq >> 4;
        const float pi_over_rc = kPi / cutoff;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int i = 0; i < my_tiles; ++i) {
            const int64_t e_base = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * kBT;
            const int b = i & 1;
            uint8_t* hi = smem + L::DU + b * 4 * kBlkT + (cg >> 3) * kBlkT;
            uint8_t* lo = hi + 2 * kBlkT;
            if constexpr (PAIRS) {
                // the four rows' atom ids and lengths first (one latency), so that the row gathers below start back to back
                int2 st[4];
                float dd[4];
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int64_t u = e_base + rr * 16 + ro;
                    st[rr] = make_int2(-1, 0);
                    dd[rr] = 0.f;
                    if (u < n_edges) { st[rr] = __ldg(pair_atoms + u); dd[rr] = __ldg(edge_dist + u); }
                }
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const int row = rr * 16 + ro;
                    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 xs0 = z4, xs1 = z4, gt0 = z4, gt1 = z4, xt0 = z4, xt1 = z4, gs0 = z4, gs1 = z4;
                    float cs = 0.f;
                    if (st[rr].x >= 0) {
                        const bool both = st[rr].y >= 0;
                        const int64_t sa = (int64_t)st[rr].x * 128 + cg * 8, ta = (int64_t)(both ? st[rr].y : ~st[rr].y) * 128 + cg * 8;
                        xs0 = ldg4(x + sa); xs1 = ldg4(x + sa + 4);
                        gt0 = ldg4(grad_out + ta); gt1 = ldg4(grad_out + ta + 4);
                        if (both) {
                            xt0 = ldg4(x + ta); xt1 = ldg4(x + ta + 4);
                            gs0 = ldg4(grad_out + sa); gs1 = ldg4(grad_out + sa + 4);
                        }
                        cs = 0.5f * (__cosf(dd[rr] * pi_over_rc) + 1.0f);
                    }
                    if (rr == 0) {
                        mbar_wait(bar(DU_FREE_ + b), ((i >> 1) & 1) ^ 1);   // gathers of the first row are already in flight
                        if (warp == kBwdDuWarp0) trace_b(i, 2);
                    }
                    float v[8] = {fmaf(xs0.x, gt0.x, xt0.x * gs0.x) * cs, fmaf(xs0.y, gt0.y, xt0.y * gs0.y) * cs,
                                  fmaf(xs0.z, gt0.z, xt0.z * gs0.z) * cs, fmaf(xs0.w, gt0.w, xt0.w * gs0.w) * cs,
                                  fmaf(xs1.x, gt1.x, xt1.x * gs1.x) * cs, fmaf(xs1.y, gt1.y, xt1.y * gs1.y) * cs,
                                  fmaf(xs1.z, gt1.z, xt1.z * gs1.z) * cs, fmaf(xs1.w, gt1.w, xt1.w * gs1.w) * cs};
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += v[k];
                    store_chunk8<kBwdFP16>(hi, lo, row, (cg & 7) * 8, v);
                }
            } else {
#pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
                float4 xv[2][2], gv[2][2];
                float cs[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int64_t e = e_base + (sb * 2 + u) * 16 + ro;
                    if (e < n_edges) {
                        const float* xr = x + (int64_t)__ldg(src + e) * 128 + cg * 8;
                        const float* gr = grad_out + (int64_t)__ldg(edge_tgt + e) * 128 + cg * 8;
                        xv[u][0] = ldg4(xr); xv[u][1] = ldg4(xr + 4);
                        gv[u][0] = ldg4(gr); gv[u][1] = ldg4(gr + 4);
                        cs[u] = 0.5f * (__cosf(__ldg(edge_dist + e) * pi_over_rc) + 1.0f);
                    } else {
                        xv[u][0] = xv[u][1] = gv[u][0] = gv[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        cs[u] = 0.f;
                    }
                }
                if (sb == 0) {
                    mbar_wait(bar(DU_FREE_ + b), ((i >> 1) & 1) ^ 1);   // gathers of the first half are already in flight
                    if (warp == kBwdDuWarp0) trace_b(i, 2);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int row = (sb * 2 + u) * 16 + ro;
                    float v[8] = {xv[u][0].x * gv[u][0].x * cs[u], xv[u][0].y * gv[u][0].y * cs[u],
                                  xv[u][0].z * gv[u][0].z * cs[u], xv[u][0].w * gv[u][0].w * cs[u],
                                  xv[u][1].x * gv[u][1].x * cs[u], xv[u][1].y * gv[u][1].y * cs[u],
                                  xv[u][1].z * gv[u][1].z * cs[u], xv[u][1].w * gv[u][1].w * cs[u]};
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += v[k];
                    store_chunk8<kBwdFP16>(hi, lo, row, (cg & 7) * 8, v);
                }
            }
            }
            fence_proxy_async();
            warp_arrive(bar(DU_FULL_ + b));
            if (warp == kBwdDuWarp0) trace_b(i, 3);
        }
        // db2[o] = sum over the 16 row groups; scratch = dU buffer 0 once every MMA has retired
        mbar_wait(bar(DONE_), 0);
        // (every producer's tile stores are ordered before this point through DU_FULL -> MMA -> DONE; the named barrier makes
        // that ordering explicit among the eight producer warps, which is also what compute-sanitizer's racecheck can see)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        float* red = reinterpret_cast<float*>(smem + L::DU);
#pragma unroll
        for (int k = 0; k < 8; ++k) red[ro * 128 + cg * 8 + k] = acc[k];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tq < 128) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) s += red[r * 128 + tq];
            ws[Part::kB2 + tq] = s;
        }
    } else if (warp == kBwdMmaWarp) {
        // ===================== MMA issuer
        {
            const bool leader = elect_one_sync();
            constexpr uint32_t fmt = Split<kBwdFP16>::kFmt;
            const uint32_t id_t = idesc_f16(fmt, 128, kBT);                 // D^T tiles: M = features, N = 64 edges
            const uint32_t id_w2 = idesc_f16(fmt, 128, 128, 1, 1), id_w1 = idesc_f16(fmt, 128, 64, 1, 1);
            const uint64_t dAh = desc_mn_sw128(sbase + L::DA, kBlkT), dAl = desc_mn_sw128(sbase + L::DA + 2 * kBlkT, kBlkT);
            auto mma1 = [&](int i) {
                const int b = i & 1;
                mbar_wait(bar(PHI_FULL_ + b), (i >> 1) & 1);
                mbar_wait(bar(D1_FREE_), (i & 1) ^ 1);
                tc_fence_after();
                const uint64_t ph = desc_k_sw128(sbase + L::PHI + b * 2 * kBlkT), pl = desc_k_sw128(sbase + L::PHI + b * 2 * kBlkT + kBlkT);
                if (leader) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) mma3_ts(tD1, tW1h + 8 * kk, tW1l + 8 * kk, ph + 2 * kk, pl + 2 * kk, id_t, kk > 0);
                    tc_commit(bar(D1_FULL_));
                }
                __syncwarp();
            };
            if (my_tiles > 0) mma1(0);
            for (int i = 0; i < my_tiles; ++i) {
                const int b = i & 1;
                const uint32_t du = sbase + L::DU + b * 4 * kBlkT;
                const uint32_t sa = sbase + L::S + b * 4 * kBlkT;
                // ---- MMA3(i): ds^T = W2^T . dU^T
                mbar_wait(bar(DU_FULL_ + b), (i >> 1) & 1);
                mbar_wait(bar(D3_FREE_), (i & 1) ^ 1);
                tc_fence_after();
                trace_b(i, 4);
                {
                    const uint64_t uh = desc_k_sw128(du), ul = desc_k_sw128(du + 2 * kBlkT);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t ot = (ks >> 2) * (kBlkT >> 4) + 2 * (ks & 3);
                            mma3_ts(tD3, tW2h + 8 * ks, tW2l + 8 * ks, uh + ot, ul + ot, id_t, ks > 0);
                        }
                        tc_commit(bar(D3_FULL_));
                    }
                    __syncwarp();
                }
                // ---- MMA1(i+1) ahead of the weight-gradient MMAs: it only needs the next rbf tile and a drained D1, and the
                // epilogue warps -- the bottleneck -- can then start E1(i+1) the moment they finish E3(i)
                if (i + 1 < my_tiles) mma1(i + 1);
                // ---- WG2(i): DW2 += dU^T s
                mbar_wait(bar(S_FULL_ + b), (i >> 1) & 1);
                tc_fence_after();
                trace_b(i, 5);
                {
                    const uint64_t uh = desc_mn_sw128(du, kBlkT), ul = desc_mn_sw128(du + 2 * kBlkT, kBlkT);
                    const uint64_t sh = desc_mn_sw128(sa, kBlkT), sl = desc_mn_sw128(sa + 2 * kBlkT, kBlkT);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t o = ks * (2048 >> 4);
                            mma3(tDW2, uh + o, ul + o, sh + o, sl + o, id_w2, (i | ks) > 0);
                        }
                        tc_commit(bar(S_FREE_ + b));
                        tc_commit(bar(DU_FREE_ + b));
                    }
                    __syncwarp();
                }
                trace_b(i, 6);
                // ---- WG1(i): DW1 += dA^T rbf
                mbar_wait(bar(DA_FULL_), i & 1);
                tc_fence_after();
                trace_b(i, 7);
                {
                    const uint32_t ph_addr = sbase + L::PHI + b * 2 * kBlkT;
                    const uint64_t ph = desc_mn_sw128(ph_addr, kBlkT), pl = desc_mn_sw128(ph_addr + kBlkT, kBlkT);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t o = ks * (2048 >> 4);
                            mma3(tDW1, dAh + o, dAl + o, ph + o, pl + o, id_w1, (i | ks) > 0);
                        }
                        tc_commit(bar(DA_FREE_));
                        tc_commit(bar(PHI_FREE_ + b));
                    }
                    __syncwarp();
                }
                trace_b(i, 8);
            }
            if (leader) tc_commit(bar(DONE_));
            __syncwarp();
        }
    } else {
        // ===================== epilogue warps (TMEM lane = feature, columns = edges): warp = (quadrant q, edge quarter eq)
        const int q = warp & 3, eq = warp >> 2;
        const int f = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const float b1f = __ldg(b1 + f);
        const uint32_t kcol = f & 63;
        const int blk = (f >> 6) * kBlkT;
        float sig[16];
        auto e1 = [&](int i) {
            const int b = i & 1;
            mbar_wait(bar(D1_FULL_), i & 1);
            tc_fence_after();
            if (warp == 0) trace_b(i, 9);
            float a[16];
            tmem_ld16(tD1 + lane_base + eq * 16, a);
            tc_fence_before();
            warp_arrive(bar(D1_FREE_));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float av = a[j] + b1f;
                const float t = ex2_approx(-fabsf(av) * 1.4426950408889634f);       // exp(-|a|)
                const float r = rcp_approx(1.0f + t);
                a[j] = fmaf(lg2_approx(1.0f + t), 0.6931471805599453f, fmaxf(av, 0.f)) - kLog2;
                sig[j] = av >= 0.f ? r : t * r;                                      // sigmoid(a): three MUFU per element in all
            }
            mbar_wait(bar(S_FREE_ + b), ((i >> 1) & 1) ^ 1);             // WG2(i-2) has consumed this S buffer
            uint8_t* s_hi = smem + L::S + b * 4 * kBlkT + blk;
            uint8_t* s_lo = s_hi + 2 * kBlkT;
            store_split16_paired<kBwdFP16>(s_hi, s_lo, eq * 16, kcol, a, lane);
            fence_proxy_async();
            warp_arrive(bar(S_FULL_ + b));
            if (warp == 0) trace_b(i, 10);
        };
        uint8_t* da_hi = smem + L::DA + blk;
        uint8_t* da_lo = da_hi + 2 * kBlkT;
        if (my_tiles > 0) e1(0);
        for (int i = 0; i < my_tiles; ++i) {
            mbar_wait(bar(D3_FULL_), i & 1);
            tc_fence_after();
            if (warp == 0) trace_b(i, 11);
            float ds[16];
            tmem_ld16(tD3 + lane_base + eq * 16, ds);
            tc_fence_before();
            warp_arrive(bar(D3_FREE_));
            if (i > 0) mbar_wait(bar(DA_FREE_), (i - 1) & 1);             // WG1(i-1) has consumed dA
            if (warp == 0) trace_b(i, 12);
#pragma unroll
            for (int j = 0; j < 16; ++j) ds[j] *= sig[j];
            store_split16_paired<kBwdFP16>(da_hi, da_lo, eq * 16, kcol, ds, lane);
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

// Sum of the per-CTA partials in a fixed order (deterministic).  Four adjacent lanes share one output and take every fourth
// partial each (two independent chains per lane keep loads in flight), then combine by two shuffles: 4x the parallelism of
// one thread per output for the same 15 MB read (the kernel sits on the critical path of every layer's backward).
__global__ void __launch_bounds__(256)
filter_bwd_tc_reduce_kernel(const float* __restrict__ workspace, int n_parts, int G,
                            float* __restrict__ gw1, float* __restrict__ gb1,
                            float* __restrict__ gw2, float* __restrict__ gb2) {
    pdl_launch_dependents();
    pdl_wait();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int idx = gid >> 2, part = gid & 3;
    const bool live = idx < Part::kFloats;
    const bool w1_slot = live && idx >= Part::kW1 && idx < Part::kB2;
    const int g = w1_slot ? (idx - Part::kW1) / 128 : 0;
    float s0 = 0.f, s1 = 0.f;
    if (live && !(w1_slot && g >= G)) {                                  // padded gaussians are never written
        int p = part;
        for (; p + 4 < n_parts; p += 8) {
            s0 += workspace[(int64_t)p * Part::kFloats + idx];
            s1 += workspace[(int64_t)(p + 4) * Part::kFloats + idx];
        }
        if (p < n_parts) s0 += workspace[(int64_t)p * Part::kFloats + idx];
    }
    float s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (!live || part != 0 || (w1_slot && g >= G)) return;
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
                         const int32_t* pair_atoms,
                         float* workspace, float* gw1, float* gb1, float* gw2, float* gb2, void* stream) {
    GEOSSL_REQUIRE(edge_dist && n_edges_dev && offset && w1 && b1 && w2 && x && grad_out && workspace &&
                   gw1 && gb1 && gw2 && gb2 && (pair_atoms || (src && edge_tgt)), "null pointer");
    GEOSSL_REQUIRE(F == 128, "the tensor-core filter kernel is built for num_filters = 128");
    GEOSSL_REQUIRE(G >= 1 && G <= 63, "num_gaussians must be in [1,63] (column 63 of the rbf tile carries the bias sum)");
    const size_t smem = tc::BwdLayout::kBytes + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::filter_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::filter_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    if (pair_atoms) {
        GEOSSL_CUDA(launch_pdl(tc::filter_bwd_tc_kernel<true>, dim3(kNumSM), dim3(tc::kBwdThreads), smem, as_stream(stream),
                               edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, x, grad_out, src, edge_tgt,
                               reinterpret_cast<const int2*>(pair_atoms), workspace));
    } else {
        GEOSSL_CUDA(launch_pdl(tc::filter_bwd_tc_kernel<false>, dim3(kNumSM), dim3(tc::kBwdThreads), smem, as_stream(stream),
                               edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, x, grad_out, src, edge_tgt,
                               (const int2*)nullptr, workspace));
    }
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(tc::filter_bwd_tc_reduce_kernel, dim3((4 * tc::Part::kFloats + 255) / 256), dim3(256), 0, as_stream(stream),
                           workspace, (int)kNumSM, G, gw1, gb1, gw2, gb2));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
