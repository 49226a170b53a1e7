// Node-level dense layers (128 -> 128) on the sm_100a tensor cores: forward / data-gradient and
// weight-gradient kernels with split-precision operands (tc.cuh), fused bias, shifted-softplus
// pre-activation, activation-gradient and residual epilogues.
//
// Replaces the atom-wise nn.Linear calls of the reference encoder (Geom3D/models/schnet.py:99-101 head,
// :165-166 InteractionBlock act + lin, :189,191 CFConv lin1 / lin2, and the `h + interaction(...)` residual
// of :97), which run as fp32 SIMT cuBLAS GEMMs + separate elementwise kernels in PyTorch.
//
//   linear_tc:        Y[r][n] = epi( sum_k pre(X[r][k]) * Wm[n][k] )
//                       pre  in {identity, shifted softplus}
//                       Wm   = W (nn.Linear layout, forward) or W^T (data gradient)
//                       epi  : + bias[n], * sigmoid(Z[r][n]) (activation gradient), + R[r][n] (residual)
//                     computed transposed (D^T = Wm . X^T: TMEM lanes = output features, columns = rows) so
//                     that bias is a per-thread scalar and every store instruction writes 128 contiguous bytes.
//   linear_wgrad_tc:  DW[o][i] = sum_r dY[r][o] * pre(X[r][i]),  db[o] = sum_r dY[r][o]
//                     MN-major views of [row][feature] tiles, accumulator resident in TMEM, per-CTA partials
//                     reduced in a fixed order (deterministic).
#include "common.cuh"
#include "tc.cuh"

namespace geossl {
namespace tc {

__device__ long long* g_trace_lin = nullptr;   // optional clock64() trace of CTA 0 (geossl_debug_set_trace_linear)
__device__ __forceinline__ void trace_l(int event) {
    long long* t = g_trace_lin;
    if (t != nullptr && blockIdx.x == 0 && threadIdx.x == 0) t[event] = clock64();
}

// Pre-activation applied to X while it is staged (and its derivative in the data-gradient epilogue):
// 1 = shifted softplus (SchNet, schnet.py:210-216), 2 = SiLU (PaiNN's Dense layers, painn.py:21-24 / painn_utils.py:9-35).
constexpr int kActSsp = 1, kActSilu = 2;
__device__ __forceinline__ float act_fwd(float x, int kind) {
    return kind == kActSilu ? x * sigmoid_fast(x) : ssp_fast(x);
}
__device__ __forceinline__ float act_grad(float z, int kind) {
    const float sg = sigmoid_fast(z);
    return kind == kActSilu ? sg * fmaf(z, 1.f - sg, 1.f) : sg;
}

constexpr int kNR = 64;                   // rows per tile
constexpr int kNBlkW = 128 * 128;         // [128 rows x 64 k] 16-bit weight block
constexpr int kNBlkT = kNR * 128;         // [64 rows x 64 k] 16-bit tile block

template <int NR>                                     // NR rows per tile (64: two CTAs per SM; 128: one wave of 1-CTA SMs)
struct LinLayoutT {
    static constexpr int W = 0;                       // hi (2 k-blocks) | lo (2 k-blocks)   64 KB
    static constexpr int X = W + 4 * kNBlkW;          // hi (2 k-blocks) | lo (2 k-blocks)   32 / 64 KB
    static constexpr int BAR = X + 4 * NR * 128;          // [0] MMA done, [1] weight image landed
    static constexpr int TMEM_PTR = BAR + 16;
    static constexpr int kBytes = TMEM_PTR + 16;
};
constexpr int kWImage = 4 * kNBlkW;                   // bytes of a packed weight image: hi (2 k-blocks) | lo (2 k-blocks)

// Weight (128,128) fp32 -> the exact shared-memory operand image (split, swizzled) in global memory, so that every CTA
// of the layer kernels fetches it with one bulk async copy instead of re-splitting it.
template <bool FP16>
__device__ __forceinline__ void pack_weight_body(const float* __restrict__ Wt, int ldw, int trans, uint8_t* __restrict__ image) {
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 128 * 16; idx += gridDim.x * blockDim.x) {
        float v[8];
        int n, c;
        if (!trans) {
            n = idx >> 4; c = idx & 15;
            if ((ldw & 3) == 0) {
                const float4 a = ldg4(Wt + (size_t)n * ldw + c * 8), b = ldg4(Wt + (size_t)n * ldw + c * 8 + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {                                       // rows not 16-byte aligned (the (128,129) first layer of the DDM head)
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __ldg(Wt + (size_t)n * ldw + c * 8 + j);
            }
        } else {
            n = idx & 127; c = idx >> 7;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldg(Wt + (size_t)(c * 8 + j) * ldw + n);
        }
        const int blk = c >> 3;
        store_chunk8<FP16>(image + blk * kNBlkW, image + 2 * kNBlkW + blk * kNBlkW, n, (c & 7) * 8, v);
    }
}

template <bool FP16>
__global__ void __launch_bounds__(256)
pack_weight_kernel(const float* __restrict__ Wt, int ldw, int trans, uint8_t* __restrict__ image) {
    pdl_launch_dependents();
    pdl_wait();
    pack_weight_body<FP16>(Wt, ldw, trans, image);
}

// Both images of ONE 128 x 128 block (forward, fp16 parts | transposed, bf16 parts) in one launch: blockIdx.y = orientation.
__global__ void __launch_bounds__(256)
pack_weight_pair_kernel(const float* __restrict__ Wt, int ldw, uint8_t* __restrict__ images) {
    pdl_launch_dependents();
    pdl_wait();
    if (blockIdx.y == 0) pack_weight_body<true>(Wt, ldw, 0, images);
    else pack_weight_body<false>(Wt, ldw, 1, images + kWImage);
}

// All layers of a model in ONE launch: blockIdx.y = layer, blockIdx.z = 0: forward image (fp16 parts, nn.Linear
// orientation) / 1: data-gradient image (bf16 parts, transposed).  images = [layer][2][kWImage] bytes.
__global__ void __launch_bounds__(256)
pack_weights_batched_kernel(const float* const* __restrict__ weights, const int32_t* __restrict__ lds, uint8_t* __restrict__ images) {
    // weights[b] = first element of a 128 x 128 block of a row-major matrix with leading dimension lds[b] (NULL => 128)
    const float* Wt = weights[blockIdx.y];
    const int ldw = lds ? lds[blockIdx.y] : 128;
    uint8_t* image = images + ((size_t)blockIdx.y * 2 + blockIdx.z) * kWImage;
    if (blockIdx.z == 0) pack_weight_body<true>(Wt, ldw, 0, image);
    else pack_weight_body<false>(Wt, ldw, 1, image);
}

// X tile (NR rows x 128 k, fp32, optional ssp) -> split K-major SW128 image.  4 NR threads, four (row, 8-column chunk)
// items each: 16 lanes cover one 512-byte row, so every load instruction of a warp reads 1 KB of contiguous memory
// (full 32-byte sectors) and all eight loads of a thread are in flight before the first conversion.
template <bool FP16, int NR>
__device__ __forceinline__ void stage_rows_kmajor(const float* __restrict__ X, int64_t ldx, int64_t row0, int64_t n_rows, int pre_act,
                                                  uint8_t* hi, uint8_t* lo, int k_cols = 128) {
    constexpr int kBlkT = NR * 128, kRowsPerPass = NR / 4;
    const int tid = threadIdx.x, c = tid & 15, r0 = tid >> 4;
    float4 a[4][2];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int64_t row = row0 + r0 + it * kRowsPerPass;
        if (row < n_rows && c * 8 < k_cols) {                 // (k_cols < 128: a K-padded operand, only its live columns exist)
            const float* p = X + row * ldx + c * 8;
            a[it][0] = ldg4(p);
            a[it][1] = ldg4(p + 4);
        } else {
            a[it][0] = a[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        float v[8] = {a[it][0].x, a[it][0].y, a[it][0].z, a[it][0].w, a[it][1].x, a[it][1].y, a[it][1].z, a[it][1].w};
        if (pre_act && row0 + r0 + it * kRowsPerPass < n_rows) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = act_fwd(v[k], pre_act);
        }
        const int blk = c >> 3;                       // k-block of 64
        store_chunk8<FP16>(hi + blk * kBlkT, lo + blk * kBlkT, r0 + it * kRowsPerPass, (c & 7) * 8, v);
    }
}

template <bool FP16, bool PRE_SSP, bool HAS_Z, bool HAS_R, int NR>
__global__ void __launch_bounds__(4 * NR, NR == 64 ? 2 : 1)
linear_tc_kernel(const float* __restrict__ X, int64_t n_rows, const uint8_t* __restrict__ w_image, const float* __restrict__ bias,
                 const float* __restrict__ Z, const float* __restrict__ R, float* __restrict__ Y,
                 int64_t ldx, int64_t ldz, int64_t ldr, int64_t ldy, int act, int k_cols) {
    // ld*: row strides in floats (128 for the plain 128 -> 128 layer; wider layers are tiled into 128 x 128 blocks by the
    // host, each block one launch over a column window of the wider tensors).  act: kActSsp / kActSilu where PRE_SSP / HAS_Z apply.
    extern __shared__ uint8_t smem_raw[];
    trace_l(0);
    uint8_t* smem = align1024(smem_raw);
    using L = LinLayoutT<NR>;
    constexpr int kBlkT = NR * 128;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + L::BAR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // weights: the packed A-operand image (rows n, K-major, split) arrives by bulk async copy while the X tile is staged
    const uint32_t wbar = bar + 8;
    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(wbar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(sbase + L::TMEM_PTR, NR);
    pdl_wait();                                                // everything above overlaps the previous kernel's tail
    if (tid == 0) {
        mbar_expect_tx(wbar, kWImage);
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(sbase + L::W + c * kNBlkW, w_image + c * kNBlkW, kNBlkW, wbar);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, NR);
    const int q = warp & 3, eh = warp >> 2, f = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const float bf = bias ? __ldg(bias + f) : 0.f;

    const int64_t n_tiles = (n_rows + NR - 1) / NR;
    uint32_t phase = 0;
    trace_l(1);
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t row0 = t * NR;
        stage_rows_kmajor<FP16, NR>(X, ldx, row0, n_rows, PRE_SSP ? act : 0, smem + L::X, smem + L::X + 2 * kBlkT, k_cols);
        trace_l(2);
        // epilogue operands do not depend on the MMA: fetch them now so their latency hides behind it
        const bool full = row0 + NR <= n_rows;                 // warp-uniform: no per-row bounds checks on full tiles
        const int64_t erow = row0 + eh * 32;
        const int64_t ebase = erow * ldy + f;
        float zr[HAS_Z ? 32 : 1], rr[HAS_R ? 32 : 1];
        if constexpr (HAS_Z) {
#pragma unroll
            for (int j = 0; j < 32; ++j) zr[j] = (full || erow + j < n_rows) ? __ldg(Z + (erow + j) * ldz + f) : 0.f;
        }
        if constexpr (HAS_R) {
#pragma unroll
            for (int j = 0; j < 32; ++j) rr[j] = (full || erow + j < n_rows) ? __ldg(R + (erow + j) * ldr + f) : 0.f;
        }
        fence_proxy_async();
        __syncthreads();
        trace_l(3);
        if (warp == 0) {                                       // whole warp (uniform operands); the elected lane issues
            mbar_wait(wbar, 0);
            trace_l(4);                                // (completes once; later tiles pass immediately)
            tc_fence_after();
            const uint64_t wh = desc_k_sw128(sbase + L::W), wl = desc_k_sw128(sbase + L::W + 2 * kNBlkW);
            const uint64_t xh = desc_k_sw128(sbase + L::X), xl = desc_k_sw128(sbase + L::X + 2 * kBlkT);
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    if (ks * 16 >= k_cols) break;              // K-padded operand: the remaining k-steps multiply zeros
                    const uint32_t ow = (ks >> 2) * (kNBlkW >> 4) + 2 * (ks & 3), ox = (ks >> 2) * (kBlkT >> 4) + 2 * (ks & 3);
                    mma3(tmem, wh + ow, wl + ow, xh + ox, xl + ox, idesc, ks > 0);
                }
                tc_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        trace_l(5);
        float v[32];
        tmem_ld32(tmem + lane_base + eh * 32, v);          // lane = output feature f, columns = rows eh*32..+31
        tc_fence_before();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float y = v[j] + bf;
            if constexpr (HAS_Z) y *= act_grad(zr[j], act);
            if constexpr (HAS_R) y += rr[j];
            v[j] = y;
        }
        if (full) {
#pragma unroll
            for (int j = 0; j < 32; ++j) Y[ebase + j * ldy] = v[j];                // 128 contiguous bytes per warp store
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (row0 + eh * 32 + j < n_rows) Y[ebase + j * ldy] = v[j];
        }
        __syncthreads();                                   // TMEM / X tile reuse by the next tile
        trace_l(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, NR);
    }
    trace_l(7);
}

// ------------------------------------------------------------------------------------------ chained layers
// Several 128 -> 128 layers applied to the SAME row tile without leaving the SM.  An interaction block of SchNet applies
// conv.lin2 -> ssp -> lin (+ residual) and then the next block's conv.lin1 to every atom row (schnet.py:191,165-166,97,189);
// the head applies lin1 -> ssp -> lin2 (:99-101).  As separate launches each of these ~0.5 GFLOP layers is latency bound
// (launch + weight fetch + row staging + one MMA burst + store: 12 us inside the step for ~1 us of tensor work).  Here the
// row tile is staged once, each stage's accumulator is read back from TMEM, finished (bias / activation gradient / residual),
// optionally stored, and written straight back into shared memory as the next stage's operand (paired 4-byte stores,
// tc.cuh), while the weight images stream through a two-slot ring of bulk copies.  The data-gradient chain of the backward
// pass (next lin1^T -> + residual gradient -> lin^T -> * sigmoid(y1) -> lin2^T) is the same kernel with transposed images.
struct ChainStage {
    const uint8_t* w_image;   // packed weight image (orientation / parts chosen when it was packed)
    const float* bias;        // NULL or (128)
    const float* z;           // NULL or rows of 128 (stride ldz): v *= act'(z)
    const float* residual;    // NULL or rows of 128 (stride ldr): v += residual
    float* store;             // NULL or rows of 128 (stride lds): store = v
    const float* x;           // NULL: the operand is the previous stage's result (or, with `keep`, the previous stage's operand);
                              // else this stage's operand tile is loaded from here (stride ldx, pre-activation x_act)
    int64_t ldz, ldr, lds, ldx;
    int act_next;             // 0 | kActSsp | kActSilu: activation applied to v before it becomes the next stage's operand
    int x_act;                // pre-activation applied to a loaded operand (0 | kActSsp | kActSilu)
    int keep;                 // the operand tile of the previous stage is used again (fan-out: one input, several weight blocks)
    int accumulate;           // the MMAs add to the previous (partial) stage's accumulator (fan-in: K > 128)
    int partial;              // no epilogue: the next stage accumulates onto this one
};
constexpr int kMaxChain = 6;
struct ChainArgs { ChainStage st[kMaxChain]; int n; };

struct ChainLayout {
    static constexpr int NR = 128;
    static constexpr int W = 0;                       // two slots x (hi 2 k-blocks | lo 2 k-blocks)   128 KB
    static constexpr int X = W + 2 * kWImage;         // hi (2 k-blocks) | lo (2 k-blocks)             64 KB
    static constexpr int BAR = X + 4 * NR * 128;      // [0] MMA done, [1..2] weight slot landed
    static constexpr int TMEM_PTR = BAR + 32;
    static constexpr int kBytes = TMEM_PTR + 16;
};
static_assert(ChainLayout::kBytes + 1024 <= 227 * 1024, "shared memory budget");

// EX = false: the plain chain (every tensor (n_rows,128) contiguous, operands handed on from stage to stage) compiles without
// the stride / fan-out / fan-in machinery -- it is the kernel of SchNet's step and runs 14 times in it.
template <bool FP16, bool EX>
__global__ void __launch_bounds__(512, 1)
linear_chain_tc_kernel(const float* __restrict__ X, int64_t n_rows, const __grid_constant__ ChainArgs args, int act) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    using L = ChainLayout;
    constexpr int NR = L::NR, kBlkT = NR * 128;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + L::BAR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar + 8, 1);
        mbar_init(bar + 16, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(sbase + L::TMEM_PTR, NR);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, NR);
    const int q = warp & 3, eh = warp >> 2, f = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int n_st = args.n;

    auto fetch_weights = [&](int stage, int slot) {           // one thread: 64 KB image -> ring slot, completion on its mbarrier
        mbar_expect_tx(bar + 8 + 8 * slot, kWImage);
#pragma unroll
        for (int c = 0; c < 4; ++c) bulk_g2s(sbase + L::W + slot * kWImage + c * kNBlkW, args.st[stage].w_image + c * kNBlkW, kNBlkW, bar + 8 + 8 * slot);
    };

    const int64_t n_tiles = (n_rows + NR - 1) / NR;
    uint32_t mma_phase = 0, w_uses[2] = {0, 0};
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t row0 = t * NR;
        const bool full = row0 + NR <= n_rows;
        if (tid == 0) {
            fetch_weights(0, 0);
            if (n_st > 1) fetch_weights(1, 1);
        }
        stage_rows_kmajor<FP16, NR>(X, EX ? args.st[0].ldx : 128, row0, n_rows, EX ? args.st[0].x_act : 0, smem + L::X,
                                    smem + L::X + 2 * kBlkT);
        const int64_t erow = row0 + eh * 32;
        for (int s = 0; s < n_st; ++s) {
            const ChainStage& S = args.st[s];
            const int slot = s & 1;
            // a stage that brings its own operand: every thread has passed the previous stage's MMA wait, the tile is free
            if (EX && s > 0 && S.x != nullptr)
                stage_rows_kmajor<FP16, NR>(S.x, S.ldx, row0, n_rows, S.x_act, smem + L::X, smem + L::X + 2 * kBlkT);
            const bool partial = EX && S.partial;
            const int64_t ldz = EX ? S.ldz : 128, ldr = EX ? S.ldr : 128, lds = EX ? S.lds : 128;
            // epilogue operands that do not depend on the MMA: fetched now, their latency hides behind it
            float zr[32], rr[32];
            const bool has_z = S.z != nullptr && !partial, has_r = S.residual != nullptr && !partial;
            if (has_z) {
#pragma unroll
                for (int j = 0; j < 32; ++j) zr[j] = (full || erow + j < n_rows) ? __ldg(S.z + (erow + j) * ldz + f) : 0.f;
            }
            if (has_r) {
#pragma unroll
                for (int j = 0; j < 32; ++j) rr[j] = (full || erow + j < n_rows) ? __ldg(S.residual + (erow + j) * ldr + f) : 0.f;
            }
            const float bf = S.bias ? __ldg(S.bias + f) : 0.f;
            fence_proxy_async();
            __syncthreads();                                    // operand tile of this stage is complete
            if (warp == 0) {
                mbar_wait(bar + 8 + 8 * slot, w_uses[slot] & 1);
                tc_fence_after();
                const uint32_t wb = sbase + L::W + slot * kWImage;
                const uint64_t wh = desc_k_sw128(wb), wl = desc_k_sw128(wb + 2 * kNBlkW);
                const uint64_t xh = desc_k_sw128(sbase + L::X), xl = desc_k_sw128(sbase + L::X + 2 * kBlkT);
                if (elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const uint32_t ow = (ks >> 2) * (kNBlkW >> 4) + 2 * (ks & 3), ox = (ks >> 2) * (kBlkT >> 4) + 2 * (ks & 3);
                        mma3(tmem, wh + ow, wl + ow, xh + ox, xl + ox, idesc, (ks > 0) || (EX && S.accumulate));
                    }
                    tc_commit(bar);
                }
                __syncwarp();
            }
            ++w_uses[slot];
            mbar_wait(bar, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
            // the MMAs have consumed this slot's image and the operand tile: refill the slot two stages ahead
            if (tid == 0 && s + 2 < n_st) fetch_weights(s + 2, slot);
            if (partial) continue;                                 // the next stage accumulates onto this result
            float v[32];
            tmem_ld32(tmem + lane_base + eh * 32, v);              // lane = output feature f, columns = rows eh*32..+31
            tc_fence_before();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float y = v[j] + bf;
                if (has_z) y *= act_grad(zr[j], act);
                if (has_r) y += rr[j];
                v[j] = y;
            }
            if (S.store != nullptr) {
                float* out = S.store + erow * lds + f;
                if (full) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) out[j * lds] = v[j];  // 128 contiguous bytes per warp store
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (erow + j < n_rows) out[j * lds] = v[j];
                }
            }
            if (s + 1 < n_st && (!EX || (args.st[s + 1].x == nullptr && !args.st[s + 1].keep))) {
                // next stage's operand: element (row = eh*32 + j, k = f) of the K-major tile
                if (S.act_next) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (full || erow + j < n_rows) ? act_fwd(v[j], S.act_next) : 0.f;
                } else if (!full) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = (erow + j < n_rows) ? v[j] : 0.f;
                }
                uint8_t* hi = smem + L::X + (f >> 6) * kBlkT;
                uint8_t* lo = hi + 2 * kBlkT;
                float half[16];
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) half[j] = v[h2 * 16 + j];
                    store_split16_paired<FP16>(hi, lo, eh * 32 + h2 * 16, f & 63, half, lane);
                }
            }
        }
        __syncthreads();                                           // TMEM / tile / barrier reuse by the next tile
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, NR);
    }
}

// ------------------------------------------------------------------------------------------ weight gradient
struct WgLayout {
    static constexpr int DY = 0;                      // [64 r][128 o] hi (2 MN blocks) | lo                32 KB
    static constexpr int XT = DY + 4 * kNBlkT;        // [64 r][128 i] hi | lo                              32 KB
    static constexpr int RED = XT + 4 * kNBlkT;       // 16 x 128 floats (bias column sums)                  8 KB
    static constexpr int BAR = RED + 16 * 128 * 4;
    static constexpr int TMEM_PTR = BAR + 16;
    static constexpr int kBytes = TMEM_PTR + 16;
};
constexpr int kWgPart = 128 * 128 + 128;              // per-CTA partial: DW [o][i] then db [o]

// [row][feature] tile -> split image (same memory layout as K-major rows; read by the MMA as MN-major).
// thread = (8 columns 8cg.., rows ro + 16 s); returns the per-thread column sums in acc when requested.
template <bool FP16, bool SUM>
__device__ __forceinline__ void stage_rows_mn(const float* __restrict__ X, int64_t ldx, int64_t row0, int64_t n_rows, int pre_act,
                                              uint8_t* hi, uint8_t* lo, float (&acc)[8], int x_cols = 128) {
    const int tid = threadIdx.x, cg = tid & 15, ro = tid >> 4;
    float4 a[4][2];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int64_t row = row0 + s * 16 + ro;
        if (row < n_rows && cg * 8 < x_cols) {
            const float* p = X + row * ldx + cg * 8;
            a[s][0] = ldg4(p); a[s][1] = ldg4(p + 4);
        } else {
            a[s][0] = a[s][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        float v[8] = {a[s][0].x, a[s][0].y, a[s][0].z, a[s][0].w, a[s][1].x, a[s][1].y, a[s][1].z, a[s][1].w};
        const bool valid = row0 + s * 16 + ro < n_rows;
        if (pre_act) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = valid ? act_fwd(v[k], pre_act) : 0.f;
        }
        if (SUM) {
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += v[k];
        }
        store_chunk8<FP16>(hi + (cg >> 3) * kNBlkT, lo + (cg >> 3) * kNBlkT, s * 16 + ro, (cg & 7) * 8, v);
    }
}

// Up to kMaxWgBatch independent weight-gradient products over the SAME rows in one launch (blockIdx.y = problem): the three
// layers of an interaction tail, the blocks of PaiNN's wider Dense layers, ... -- each of them is a ~10 us latency-bound
// launch on its own, and they come in groups that are ready at the same time.
constexpr int kMaxWgBatch = 6;
struct WgProblem {
    const float* dY; const float* X;
    float* workspace; float* gw; float* gb;
    int64_t ld_dy, ld_x;
    int pre_act, x_cols, ld_gw;
};
struct WgBatch { WgProblem p[kMaxWgBatch]; };

template <bool FP16>
__global__ void __launch_bounds__(256, 2)
linear_wgrad_tc_kernel(const __grid_constant__ WgBatch batch, int64_t n_rows) {
    const WgProblem& P = batch.p[blockIdx.y];
    const float* __restrict__ dY = P.dY;
    const float* __restrict__ X = P.X;
    float* __restrict__ workspace = P.workspace;
    const int64_t ld_dy = P.ld_dy, ld_x = P.ld_x;
    const int pre_act = P.pre_act, x_cols = P.x_cols;
    extern __shared__ uint8_t smem_raw[];
    pdl_launch_dependents();
    pdl_wait();
    uint8_t* smem = align1024(smem_raw);
    using L = WgLayout;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + L::BAR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(sbase + L::TMEM_PTR, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    const uint32_t idesc = idesc_f16(Split<FP16>::kFmt, 128, 128, 1, 1);
    float acc[8], dummy[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;

    const int64_t n_tiles = (n_rows + kNR - 1) / kNR;
    uint32_t phase = 0;
    int done = 0;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++done) {
        const int64_t row0 = t * kNR;
        stage_rows_mn<FP16, true>(dY, ld_dy, row0, n_rows, 0, smem + L::DY, smem + L::DY + 2 * kNBlkT, acc);
        stage_rows_mn<FP16, false>(X, ld_x, row0, n_rows, pre_act, smem + L::XT, smem + L::XT + 2 * kNBlkT, dummy, x_cols);
        fence_proxy_async();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const uint64_t ah = desc_mn_sw128(sbase + L::DY, kNBlkT), al = desc_mn_sw128(sbase + L::DY + 2 * kNBlkT, kNBlkT);
            const uint64_t bh = desc_mn_sw128(sbase + L::XT, kNBlkT), bl = desc_mn_sw128(sbase + L::XT + 2 * kNBlkT, kNBlkT);
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t o = ks * (2048 >> 4);
                    mma3(tmem, ah + o, al + o, bh + o, bl + o, idesc, (done | ks) > 0);
                }
                tc_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase);                               // tiles are re-staged only after the MMAs have read them
        phase ^= 1;
    }
    tc_fence_after();
    // ---- partials: DW from TMEM (lane = o, columns = i), db from the staged column sums
    float* ws = workspace + (int64_t)blockIdx.x * kWgPart;
    const int q = warp & 3, eh = warp >> 2, o = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        float v[32];
        if (done > 0) {
            tmem_ld32(tmem + lane_base + eh * 64 + h * 32, v);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
        }
        float* dst = ws + o * 128 + eh * 64 + h * 32;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    float* red = reinterpret_cast<float*>(smem + L::RED);
    {
        const int cg = tid & 15, ro = tid >> 4;
#pragma unroll
        for (int k = 0; k < 8; ++k) red[ro * 128 + cg * 8 + k] = acc[k];
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 128) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) s += red[r * 128 + tid];
        ws[128 * 128 + tid] = s;
    }
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, 128);
    }
}

// 256 threads = 32 outputs x 8 slices of the partial list; slices are combined in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
linear_wgrad_reduce_kernel(const __grid_constant__ WgBatch batch, int n_parts) {
    const WgProblem& P = batch.p[blockIdx.y];
    const float* __restrict__ workspace = P.workspace;
    float* __restrict__ gw = P.gw;
    float* __restrict__ gb = P.gb;
    const int ld_gw = P.ld_gw;
    __shared__ float red[8][33];
    pdl_launch_dependents();
    pdl_wait();

    const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + o;
    float s0 = 0.f, s1 = 0.f;
    if (idx < kWgPart) {
        int p = sl;
        for (; p + 8 < n_parts; p += 16) {
            s0 += workspace[(int64_t)p * kWgPart + idx];
            s1 += workspace[(int64_t)(p + 8) * kWgPart + idx];
        }
        if (p < n_parts) s0 += workspace[(int64_t)p * kWgPart + idx];
    }
    red[sl][o] = s0 + s1;
    __syncthreads();
    if (sl == 0 && idx < kWgPart) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) s += red[k][o];
        if (idx < 128 * 128) gw[(size_t)(idx >> 7) * ld_gw + (idx & 127)] = s;
        else if (gb) gb[idx - 128 * 128] = s;
    }
}

static int node_grid(int64_t n_rows) {
    int64_t tiles = (n_rows + kNR - 1) / kNR;
    const int cap = 2 * kNumSM;                              // two CTAs per SM fit (96 KB / 72 KB of shared memory)
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

// Weight-gradient kernels run on the side stream next to the main backward chain and every CTA dumps a 66 KB partial:
// a smaller grid (several row tiles per CTA, accumulated in TMEM) cuts the partial traffic and leaves SMs to the main chain.
static int wgrad_grid(int64_t n_rows) {
    int64_t tiles = (n_rows + kNR - 1) / kNR;
    static const int cap = [] { const char* e = getenv("GEOSSL_WGRAD_CTAS"); return e ? atoi(e) : 120; }();
    if (tiles >= 8 * kNumSM) return kNumSM;                  // edge-sized inputs (PaiNN's filter GEMM): a full persistent grid
    return (int)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
}

}  // namespace tc
}  // namespace geossl

using namespace geossl;

template <bool FP16, bool PRE_SSP, bool HAS_Z, bool HAS_R>
static int launch_linear_tc(const float* x, int64_t n_rows, const uint8_t* weight, const float* bias, const float* z, const float* r,
                            float* y, int64_t ldx, int64_t ldz, int64_t ldr, int64_t ldy, int act, int k_cols, cudaStream_t st) {
    static const int wide_min = [] { const char* e = getenv("GEOSSL_LINEAR_WIDE_MIN"); return e ? atoi(e) : 64 * kNumSM; }();
    if (n_rows >= wide_min) {
        // enough rows for every SM: 128-row tiles, one 512-thread CTA per SM (no co-resident CTA competing for the
        // tensor pipe / shared memory during the same phases), half as many weight-image fetches
        constexpr int NR = 128;
        const size_t smem = tc::LinLayoutT<NR>::kBytes + 1024;
        static PerDeviceFlag configured;                        // one flag per instantiation
        if (!configured.get()) {
            GEOSSL_CUDA(cudaFuncSetAttribute(tc::linear_tc_kernel<FP16, PRE_SSP, HAS_Z, HAS_R, NR>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured.set();
        }
        const int64_t tiles = (n_rows + NR - 1) / NR;
        GEOSSL_CUDA(launch_pdl(tc::linear_tc_kernel<FP16, PRE_SSP, HAS_Z, HAS_R, NR>, dim3((unsigned)(tiles < kNumSM ? tiles : kNumSM)),
                               dim3(4 * NR), smem, st, x, n_rows, weight, bias, z, r, y, ldx, ldz, ldr, ldy, act, k_cols));
        GEOSSL_LAUNCH_CHECK();
        return 0;
    }
    constexpr int NR = 64;
    const size_t smem = tc::LinLayoutT<NR>::kBytes + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::linear_tc_kernel<FP16, PRE_SSP, HAS_Z, HAS_R, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        configured.set();
    }
    GEOSSL_CUDA(launch_pdl(tc::linear_tc_kernel<FP16, PRE_SSP, HAS_Z, HAS_R, NR>, dim3(tc::node_grid(n_rows)), dim3(4 * NR), smem, st,
                           x, n_rows, weight, bias, z, r, y, ldx, ldz, ldr, ldy, act, k_cols));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

template <bool FP16, bool EX>
static int launch_chain_t(const float* x, int64_t n_rows, const tc::ChainArgs& a, int act, void* stream) {
    const size_t smem = tc::ChainLayout::kBytes + 1024;
    const int64_t tiles = (n_rows + 127) / 128;
    const dim3 grid((unsigned)(tiles < kNumSM ? tiles : kNumSM));
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::linear_chain_tc_kernel<FP16, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    GEOSSL_CUDA(launch_pdl(tc::linear_chain_tc_kernel<FP16, EX>, grid, dim3(512), smem, as_stream(stream), x, n_rows, a, act));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

static int launch_chain(const float* x, int64_t n_rows, const tc::ChainArgs& a, int bf16_parts, int act, bool ex, void* stream) {
    if (ex) return bf16_parts ? launch_chain_t<false, true>(x, n_rows, a, act, stream) : launch_chain_t<true, true>(x, n_rows, a, act, stream);
    return bf16_parts ? launch_chain_t<false, false>(x, n_rows, a, act, stream) : launch_chain_t<true, false>(x, n_rows, a, act, stream);
}

extern "C" {

int64_t geossl_weight_image_bytes(void) { return tc::kWImage; }

int geossl_pack_weight_ld(const float* weight, int ldw, int transpose_weight, int bf16_parts, void* image, void* stream) {
    GEOSSL_REQUIRE(weight && image && ldw >= 128, "null pointer / leading dimension < 128");
    if (bf16_parts) {
        GEOSSL_CUDA(launch_pdl(tc::pack_weight_kernel<false>, dim3(8), dim3(256), 0, as_stream(stream), weight, ldw, transpose_weight, (uint8_t*)image));
    } else {
        GEOSSL_CUDA(launch_pdl(tc::pack_weight_kernel<true>, dim3(8), dim3(256), 0, as_stream(stream), weight, ldw, transpose_weight, (uint8_t*)image));
    }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_pack_weight_pair(const float* weight, int ldw, void* images, void* stream) {
    GEOSSL_REQUIRE(weight && images && ldw >= 128, "null pointer / leading dimension < 128");
    GEOSSL_CUDA(launch_pdl(tc::pack_weight_pair_kernel, dim3(8, 2), dim3(256), 0, as_stream(stream), weight, ldw, (uint8_t*)images));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_pack_weight(const float* weight, int transpose_weight, int bf16_parts, void* image, void* stream) {
    return geossl_pack_weight_ld(weight, 128, transpose_weight, bf16_parts, image, stream);
}

int geossl_pack_weights_batched(const float* const* weights, const int32_t* lds, int n_weights, void* images, void* stream) {
    if (n_weights == 0) return 0;
    GEOSSL_REQUIRE(weights && images && n_weights > 0, "null pointer");
    tc::pack_weights_batched_kernel<<<dim3(8, n_weights, 2), 256, 0, as_stream(stream)>>>(weights, lds, (uint8_t*)images);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_debug_set_trace_linear(long long* device_buffer) {
    GEOSSL_CUDA(cudaMemcpyToSymbol(tc::g_trace_lin, &device_buffer, sizeof(device_buffer)));
    return 0;
}

static int linear_tc_dispatch(const float* x, int64_t n_rows, const uint8_t* weight, const float* bias, int pre_act,
                              const float* act_grad_input, const float* residual, float* y, int bf16_parts,
                              int64_t ldx, int64_t ldz, int64_t ldr, int64_t ldy, int act, int k_cols, cudaStream_t st) {
    const int key = (bf16_parts ? 0 : 8) | (pre_act ? 4 : 0) | (act_grad_input ? 2 : 0) | (residual ? 1 : 0);
#define GEOSSL_LIN_CASE(K, A, B, C, D) case K: return launch_linear_tc<A, B, C, D>(x, n_rows, weight, bias, act_grad_input, residual, y, ldx, ldz, ldr, ldy, act, k_cols, st);
    switch (key) {
        GEOSSL_LIN_CASE(0, false, false, false, false) GEOSSL_LIN_CASE(1, false, false, false, true)
        GEOSSL_LIN_CASE(2, false, false, true, false)  GEOSSL_LIN_CASE(3, false, false, true, true)
        GEOSSL_LIN_CASE(4, false, true, false, false)  GEOSSL_LIN_CASE(5, false, true, false, true)
        GEOSSL_LIN_CASE(6, false, true, true, false)   GEOSSL_LIN_CASE(7, false, true, true, true)
        GEOSSL_LIN_CASE(8, true, false, false, false)  GEOSSL_LIN_CASE(9, true, false, false, true)
        GEOSSL_LIN_CASE(10, true, false, true, false)  GEOSSL_LIN_CASE(11, true, false, true, true)
        GEOSSL_LIN_CASE(12, true, true, false, false)  GEOSSL_LIN_CASE(13, true, true, false, true)
        GEOSSL_LIN_CASE(14, true, true, true, false)   GEOSSL_LIN_CASE(15, true, true, true, true)
    }
#undef GEOSSL_LIN_CASE
    return GEOSSL_EINVAL;
}

int geossl_linear_tc(const float* x, int64_t n_rows, const void* weight_image, const float* bias, int pre_ssp,
                     const float* act_grad_input, const float* residual, float* y, int bf16_parts, void* stream) {
    if (n_rows == 0) return 0;
    GEOSSL_REQUIRE(x && weight_image && y && n_rows > 0, "null pointer");
    return linear_tc_dispatch(x, n_rows, (const uint8_t*)weight_image, bias, pre_ssp, act_grad_input, residual, y, bf16_parts,
                              128, 128, 128, 128, tc::kActSsp, 128, as_stream(stream));
}

int geossl_linear_tc_block(const float* x, int64_t ldx, int64_t n_rows, const void* weight_image, const float* bias, int act,
                           int pre_act, const float* act_grad_input, int64_t ldz, const float* residual, int64_t ldr,
                           float* y, int64_t ldy, int bf16_parts, int k_cols, void* stream) {
    if (n_rows == 0) return 0;
    GEOSSL_REQUIRE(x && weight_image && y && n_rows > 0, "null pointer");
    GEOSSL_REQUIRE(k_cols >= 16 && k_cols <= 128 && k_cols % 16 == 0, "k_cols must be a multiple of 16 in [16,128]");
    GEOSSL_REQUIRE(ldx >= k_cols && ldy >= 128 && ldx % 4 == 0 && (!act_grad_input || ldz >= 128) && (!residual || ldr >= 128), "bad leading dimension");
    GEOSSL_REQUIRE(act == tc::kActSsp || act == tc::kActSilu || !(pre_act || act_grad_input), "act must be 1 (ssp) or 2 (silu)");
    return linear_tc_dispatch(x, n_rows, (const uint8_t*)weight_image, bias, pre_act, act_grad_input, residual, y, bf16_parts,
                              ldx, ldz, ldr, ldy, act, k_cols, as_stream(stream));
}

int geossl_linear_chain_tc(const float* x, int64_t n_rows, const geossl_chain_stage* stages, int n_stages, int bf16_parts, int act,
                           void* stream) {
    if (n_rows == 0) return 0;
    GEOSSL_REQUIRE(x && stages && n_rows > 0, "null pointer");
    GEOSSL_REQUIRE(n_stages >= 1 && n_stages <= 4, "1..4 stages");
    GEOSSL_REQUIRE(act == tc::kActSsp || act == tc::kActSilu, "act must be 1 (ssp) or 2 (silu)");
    tc::ChainArgs a = {};
    a.n = n_stages;
    for (int i = 0; i < n_stages; ++i) {
        GEOSSL_REQUIRE(stages[i].weight_image != nullptr, "stage without a weight image");
        GEOSSL_REQUIRE(stages[i].act_next >= 0 && stages[i].act_next <= 2, "act_next must be 0, 1 or 2");
        a.st[i].w_image = (const uint8_t*)stages[i].weight_image;
        a.st[i].bias = stages[i].bias;
        a.st[i].z = stages[i].act_grad_input;
        a.st[i].residual = stages[i].residual;
        a.st[i].store = stages[i].store;
        a.st[i].act_next = stages[i].act_next;
        a.st[i].ldz = a.st[i].ldr = a.st[i].lds = 128;
    }
    GEOSSL_REQUIRE(stages[n_stages - 1].store != nullptr, "the last stage must store its result");
    return launch_chain(x, n_rows, a, bf16_parts, act, false, stream);
}

int geossl_linear_chain_ex(int64_t n_rows, const geossl_chain_stage_ex* stages, int n_stages, int bf16_parts, int act, void* stream) {
    if (n_rows == 0) return 0;
    GEOSSL_REQUIRE(stages && n_rows > 0, "null pointer");
    GEOSSL_REQUIRE(n_stages >= 1 && n_stages <= tc::kMaxChain, "1..6 stages");
    GEOSSL_REQUIRE(act == tc::kActSsp || act == tc::kActSilu, "act must be 1 (ssp) or 2 (silu)");
    tc::ChainArgs a = {};
    a.n = n_stages;
    for (int i = 0; i < n_stages; ++i) {
        const geossl_chain_stage_ex& S = stages[i];
        GEOSSL_REQUIRE(S.weight_image != nullptr, "stage without a weight image");
        GEOSSL_REQUIRE(S.act_next >= 0 && S.act_next <= 2 && S.x_act >= 0 && S.x_act <= 2, "activations must be 0, 1 or 2");
        GEOSSL_REQUIRE(i > 0 || S.x != nullptr, "the first stage needs an operand");
        GEOSSL_REQUIRE(!(S.x && S.keep), "a stage either loads its operand or keeps the previous one");
        GEOSSL_REQUIRE(i > 0 || !(S.keep || S.accumulate), "the first stage cannot keep / accumulate");
        GEOSSL_REQUIRE(!S.accumulate || stages[i - 1].partial, "an accumulating stage must follow a partial stage");
        GEOSSL_REQUIRE(!S.partial || (i + 1 < n_stages && stages[i + 1].accumulate), "a partial stage must be followed by an accumulating one");
        GEOSSL_REQUIRE((!S.x || (S.ldx >= 128 && S.ldx % 4 == 0)) && (!S.store || S.ld_store >= 128) && (!S.act_grad_input || S.ldz >= 128)
                       && (!S.residual || S.ldr >= 128), "bad leading dimension");
        a.st[i].w_image = (const uint8_t*)S.weight_image;
        a.st[i].bias = S.bias;
        a.st[i].z = S.act_grad_input;
        a.st[i].residual = S.residual;
        a.st[i].store = S.store;
        a.st[i].x = S.x;
        a.st[i].ldz = S.ldz; a.st[i].ldr = S.ldr; a.st[i].lds = S.ld_store; a.st[i].ldx = S.ldx;
        a.st[i].act_next = S.act_next; a.st[i].x_act = S.x_act;
        a.st[i].keep = S.keep; a.st[i].accumulate = S.accumulate; a.st[i].partial = S.partial;
    }
    GEOSSL_REQUIRE(stages[n_stages - 1].store != nullptr, "the last stage must store its result");
    return launch_chain(stages[0].x, n_rows, a, bf16_parts, act, true, stream);
}

int64_t geossl_linear_wgrad_tc_workspace(int64_t n_rows) { return (int64_t)tc::wgrad_grid(n_rows) * tc::kWgPart; }

static int wgrad_block(const float* grad_y, int64_t ld_dy, const float* x, int64_t ld_x, int64_t n_rows, int pre_act, float* workspace,
                       float* grad_weight, int ld_gw, float* grad_bias, int x_cols, void* stream);

int geossl_linear_wgrad_tc(const float* grad_y, const float* x, int64_t n_rows, int pre_ssp, float* workspace,
                           float* grad_weight, float* grad_bias, void* stream) {
    return wgrad_block(grad_y, 128, x, 128, n_rows, pre_ssp ? tc::kActSsp : 0, workspace, grad_weight, 128, grad_bias, 128, stream);
}

int geossl_linear_wgrad_tc_block(const float* grad_y, int64_t ld_dy, const float* x, int64_t ld_x, int64_t n_rows, int pre_act,
                                 float* workspace, float* grad_weight, int ld_gw, float* grad_bias, int x_cols, void* stream) {
    GEOSSL_REQUIRE(x_cols >= 8 && x_cols <= 128 && x_cols % 8 == 0, "x_cols must be a multiple of 8 in [8,128]");
    GEOSSL_REQUIRE(ld_dy >= 128 && ld_x >= x_cols && ld_gw >= 128 && ld_dy % 4 == 0 && ld_x % 4 == 0, "bad leading dimension");
    GEOSSL_REQUIRE(pre_act >= 0 && pre_act <= 2, "pre_act must be 0, 1 (ssp) or 2 (silu)");
    return wgrad_block(grad_y, ld_dy, x, ld_x, n_rows, pre_act, workspace, grad_weight, ld_gw, grad_bias, x_cols, stream);
}

static int wgrad_launch(const tc::WgBatch& b, int n_problems, int64_t n_rows, void* stream) {
    const size_t smem = tc::WgLayout::kBytes + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::linear_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    // the launch shares the SMs with the main backward chain: the CTA budget of ONE weight-gradient launch is split among
    // the problems (a 3 x 120-CTA burst measurably slows the chain: 2.33 vs 2.22 ms per SchNet step)
    int grid = (tc::wgrad_grid(n_rows) + n_problems - 1) / n_problems;
    if (grid < 1) grid = 1;
    GEOSSL_CUDA(launch_pdl(tc::linear_wgrad_tc_kernel<false>, dim3(grid, n_problems), dim3(256), smem, as_stream(stream), b, n_rows));
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(tc::linear_wgrad_reduce_kernel, dim3((tc::kWgPart + 31) / 32, n_problems), dim3(256), 0, as_stream(stream), b, grid));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

static int wgrad_block(const float* grad_y, int64_t ld_dy, const float* x, int64_t ld_x, int64_t n_rows, int pre_act, float* workspace,
                       float* grad_weight, int ld_gw, float* grad_bias, int x_cols, void* stream) {
    GEOSSL_REQUIRE(grad_y && x && workspace && grad_weight && n_rows > 0, "null pointer or empty input");
    tc::WgBatch b = {};
    b.p[0] = tc::WgProblem{grad_y, x, workspace, grad_weight, grad_bias, ld_dy, ld_x, pre_act, x_cols, ld_gw};
    return wgrad_launch(b, 1, n_rows, stream);
}

int geossl_linear_wgrad_tc_batch(const geossl_wgrad_problem* problems, int n_problems, int64_t n_rows, void* stream) {
    if (n_problems == 0 || n_rows == 0) return 0;
    GEOSSL_REQUIRE(problems && n_problems >= 1 && n_problems <= tc::kMaxWgBatch && n_rows > 0, "1..6 problems over a non-empty row range");
    tc::WgBatch b = {};
    for (int i = 0; i < n_problems; ++i) {
        const geossl_wgrad_problem& P = problems[i];
        GEOSSL_REQUIRE(P.grad_y && P.x && P.workspace && P.grad_weight, "null pointer");
        GEOSSL_REQUIRE(P.x_cols >= 8 && P.x_cols <= 128 && P.x_cols % 8 == 0, "x_cols must be a multiple of 8 in [8,128]");
        GEOSSL_REQUIRE(P.ld_dy >= 128 && P.ld_x >= P.x_cols && P.ld_gw >= 128 && P.ld_dy % 4 == 0 && P.ld_x % 4 == 0, "bad leading dimension");
        GEOSSL_REQUIRE(P.pre_act >= 0 && P.pre_act <= 2, "pre_act must be 0, 1 (ssp) or 2 (silu)");
        b.p[i] = tc::WgProblem{P.grad_y, P.x, P.workspace, P.grad_weight, P.grad_bias, P.ld_dy, P.ld_x, P.pre_act, P.x_cols, P.ld_gw};
    }
    return wgrad_launch(b, n_problems, n_rows, stream);
}

}  // extern "C"
