// DDM (denoising distance matching) head, fp32 SIMT edition.
//
// Replaces the distance block of do_DDM (examples/pretrain_GeoSSL.py:197-205) and
// NCSN_version_03.forward (examples/NCSN.py:183-212) with its two MultiLayerPerceptrons (NCSN.py:9-43):
// sigma-indexed distance perturbation, 1->H->1 distance embedding, gather h[u]+h[v], score MLP
// (H+1)->H->H/2->1, sigma^alpha weighted squared error, mean over graphs -- one kernel forward, one
// backward (which recomputes the forward per 64-pair tile, so nothing per-pair is saved in HBM).
//
// Persistent CTAs (one per SM) walk 64-pair tiles; the score MLP runs as register-tiled SIMT GEMMs on
// k-major shared-memory tiles (simt_tile.cuh).  Weight matrices sit in shared memory once, row-major with
// an odd row stride (H+1), which serves both the forward (B[k][n] = W[n][k]) and the data-gradient
// (B[k][n] = W[k][n]) orientation without bank conflicts.  Weight gradients accumulate in registers across
// the CTA's tiles and are reduced over CTAs in a fixed order (deterministic); only dL/dh uses fp32
// atomics (RED), because pairs scatter to both endpoints.
#include "common.cuh"
#include "simt_tile.cuh"

namespace geossl {

template <int H>
struct HeadCfg {
    static constexpr int HH = H / 2;
    static constexpr int LD = H + 1;
    static constexpr int TP = 64;
    static constexpr int S = TP + 4;
    using C1 = TileCfg<H>;      // (pairs x H) outputs
    using C2 = TileCfg<HH>;     // (pairs x H/2) outputs
    static constexpr int M0 = (H / C1::TY) > 0 ? (H / C1::TY) : 1;     // dW0 rows per thread
    static constexpr int M1 = (HH / C1::TY) > 0 ? (HH / C1::TY) : 1;   // dW1 rows per thread
    // shared memory (floats)
    static constexpr int kW0 = 0;                        // [H][LD]    output_mlp.layers.0.weight
    static constexpr int kW1 = kW0 + H * LD;             // [HH][LD]   output_mlp.layers.1.weight (padded rows)
    static constexpr int kFeat = kW1 + HH * LD + 3;      // [H+1][S]   (offset keeps 16 B alignment below)
    static constexpr int kFeatA = (kFeat + 3) / 4 * 4;
    static constexpr int kZ1 = kFeatA + (H + 1) * S;     // [H][S]     z1, later dz1
    static constexpr int kZ2 = kZ1 + H * S;              // [HH][S]    dz2 (backward only)
    static constexpr int kVec = kZ2 + HH * S;            // small vectors
    static constexpr int oB0 = kVec;                     // [H]
    static constexpr int oB1 = oB0 + H;                  // [HH]
    static constexpr int oW2 = oB1 + HH;                 // [HH]
    static constexpr int oIW0 = oW2 + HH;                // [H] input mlp layer 0 weight
    static constexpr int oIB0 = oIW0 + H;                // [H]
    static constexpr int oIW1 = oIB0 + H;                // [H] input mlp layer 1 weight
    static constexpr int oU = oIW1 + H;                  // int [TP]
    static constexpr int oV = oU + TP;                   // int [TP]
    static constexpr int oSig = oV + TP;                 // sigma
    static constexpr int oDt = oSig + TP;                // perturbed distance
    static constexpr int oEmb = oDt + TP;
    static constexpr int oTgt = oEmb + TP;               // target
    static constexpr int oSa = oTgt + TP;                // sigma^alpha (0 for invalid pairs)
    static constexpr int oDemb = oSa + TP;
    static constexpr int oRed = oDemb + TP;              // [64] reduction scratch
    static constexpr int kFloats = oRed + 64;
    // per-CTA partial gradient layout
    static constexpr int pW0 = 0;                        // [H][H+1]
    static constexpr int pB0 = pW0 + H * LD;
    static constexpr int pW1 = pB0 + H;                  // [HH][H]
    static constexpr int pB1 = pW1 + HH * H;
    static constexpr int pW2 = pB1 + HH;                 // [HH]
    static constexpr int pB2 = pW2 + HH;                 // [1]
    static constexpr int pIW0 = pB2 + 1;                 // [H]
    static constexpr int pIB0 = pIW0 + H;
    static constexpr int pIW1 = pIB0 + H;
    static constexpr int pIB1 = pIW1 + H;                // [1]
    static constexpr int kPartial = pIB1 + 1;
};

struct HeadIn {
    const float* h; const int64_t* sei; const int64_t* batch; int64_t n_pairs;
    const float* dist; const float* noise; const int64_t* noise_level; const float* sigmas; int n_levels;
    float anneal_power;
    geossl_ddm_params p;
    // Capacity-padded batches (CUDA-graph replay over variable-size batches): n_pairs is the CAPACITY (row stride of
    // sei, length of dist / noise); the number of live pairs is read from device memory.  NULL => all n_pairs are live.
    const int32_t* n_live;
    __device__ __forceinline__ int64_t live() const {
        if (n_live == nullptr) return n_pairs;
        const int64_t l = (int64_t)__ldg(n_live);
        return l < 0 ? 0 : (l < n_pairs ? l : n_pairs);
    }
};

template <int H>
__device__ __forceinline__ void head_load_weights(const HeadIn& in, float* smem) {
    using K = HeadCfg<H>;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < H * K::LD; idx += 256) smem[K::kW0 + idx] = __ldg(in.p.out_w0 + idx);
    for (int idx = tid; idx < K::HH * H; idx += 256) smem[K::kW1 + (idx / H) * K::LD + idx % H] = __ldg(in.p.out_w1 + idx);
    if (tid < H) {
        smem[K::oB0 + tid] = __ldg(in.p.out_b0 + tid);
        smem[K::oIW0 + tid] = __ldg(in.p.in_w0 + tid);
        smem[K::oIB0 + tid] = __ldg(in.p.in_b0 + tid);
        smem[K::oIW1 + tid] = __ldg(in.p.in_w1 + tid);
    }
    if (tid < K::HH) {
        smem[K::oB1 + tid] = __ldg(in.p.out_b1 + tid);
        smem[K::oW2 + tid] = __ldg(in.p.out_w2 + tid);
    }
}

// Per-pair scalars, distance embedding and the k-major feature tile.  Returns the largest graph id seen
// by this thread (-1 if none).  Ends with a __syncthreads().
template <int H>
__device__ __forceinline__ int head_build_tile(const HeadIn& in, int64_t p0, float* smem) {
    using K = HeadCfg<H>;
    using C1 = typename K::C1;
    const int tid = threadIdx.x;
    int* sU = reinterpret_cast<int*>(smem + K::oU);
    int* sV = reinterpret_cast<int*>(smem + K::oV);
    int gmax = -1;
    if (tid < K::TP) {
        const int64_t p = p0 + tid;
        int u = 0, v = 0;
        float sigma = 1.f, dt = 0.f, tgt = 0.f, sa = 0.f;
        if (p < in.live()) {
            u = (int)in.sei[p];
            v = (int)in.sei[in.n_pairs + p];
            const int g = (int)in.batch[u];
            gmax = g;
            int lvl = (int)in.noise_level[g];
            lvl = lvl < 0 ? 0 : (lvl >= in.n_levels ? in.n_levels - 1 : lvl);
            sigma = __ldg(in.sigmas + lvl);
            const float d = __ldg(in.dist + p);
            dt = __fadd_rn(d, __fmul_rn(__ldg(in.noise + p), sigma));           // NCSN.py:196
            tgt = __fmul_rn(-(1.f / __fmul_rn(sigma, sigma)), __fsub_rn(dt, d)); // NCSN.py:199 (op order kept)
            sa = (in.anneal_power == 2.f) ? sigma * sigma : powf(sigma, in.anneal_power);
        }
        sU[tid] = u; sV[tid] = v;
        smem[K::oSig + tid] = sigma; smem[K::oDt + tid] = dt; smem[K::oTgt + tid] = tgt; smem[K::oSa + tid] = sa;
    }
    __syncthreads();
    {   // distance embedding: 4 threads per pair
        const int pl = tid >> 2, q = tid & 3;
        const float dt = smem[K::oDt + pl];
        float s = 0.f;
        for (int k = q; k < H; k += 4) {
            const float pre = fmaf(smem[K::oIW0 + k], dt, smem[K::oIB0 + k]);
            s = fmaf(smem[K::oIW1 + k], fmaxf(pre, 0.f), s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (q == 0) {
            const float emb = s + __ldg(in.p.in_b1);
            smem[K::oEmb + pl] = emb;
            smem[K::kFeatA + H * K::S + pl] = emb;
        }
    }
    {   // feat[c][p] = h[u][c] + h[v][c]
        const int tx = tid % C1::TX, ty = tid / C1::TX;
#pragma unroll
        for (int i = 0; i < C1::ME; ++i) {
            const int pl = ty * C1::ME + i;
            const float* hu = in.h + (int64_t)sU[pl] * H;
            const float* hv = in.h + (int64_t)sV[pl] * H;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = tx + C1::TX * j;
                smem[K::kFeatA + c * K::S + pl] = __ldg(hu + c) + __ldg(hv + c);
            }
        }
    }
    __syncthreads();
    return gmax;
}

// z1 = relu(feat W0^T + b0) -> sZ1 (k-major).  Ends with a __syncthreads().
template <int H>
__device__ __forceinline__ void head_layer0(float* smem) {
    using K = HeadCfg<H>;
    using C1 = typename K::C1;
    const int tid = threadIdx.x, tx = tid % C1::TX, r0 = (tid / C1::TX) * C1::ME;
    float acc[C1::ME][4];
    zero_acc(acc);
    gemm_rows<H>(acc, smem + K::kFeatA, smem + K::kW0, 1, K::LD, H + 1, tx, r0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = tx + C1::TX * j;
        const float b = smem[K::oB0 + n];
#pragma unroll
        for (int i = 0; i < C1::ME; ++i) smem[K::kZ1 + n * K::S + r0 + i] = fmaxf(acc[i][j] + b, 0.f);
    }
    __syncthreads();
}

// z2 = relu(z1 W1^T + b1) in registers (pairs r0.., columns tx2 + TX2*j); raw score reduced over the row.
template <int H>
__device__ __forceinline__ void head_layer1(float* smem, float (&z2)[HeadCfg<H>::C2::ME][4],
                                            float (&score)[HeadCfg<H>::C2::ME], float out_b2) {
    using K = HeadCfg<H>;
    using C2 = typename K::C2;
    const int tid = threadIdx.x, tx = tid % C2::TX, r0 = (tid / C2::TX) * C2::ME;
    zero_acc(z2);
    gemm_rows<K::HH>(z2, smem + K::kZ1, smem + K::kW1, 1, K::LD, H, tx, r0);
#pragma unroll
    for (int i = 0; i < C2::ME; ++i) score[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = tx + C2::TX * j;
        const float b = smem[K::oB1 + n], w = smem[K::oW2 + n];
#pragma unroll
        for (int i = 0; i < C2::ME; ++i) {
            z2[i][j] = fmaxf(z2[i][j] + b, 0.f);
            score[i] = fmaf(z2[i][j], w, score[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < C2::ME; ++i) {
#pragma unroll
        for (int o = C2::TX / 2; o > 0; o >>= 1) score[i] += __shfl_xor_sync(0xffffffffu, score[i], o);
        score[i] += out_b2;
    }
}

template <int H>
__global__ void __launch_bounds__(256, 1) ddm_head_fwd_kernel(HeadIn in, float* __restrict__ workspace) {
    using K = HeadCfg<H>;
    using C2 = typename K::C2;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    head_load_weights<H>(in, smem);
    __syncthreads();
    const float out_b2 = __ldg(in.p.out_b2);
    float loss_acc = 0.f;
    int gmax = -1;
    const int64_t n_tiles = (in.live() + K::TP - 1) / K::TP;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        gmax = max(gmax, head_build_tile<H>(in, t * K::TP, smem));
        head_layer0<H>(smem);
        float z2[C2::ME][4], score[C2::ME];
        head_layer1<H>(smem, z2, score, out_b2);
        if (tid % C2::TX == 0) {
            const int r0 = (tid / C2::TX) * C2::ME;
#pragma unroll
            for (int i = 0; i < C2::ME; ++i) {
                const int pl = r0 + i;
                const float s = score[i] * (1.f / smem[K::oSig + pl]);          // NCSN.py:205
                const float diff = s - smem[K::oTgt + pl];
                loss_acc += 0.5f * (diff * diff) * smem[K::oSa + pl];           // NCSN.py:209 (sa = 0 if invalid)
            }
        }
        __syncthreads();
    }
    // block reduction (fixed order) of the loss partial and the largest graph id
    __shared__ float red_l[8];
    __shared__ int red_g[8];
    loss_acc = warp_sum(loss_acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
    if ((tid & 31) == 0) { red_l[tid >> 5] = loss_acc; red_g[tid >> 5] = gmax; }
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        int g = -1;
        for (int w = 0; w < 8; ++w) { s += red_l[w]; g = max(g, red_g[w]); }
        workspace[2 * blockIdx.x] = s;
        workspace[2 * blockIdx.x + 1] = (float)g;
    }
}

__global__ void ddm_loss_finalize_kernel(const float* __restrict__ workspace, int n_parts, float* __restrict__ loss) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float s = 0.f, g = -1.f;
        for (int p = 0; p < n_parts; ++p) { s += workspace[2 * p]; g = fmaxf(g, workspace[2 * p + 1]); }
        const float ng = g + 1.f;
        loss[0] = ng > 0.f ? s / ng : 0.f;
        loss[1] = ng;
    }
}

template <int H>
__global__ void __launch_bounds__(256, 1)
ddm_head_bwd_kernel(HeadIn in, const float* __restrict__ loss_aux, const float* __restrict__ grad_loss,
                    float* __restrict__ grad_h, float* __restrict__ workspace) {
    using K = HeadCfg<H>;
    using C1 = typename K::C1;
    using C2 = typename K::C2;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int tx1 = tid % C1::TX, ty1 = tid / C1::TX, r01 = ty1 * C1::ME;
    const int tx2 = tid % C2::TX, ty2 = tid / C2::TX, r02 = ty2 * C2::ME;
    int* sU = reinterpret_cast<int*>(smem + K::oU);
    int* sV = reinterpret_cast<int*>(smem + K::oV);
    head_load_weights<H>(in, smem);
    __syncthreads();
    const float out_b2 = __ldg(in.p.out_b2);
    const float ng = __ldg(loss_aux + 1);
    const float gscale = ng > 0.f ? __ldg(grad_loss) / ng : 0.f;

    float aW0[K::M0][4], aW1[K::M1][4], aW2[4];
    zero_acc(aW0); zero_acc(aW1);
#pragma unroll
    for (int j = 0; j < 4; ++j) aW2[j] = 0.f;
    float aB2 = 0.f, aW0last = 0.f, aB0 = 0.f, aB1 = 0.f, aIW0 = 0.f, aIB0 = 0.f, aIW1 = 0.f, aIB1 = 0.f;

    const int64_t n_tiles = (in.live() + K::TP - 1) / K::TP;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        head_build_tile<H>(in, t * K::TP, smem);
        head_layer0<H>(smem);
        {
            float z2[C2::ME][4], score[C2::ME];
            head_layer1<H>(smem, z2, score, out_b2);
#pragma unroll
            for (int i = 0; i < C2::ME; ++i) {
                const int pl = r02 + i;
                const float inv = 1.f / smem[K::oSig + pl];
                const float s = score[i] * inv;
                const float dsr = (s - smem[K::oTgt + pl]) * smem[K::oSa + pl] * gscale * inv;   // d loss / d raw score
                if (tx2 == 0) aB2 += dsr;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = tx2 + C2::TX * j;
                    aW2[j] = fmaf(dsr, z2[i][j], aW2[j]);
                    smem[K::kZ2 + n * K::S + pl] = z2[i][j] > 0.f ? dsr * smem[K::oW2 + n] : 0.f;
                }
            }
        }
        __syncthreads();
        // dW1[k][i] += sum_p dz2[p][k] z1[p][i]; db1[k] += sum_p dz2[p][k]
        gemm_wgrad<H, K::M1>(aW1, smem + K::kZ2, ty1 * K::M1, K::HH, smem + K::kZ1, tx1);
        if (tid < K::HH) {
            float s = 0.f;
            for (int p = 0; p < K::TP; ++p) s += smem[K::kZ2 + tid * K::S + p];
            aB1 += s;
        }
        // dz1 = (dz2 W1) * [z1 > 0]
        float acc[C1::ME][4];
        zero_acc(acc);
        gemm_rows<H>(acc, smem + K::kZ2, smem + K::kW1, K::LD, 1, K::HH, tx1, r01);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < C1::ME; ++i)
                if (!(smem[K::kZ1 + (tx1 + C1::TX * j) * K::S + r01 + i] > 0.f)) acc[i][j] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < C1::ME; ++i) smem[K::kZ1 + (tx1 + C1::TX * j) * K::S + r01 + i] = acc[i][j];
        __syncthreads();
        // dW0[k][i] += sum_p dz1[p][k] feat[p][i] (i < H), last column and db0 per thread k
        gemm_wgrad<H, K::M0>(aW0, smem + K::kZ1, ty1 * K::M0, H, smem + K::kFeatA, tx1);
        if (tid < H) {
            float s = 0.f, sl = 0.f;
            for (int p = 0; p < K::TP; ++p) {
                const float v = smem[K::kZ1 + tid * K::S + p];
                s += v;
                sl = fmaf(v, smem[K::oEmb + p], sl);
            }
            aB0 += s;
            aW0last += sl;
        }
        // dfeat[:, :H] = dz1 W0[:, :H]  -> scatter to both endpoints
        zero_acc(acc);
        gemm_rows<H>(acc, smem + K::kZ1, smem + K::kW0, K::LD, 1, H, tx1, r01);
#pragma unroll
        for (int i = 0; i < C1::ME; ++i) {
            const int pl = r01 + i;
            if (t * K::TP + pl < in.live()) {
                float* gu = grad_h + (int64_t)sU[pl] * H;
                float* gv = grad_h + (int64_t)sV[pl] * H;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = tx1 + C1::TX * j;
                    atomicAdd(gu + c, acc[i][j]);
                    atomicAdd(gv + c, acc[i][j]);
                }
            }
        }
        {   // demb_p = sum_k dz1[p][k] W0[k][H]
            const int pl = tid >> 2, q = tid & 3;
            float s = 0.f;
            for (int k = q; k < H; k += 4) s = fmaf(smem[K::kZ1 + k * K::S + pl], smem[K::kW0 + k * K::LD + H], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0) smem[K::oDemb + pl] = s;
        }
        __syncthreads();
        if (tid < H) {   // distance-embedding MLP gradients, thread per hidden unit
            const float w0 = smem[K::oIW0 + tid], b0 = smem[K::oIB0 + tid], w1 = smem[K::oIW1 + tid];
            for (int p = 0; p < K::TP; ++p) {
                const float dt = smem[K::oDt + p], de = smem[K::oDemb + p];
                const float pre = fmaf(w0, dt, b0);
                if (pre > 0.f) {
                    aIW1 = fmaf(de, pre, aIW1);
                    const float dpre = de * w1;
                    aIW0 = fmaf(dpre, dt, aIW0);
                    aIB0 += dpre;
                }
                if (tid == 0) aIB1 += de;
            }
        }
        __syncthreads();
    }

    // ---- per-CTA partials
    float* ws = workspace + (int64_t)blockIdx.x * K::kPartial;
    if (ty1 * K::M0 < H) {
#pragma unroll
        for (int i = 0; i < K::M0; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ws[K::pW0 + (ty1 * K::M0 + i) * K::LD + tx1 + C1::TX * j] = aW0[i][j];
    }
    if (ty1 * K::M1 < K::HH) {
#pragma unroll
        for (int i = 0; i < K::M1; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) ws[K::pW1 + (ty1 * K::M1 + i) * H + tx1 + C1::TX * j] = aW1[i][j];
    }
    if (tid < H) {
        ws[K::pW0 + tid * K::LD + H] = aW0last;
        ws[K::pB0 + tid] = aB0;
        ws[K::pIW0 + tid] = aIW0;
        ws[K::pIB0 + tid] = aIB0;
        ws[K::pIW1 + tid] = aIW1;
    }
    if (tid < K::HH) ws[K::pB1 + tid] = aB1;
    if (tid == 0) ws[K::pIB1] = aIB1;
    // out_w2 / out_b2: reduce over the TY2 row groups through shared memory
    __syncthreads();
    float* red = smem + K::kFeatA;     // [TY2][HH] and [TY2]
#pragma unroll
    for (int j = 0; j < 4; ++j) red[ty2 * K::HH + tx2 + C2::TX * j] = aW2[j];
    if (tx2 == 0) red[C2::TY * K::HH + ty2] = aB2;
    __syncthreads();
    if (tid < K::HH) {
        float s = 0.f;
        for (int r = 0; r < C2::TY; ++r) s += red[r * K::HH + tid];
        ws[K::pW2 + tid] = s;
    }
    if (tid == 0) {
        float s = 0.f;
        for (int r = 0; r < C2::TY; ++r) s += red[C2::TY * K::HH + r];
        ws[K::pB2] = s;
    }
}

template <int H>
__global__ void ddm_head_reduce_kernel(const float* __restrict__ workspace, int n_parts, geossl_ddm_grads g) {
    using K = HeadCfg<H>;
    pdl_launch_dependents();
    pdl_wait();

    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K::kPartial) return;
    // four independent chains (fixed association order => still deterministic) keep enough loads in flight
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 3 < n_parts; p += 4) {
        s0 += workspace[(int64_t)p * K::kPartial + idx];
        s1 += workspace[(int64_t)(p + 1) * K::kPartial + idx];
        s2 += workspace[(int64_t)(p + 2) * K::kPartial + idx];
        s3 += workspace[(int64_t)(p + 3) * K::kPartial + idx];
    }
    for (; p < n_parts; ++p) s0 += workspace[(int64_t)p * K::kPartial + idx];
    const float s = (s0 + s1) + (s2 + s3);
    if (idx < K::pB0) g.out_w0[idx] = s;
    else if (idx < K::pW1) g.out_b0[idx - K::pB0] = s;
    else if (idx < K::pB1) g.out_w1[idx - K::pW1] = s;
    else if (idx < K::pW2) g.out_b1[idx - K::pB1] = s;
    else if (idx < K::pB2) g.out_w2[idx - K::pW2] = s;
    else if (idx < K::pIW0) g.out_b2[0] = s;
    else if (idx < K::pIB0) g.in_w0[idx - K::pIW0] = s;
    else if (idx < K::pIW1) g.in_b0[idx - K::pIB0] = s;
    else if (idx < K::pIB1) g.in_w1[idx - K::pIW1] = s;
    else g.in_b1[0] = s;
}

__global__ void pair_distance_kernel(const float* __restrict__ pos, const int64_t* __restrict__ sei, int64_t n_pairs,
                                     float* __restrict__ dist) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const int64_t u = sei[p], v = sei[n_pairs + p];
    const float dx = __fsub_rn(pos[3 * u], pos[3 * v]), dy = __fsub_rn(pos[3 * u + 1], pos[3 * v + 1]),
                dz = __fsub_rn(pos[3 * u + 2], pos[3 * v + 2]);
    // sqrt(sum((u - v) ** 2, dim=1))   (pretrain_GeoSSL.py:201)
    dist[p] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

static int head_grid(int64_t n_pairs) {
    int64_t tiles = (n_pairs + 63) / 64;
    return (int)(tiles < kNumSM ? (tiles > 0 ? tiles : 1) : kNumSM);
}

template <int H>
int launch_head_fwd(const HeadIn& in, float* workspace, float* loss, cudaStream_t st) {
    const size_t smem = HeadCfg<H>::kFloats * sizeof(float);
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(ddm_head_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    const int grid = head_grid(in.n_pairs);
    ddm_head_fwd_kernel<H><<<grid, 256, smem, st>>>(in, workspace);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    count_launch();
    ddm_loss_finalize_kernel<<<1, 32, 0, st>>>(workspace, grid, loss);
    return 0;
}

template <int H>
int launch_head_bwd(const HeadIn& in, int64_t n_atoms, const float* loss_aux, const float* grad_loss, float* workspace,
                    float* grad_h, const geossl_ddm_grads& g, cudaStream_t st) {
    const size_t smem = HeadCfg<H>::kFloats * sizeof(float);
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(ddm_head_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    cudaError_t e = cudaMemsetAsync(grad_h, 0, sizeof(float) * (size_t)n_atoms * H, st);
    if (e != cudaSuccess) return (int)e;
    const int grid = head_grid(in.n_pairs);
    ddm_head_bwd_kernel<H><<<grid, 256, smem, st>>>(in, loss_aux, grad_loss, grad_h, workspace);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    count_launch();
    const int n = HeadCfg<H>::kPartial;
    ddm_head_reduce_kernel<H><<<(n + 255) / 256, 256, 0, st>>>(workspace, grid, g);
    return 0;
}

static int64_t head_workspace(int H) {
    switch (H) {
        case 32: return (int64_t)kNumSM * HeadCfg<32>::kPartial;
        case 64: return (int64_t)kNumSM * HeadCfg<64>::kPartial;
        case 128: return (int64_t)kNumSM * HeadCfg<128>::kPartial;
        default: return -1;
    }
}

}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_pair_distance(const float* pos, const int64_t* sei, int64_t n_pairs, float* dist, void* stream) {
    if (n_pairs == 0) return 0;
    GEOSSL_REQUIRE(pos && sei && dist && n_pairs > 0, "null pointer");
    pair_distance_kernel<<<(int)((n_pairs + 255) / 256), 256, 0, as_stream(stream)>>>(pos, sei, n_pairs, dist);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int64_t geossl_ddm_workspace(int H) { return head_workspace(H); }

static int64_t head_pad(int64_t n_pairs) { return (n_pairs + 63) / 64 * 64; }
constexpr int kPrepRowsHost = 8;          // == tc::kPrepRows (rows of per-pair scalars in the workspace)
static int64_t head_loss_offset() { return (head_workspace(128) + 63) / 64 * 64; }            // fused mode: 2 floats per CTA
static int64_t head_prep_offset() { return head_loss_offset() + (2 * kNumSM + 63) / 64 * 64; }
// parameter partial sums | loss partials (fused mode) | per-pair scalars (8 rows of n_pairs rounded up to the tile size)
int64_t geossl_ddm_workspace_tc(int64_t n_pairs) { return head_prep_offset() + 8 * head_pad(n_pairs > 0 ? n_pairs : 0) + 64; }

static int fill_head_in(HeadIn& in, const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs,
                        const int32_t* n_pairs_live, const float* dist, const float* noise, const int64_t* noise_level, const float* sigmas,
                        int n_levels, float anneal_power, const geossl_ddm_params* params) {
    if (!(h && sei && batch && dist && noise && noise_level && sigmas && params) || n_levels < 1) return GEOSSL_EINVAL;
    const geossl_ddm_params& p = *params;
    if (!(p.in_w0 && p.in_b0 && p.in_w1 && p.in_b1 && p.out_w0 && p.out_b0 && p.out_w1 && p.out_b1 && p.out_w2 && p.out_b2))
        return GEOSSL_EINVAL;
    in.h = h; in.sei = sei; in.batch = batch; in.n_pairs = n_pairs; in.dist = dist; in.noise = noise;
    in.noise_level = noise_level; in.sigmas = sigmas; in.n_levels = n_levels; in.anneal_power = anneal_power; in.p = p;
    in.n_live = n_pairs_live;
    return 0;
}

int geossl_ddm_head_fwd(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                        const float* dist, const float* noise, const int64_t* noise_level,
                        const float* sigmas, int n_levels, float anneal_power, int H,
                        const geossl_ddm_params* params, float* workspace, float* loss, void* stream) {
    GEOSSL_REQUIRE(workspace && loss && n_pairs >= 0, "null workspace/loss");
    HeadIn in;
    if (n_pairs == 0) {
        GEOSSL_CUDA(cudaMemsetAsync(loss, 0, 2 * sizeof(float), as_stream(stream)));
        return 0;
    }
    GEOSSL_REQUIRE(fill_head_in(in, h, sei, batch, n_pairs, n_pairs_live, dist, noise, noise_level, sigmas, n_levels, anneal_power, params) == 0,
                   "null input pointer");
    int rc;
    switch (H) {
        case 32: rc = launch_head_fwd<32>(in, workspace, loss, as_stream(stream)); break;
        case 64: rc = launch_head_fwd<64>(in, workspace, loss, as_stream(stream)); break;
        case 128: rc = launch_head_fwd<128>(in, workspace, loss, as_stream(stream)); break;
        default: set_error("%s: unsupported width H=%d (32/64/128)", __func__, H); return GEOSSL_EINVAL;
    }
    if (rc) { set_error("%s: %s", __func__, cudaGetErrorString((cudaError_t)rc)); return rc; }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_ddm_head_bwd(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                        int64_t n_atoms,
                        const float* dist, const float* noise, const int64_t* noise_level,
                        const float* sigmas, int n_levels, float anneal_power, int H,
                        const geossl_ddm_params* params, const float* loss_aux, const float* grad_loss, float* workspace,
                        float* grad_h, const geossl_ddm_grads* grads, void* stream) {
    GEOSSL_REQUIRE(workspace && grad_h && grads && loss_aux && grad_loss && n_pairs >= 0 && n_atoms >= 0, "null pointer");
    const geossl_ddm_grads& g = *grads;
    GEOSSL_REQUIRE(g.in_w0 && g.in_b0 && g.in_w1 && g.in_b1 && g.out_w0 && g.out_b0 && g.out_w1 && g.out_b1 && g.out_w2 && g.out_b2,
                   "null gradient pointer");
    HeadIn in;
    GEOSSL_REQUIRE(n_pairs > 0, "n_pairs must be > 0 (caller handles the empty case)");
    GEOSSL_REQUIRE(fill_head_in(in, h, sei, batch, n_pairs, n_pairs_live, dist, noise, noise_level, sigmas, n_levels, anneal_power, params) == 0,
                   "null input pointer");
    int rc;
    switch (H) {
        case 32: rc = launch_head_bwd<32>(in, n_atoms, loss_aux, grad_loss, workspace, grad_h, g, as_stream(stream)); break;
        case 64: rc = launch_head_bwd<64>(in, n_atoms, loss_aux, grad_loss, workspace, grad_h, g, as_stream(stream)); break;
        case 128: rc = launch_head_bwd<128>(in, n_atoms, loss_aux, grad_loss, workspace, grad_h, g, as_stream(stream)); break;
        default: set_error("%s: unsupported width H=%d (32/64/128)", __func__, H); return GEOSSL_EINVAL;
    }
    if (rc) { set_error("%s: %s", __func__, cudaGetErrorString((cudaError_t)rc)); return rc; }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"

// =====================================================================================================
// Tensor-core edition (H = 128): same math, 64-pair tiles, everything transposed so that TMEM lanes are
// hidden units / features and columns are pairs (tc.cuh: split-precision operands, three MMAs per product).
//   MMA1  D1^T[n][p]  = W0[:, :128] . feat^T          z1 = relu(D1^T + emb_p * W0[n][128] + b0[n])
//   MMA2  D2^T[n2][p] = W1 . z1^T                      z2 = relu(D2^T + b1[n2]); score_p = sum_n2 z2 w2[n2] + b2
//   (backward, recomputing the above)
//   MMA3  D3^T[n][p]  = W1^T . dz2^T                   dz1 = D3^T * [z1 > 0]
//   MMA4  D4^T[k][p]  = W0[:, :128]^T . dz1^T          -> RED into grad_h[u_p], grad_h[v_p] (128 contiguous bytes / warp)
//   WG0   DW0[n][k]  += dz1^T feat                     WG1  DW1^T[n][n2] += z1^T dz2        (accumulated in TMEM)
// Weight images live once in shared memory and serve both orientations (K-major for MMA1/2, MN-major for MMA3/4).
// Synchronous per-tile pipeline (stage -> MMA -> epilogue, __syncthreads between): the head is ~8 % of the step's
// FLOPs, so simplicity wins over overlap here.
// =====================================================================================================
#include "tc.cuh"

namespace geossl {
namespace tc {

__device__ long long* g_trace_head = nullptr;  // optional clock64() trace of CTA 0 (geossl_debug_set_trace_head)
__device__ __forceinline__ void trace_h(int tile, int event) {
    long long* t = g_trace_head;
    if (t != nullptr && blockIdx.x == 0 && tile < 32 && threadIdx.x == 0) t[tile * 16 + event] = clock64();
}

// Per-pair scalars of the DDM objective, one thread per pair, computed ONCE per launch instead of per tile inside
// the MMA pipeline (where the index -> graph id -> noise level -> sigma chain was four dependent global loads):
//   prep[0][p] = u, prep[1][p] = v, prep[2][p] = graph id   (int32 bit patterns)
//   prep[3][p] = sigma, [4] = perturbed distance d~, [5] = target, [6] = sigma^anneal, [7] = distance embedding(d~)
// Row stride = n_pad (pairs rounded up to the tile size); rows past n_pairs are zero / sigma = 1.
constexpr int kPrepRows = 8;
__global__ void __launch_bounds__(256)
ddm_pair_prep_kernel(HeadIn in, int64_t n_pad, float* __restrict__ prep) {
    __shared__ float sw0[128], sb0[128], sw1[128];
    pdl_launch_dependents();
    pdl_wait();

    if (threadIdx.x < 128) {
        sw0[threadIdx.x] = __ldg(in.p.in_w0 + threadIdx.x);
        sb0[threadIdx.x] = __ldg(in.p.in_b0 + threadIdx.x);
        sw1[threadIdx.x] = __ldg(in.p.in_w1 + threadIdx.x);
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pad) return;
    int u = 0, v = 0, g = -1;
    float sigma = 1.f, dt = 0.f, tgt = 0.f, sa = 0.f, emb = 0.f;
    if (p < in.live()) {
        u = (int)in.sei[p];
        v = (int)in.sei[in.n_pairs + p];
        g = (int)in.batch[u];
        int lvl = (int)in.noise_level[g];
        lvl = lvl < 0 ? 0 : (lvl >= in.n_levels ? in.n_levels - 1 : lvl);
        sigma = __ldg(in.sigmas + lvl);
        const float d = __ldg(in.dist + p);
        dt = __fadd_rn(d, __fmul_rn(__ldg(in.noise + p), sigma));                    // NCSN.py:194-195
        tgt = __fmul_rn(-(1.f / __fmul_rn(sigma, sigma)), __fsub_rn(dt, d));         // :196
        sa = (in.anneal_power == 2.f) ? sigma * sigma : powf(sigma, in.anneal_power);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;                                // input_distance_mlp, 1 -> 128 -> 1
#pragma unroll 4
        for (int k = 0; k < 128; k += 4) {
            s0 = fmaf(sw1[k], fmaxf(fmaf(sw0[k], dt, sb0[k]), 0.f), s0);
            s1 = fmaf(sw1[k + 1], fmaxf(fmaf(sw0[k + 1], dt, sb0[k + 1]), 0.f), s1);
            s2 = fmaf(sw1[k + 2], fmaxf(fmaf(sw0[k + 2], dt, sb0[k + 2]), 0.f), s2);
            s3 = fmaf(sw1[k + 3], fmaxf(fmaf(sw0[k + 3], dt, sb0[k + 3]), 0.f), s3);
        }
        emb = ((s0 + s1) + (s2 + s3)) + __ldg(in.p.in_b1);
    }
    prep[0 * n_pad + p] = __int_as_float(u);
    prep[1 * n_pad + p] = __int_as_float(v);
    prep[2 * n_pad + p] = __int_as_float(g);
    prep[3 * n_pad + p] = sigma;
    prep[4 * n_pad + p] = dt;
    prep[5 * n_pad + p] = tgt;
    prep[6 * n_pad + p] = sa;
    prep[7 * n_pad + p] = emb;
}

// Sum over the 32 lanes of each of 16 per-lane values with 16 shuffles (instead of 16 x 5): after the four halving
// exchanges lane l holds the total of element e(l) = (l >> 1) & 15 ... returned; both lanes of a pair (l, l^1) agree.
__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float send = h16 ? v[j] : v[j + 8], keep = h16 ? v[j + 8] : v[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float send = h8 ? a[j] : a[j + 4], keep = h8 ? a[j + 4] : a[j];
        b[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float send = h4 ? b[j] : b[j + 2], keep = h4 ? b[j + 2] : b[j];
        c[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = h2 ? c[0] : c[1], keep = h2 ? c[1] : c[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;                                                   // element index = 8*bit4 + 4*bit3 + 2*bit2 + bit1 of the lane
}
__device__ __forceinline__ int reduce16_index(int lane) { return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1); }

constexpr int kHP = 64;                        // pairs per tile
constexpr int kHBlkW = 128 * 128;              // [128 rows x 64 k] weight block (bytes)
constexpr int kHBlkT = kHP * 128;              // [64 rows x 64 k] tile block (bytes)
constexpr int kHThreads = 512;

struct HeadSmem {
    static constexpr int W0 = 0;                               // hi (2 k-blocks) | lo                       64 KB
    static constexpr int W1 = W0 + 4 * kHBlkW;                 // rows n2 < 64: hi (2 x 8 KB) | lo (2 x 8 KB)  32 KB
    static constexpr int FEAT = W1 + 4 * kHBlkT;               // hi (2 blk) | lo (2 blk)                     32 KB
    static constexpr int Z1 = FEAT + 4 * kHBlkT;               // 32 KB
    static constexpr int DZ1 = Z1 + 4 * kHBlkT;                // 32 KB
    static constexpr int DZ2 = DZ1 + 4 * kHBlkT;               // hi (1 blk) | lo (1 blk)                     16 KB
    static constexpr int SC = DZ2 + 2 * kHBlkT;                // per-pair scalars: 10 arrays x 64 x 4 B
    static constexpr int RED = SC + 10 * 64 * 4;               // [4][64] floats
    static constexpr int BAR = RED + 4 * 64 * 4;
    static constexpr int TMEM_PTR = BAR + 16;
    static constexpr int kBytes = TMEM_PTR + 16;
};
static_assert(HeadSmem::kBytes + 1024 <= 227 * 1024, "shared memory budget");

template <bool FP16, bool BWD>
__global__ void __launch_bounds__(kHThreads, 1)
ddm_head_tc_kernel(HeadIn in, const float* __restrict__ prep, int64_t n_pad, const float* __restrict__ loss_aux,
                   const float* __restrict__ grad_loss, float* __restrict__ grad_h, float* __restrict__ workspace,
                   float* __restrict__ loss_part, const float* __restrict__ projA, float* __restrict__ gradA) {
    // Atom-projected mode (projA != NULL; one-pass training path): the first layer of the score MLP is linear in
    // feat_p = h[u_p] + h[v_p], so W0[:, :128] . feat_p = A[u_p] + A[v_p] with A = h W0[:, :128]^T computed once per ATOM
    // by the dense-layer kernel (7.7 k rows instead of 111 k pairs).  MMA1 becomes a gather-add in E1, MMA4 + E4 (dfeat per
    // pair, scattered to both atoms) become a scatter of dz1 into gradA followed by one per-atom GEMM dh = gradA W0[:, :128],
    // and WG0 becomes the per-atom weight gradient gradA^T h: three of the six MMA groups and the feat tile leave this kernel.
    const bool proj = BWD && projA != nullptr;
    using K = HeadCfg<128>;
    using L = HeadSmem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = align1024(smem_raw);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar = sbase + L::BAR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    pdl_launch_dependents();
    pdl_wait();
    int* sU = reinterpret_cast<int*>(smem + L::SC);
    int* sV = sU + 64;
    float* sSig = reinterpret_cast<float*>(sV + 64);
    float* sDt = sSig + 64; float* sTgt = sDt + 64; float* sSa = sTgt + 64; float* sEmb = sSa + 64;
    float* sDsr = sEmb + 64; float* sDemb = sDsr + 64; float* sValid = sDemb + 64;
    float* sRed = reinterpret_cast<float*>(smem + L::RED);

    // ---- weights: W0[:, :128] rows n, W1 rows n2 (K-major images, split)
    for (int idx = tid; idx < (proj ? 0 : 128 * 16); idx += kHThreads) {
        const int n = idx >> 4, c = idx & 15;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(in.p.out_w0 + n * 129 + c * 8 + j);
        store_chunk8<FP16>(smem + L::W0 + (c >> 3) * kHBlkW, smem + L::W0 + 2 * kHBlkW + (c >> 3) * kHBlkW, n, (c & 7) * 8, v);
    }
    for (int idx = tid; idx < 64 * 16; idx += kHThreads) {
        const int n2 = idx >> 4, c = idx & 15;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(in.p.out_w1 + n2 * 128 + c * 8 + j);
        store_chunk8<FP16>(smem + L::W1 + (c >> 3) * kHBlkT, smem + L::W1 + 2 * kHBlkT + (c >> 3) * kHBlkT, n2, (c & 7) * 8, v);
    }
    const uint32_t bar_wg = bar + 8;                           // weight-gradient MMAs retire off the critical path
    if (tid == 0) { mbar_init(bar, 1); mbar_init(bar_wg, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(sbase + L::TMEM_PTR, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + L::TMEM_PTR);
    const uint32_t tDW0 = tmem, tDW1 = tmem + 128, tD1 = tmem + 192, tD2 = tmem + 256, tD3 = tmem + 320, tD4 = tmem + 384;

    const int q = warp & 3, cg4 = warp >> 2;                   // TMEM lane quadrant, column (pair) group of 16
    const int Ln = q * 32 + lane;                              // this thread's TMEM lane: hidden unit n / n2 / feature k
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const int col0 = cg4 * 16;
    const float wl = __ldg(in.p.out_w0 + Ln * 129 + 128), b0n = __ldg(in.p.out_b0 + Ln);
    const float b1n = Ln < 64 ? __ldg(in.p.out_b1 + Ln) : 0.f, w2n = Ln < 64 ? __ldg(in.p.out_w2 + Ln) : 0.f;
    const float out_b2 = __ldg(in.p.out_b2), in_b1 = __ldg(in.p.in_b1);
    const int unit = tid & 127, pgrp = tid >> 7;               // emb-gradient mapping: hidden unit x quarter of the tile's pairs
    const float iw0 = __ldg(in.p.in_w0 + unit), ib0 = __ldg(in.p.in_b0 + unit), iw1 = __ldg(in.p.in_w1 + unit);
    float gscale = 0.f;
    if (BWD) {
        if (loss_aux != nullptr) {
            const float ng = __ldg(loss_aux + 1);
            gscale = ng > 0.f ? __ldg(grad_loss) / ng : 0.f;
        } else {
            // fused forward + backward (geossl_ddm_head_fwd_bwd_tc): the number of graphs is not known until every CTA has
            // finished, so the gradients are accumulated UNSCALED (dL/d. of the plain sum over pairs); the caller multiplies
            // them by grad_loss / n_graphs.  The loss partials go to loss_part as in the forward kernel.
            gscale = 1.f;
        }
    }
    constexpr uint32_t fmt = Split<FP16>::kFmt;
    const uint32_t id_t = idesc_f16(fmt, 128, kHP);            // K-major A and B
    const uint32_t id_a_mn = idesc_f16(fmt, 128, kHP, 1, 0);   // MN-major A (weight image), K-major B
    const uint32_t id_wg0 = idesc_f16(fmt, 128, 128, 1, 1), id_wg1 = idesc_f16(fmt, 128, 64, 1, 1);

    float loss_acc = 0.f;
    int gmax = -1;
    float a_db0 = 0.f, a_dwl = 0.f, a_db1 = 0.f, a_dw2 = 0.f, a_db2 = 0.f, a_iw0 = 0.f, a_ib0 = 0.f, a_iw1 = 0.f, a_ib1 = 0.f;
    uint32_t phase = 0, phase_wg = 0;
    int done = 0;
    const int64_t n_tiles = (in.live() + kHP - 1) / kHP;
    float pre[kPrepRows];
    if (tid < kHP && blockIdx.x < n_tiles) {
#pragma unroll
        for (int r = 0; r < kPrepRows; ++r) pre[r] = __ldg(prep + r * n_pad + (int64_t)blockIdx.x * kHP + tid);
    }
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++done) {
        trace_h(done, 0);
        // ---- 0. per-pair scalars: precomputed by ddm_pair_prep_kernel, the next tile's are prefetched into registers
        if (tid < kHP) {
            sU[tid] = __float_as_int(pre[0]); sV[tid] = __float_as_int(pre[1]);
            const int g = __float_as_int(pre[2]);
            gmax = max(gmax, g);
            sSig[tid] = pre[3]; sDt[tid] = pre[4]; sTgt[tid] = pre[5]; sSa[tid] = pre[6]; sEmb[tid] = pre[7];
            sValid[tid] = g >= 0 ? 1.f : 0.f;
        }
        __syncthreads();
        trace_h(done, 1);
        if (tid < kHP && t + gridDim.x < n_tiles) {
            const int64_t pn = (t + gridDim.x) * kHP + tid;
#pragma unroll
            for (int r = 0; r < kPrepRows; ++r) pre[r] = __ldg(prep + r * n_pad + pn);
        }
        trace_h(done, 2);
        // ---- 1. the feat tile (lanes over columns: coalesced row gathers)
        if (BWD && done > 0) {                                 // WG0/WG1 of the previous tile still read FEAT / Z1 / dZ1 / dZ2
            mbar_wait(bar_wg, phase_wg); phase_wg ^= 1;
            tc_fence_after();
        }
        if (!proj) {
            const int cg = tid & 15, ro = tid >> 4;             // 8 columns 8cg.., rows ro and ro + 32
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                const int row = ro + 32 * s2;
                const float* hu = in.h + (int64_t)sU[row] * 128 + cg * 8;
                const float* hv = in.h + (int64_t)sV[row] * 128 + cg * 8;
                const float4 a0 = ldg4(hu), a1 = ldg4(hu + 4), c0 = ldg4(hv), c1 = ldg4(hv + 4);
                const float m = sValid[row];
                float v[8] = {(a0.x + c0.x) * m, (a0.y + c0.y) * m, (a0.z + c0.z) * m, (a0.w + c0.w) * m,
                              (a1.x + c1.x) * m, (a1.y + c1.y) * m, (a1.z + c1.z) * m, (a1.w + c1.w) * m};
                store_chunk8<FP16>(smem + L::FEAT + (cg >> 3) * kHBlkT, smem + L::FEAT + 2 * kHBlkT + (cg >> 3) * kHBlkT, row,
                                   (cg & 7) * 8, v);
            }
        }
        fence_proxy_async();
        __syncthreads();
        trace_h(done, 3);
        // ---- 2. MMA1: D1^T = W0h . feat^T
        if (!proj && warp == 0) {                             // uniform operands; the elected lane issues
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t ah = desc_k_sw128(sbase + L::W0), al = desc_k_sw128(sbase + L::W0 + 2 * kHBlkW);
                const uint64_t bh = desc_k_sw128(sbase + L::FEAT), bl = desc_k_sw128(sbase + L::FEAT + 2 * kHBlkT);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t oa = (ks >> 2) * (kHBlkW >> 4) + 2 * (ks & 3), ob = (ks >> 2) * (kHBlkT >> 4) + 2 * (ks & 3);
                    mma3(tD1, ah + oa, al + oa, bh + ob, bl + ob, id_t, ks > 0);
                }
                tc_commit(bar);
            }
            __syncwarp();
        }
        if (!proj) {
            mbar_wait(bar, phase); phase ^= 1;
            tc_fence_after();
        }
        trace_h(done, 4);
        // ---- 3. E1: z1 = relu(D1^T + emb_p * wl + b0) -> Z1 tile [p][n]
        uint32_t mask1 = 0;
        {
            float v[16];
            if (!proj) {
                tmem_ld16(tD1 + lane_base + col0, v);
            } else {                                           // lanes = hidden units: 128 contiguous bytes per warp load
                float au[16], av[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    au[j] = __ldg(projA + (int64_t)sU[col0 + j] * 128 + Ln);
                    av[j] = __ldg(projA + (int64_t)sV[col0 + j] * 128 + Ln);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = (au[j] + av[j]) * sValid[col0 + j];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                v[j] = fmaxf(fmaf(sEmb[col0 + j], wl, v[j]) + b0n, 0.f);
                if (v[j] > 0.f) mask1 |= 1u << j;
            }
            store_split16_paired<FP16>(smem + L::Z1 + (Ln >> 6) * kHBlkT, smem + L::Z1 + 2 * kHBlkT + (Ln >> 6) * kHBlkT, col0, Ln & 63, v, lane);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        trace_h(done, 5);
        // ---- 4. MMA2: D2^T = W1 . z1^T   (rows n2 >= 64 of the A operand are garbage lanes that are never read)
        if (warp == 0) {                                      // uniform operands; the elected lane issues
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t ah = desc_k_sw128(sbase + L::W1), al = desc_k_sw128(sbase + L::W1 + 2 * kHBlkT);
                const uint64_t bh = desc_k_sw128(sbase + L::Z1), bl = desc_k_sw128(sbase + L::Z1 + 2 * kHBlkT);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t o = (ks >> 2) * (kHBlkT >> 4) + 2 * (ks & 3);
                    mma3(tD2, ah + o, al + o, bh + o, bl + o, id_t, ks > 0);
                }
                tc_commit(bar);
            }
            __syncwarp();
        }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        trace_h(done, 6);
        // ---- 5. E2: z2, score (sum over the 64 lanes n2), loss / dsr, dz2
        float z2[16];
        {
            float v[16];
            tmem_ld16(tD2 + lane_base + col0, v);
            float ps[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                z2[j] = Ln < 64 ? fmaxf(v[j] + b1n, 0.f) : 0.f;
                ps[j] = z2[j] * w2n;
            }
            const float tot = warp_reduce16(ps, lane);
            if ((lane & 1) == 0 && q < 2) sRed[q * 64 + col0 + reduce16_index(lane)] = tot;
        }
        tc_fence_before();
        __syncthreads();
        if (tid < kHP) {
            const float inv = 1.f / sSig[tid];
            const float s = (sRed[tid] + sRed[64 + tid] + out_b2) * inv;
            const float diff = s - sTgt[tid];
            loss_acc += 0.5f * (diff * diff) * sSa[tid];
            sDsr[tid] = diff * sSa[tid] * gscale * inv;
        }
        if (!BWD) { __syncthreads(); trace_h(done, 7); continue; }
        __syncthreads();
        trace_h(done, 7);
        if (Ln < 64) {                                         // (whole warps: quadrants 0 and 1)
            float dzv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float dsr = sDsr[col0 + j];
                dzv[j] = z2[j] > 0.f ? dsr * w2n : 0.f;
                a_dw2 = fmaf(dsr, z2[j], a_dw2);
                a_db1 += dzv[j];
            }
            store_split16_paired<FP16>(smem + L::DZ2, smem + L::DZ2 + kHBlkT, col0, Ln, dzv, lane);
        }
        if (tid < kHP) a_db2 += sDsr[tid];
        fence_proxy_async();
        __syncthreads();
        trace_h(done, 8);
        // ---- 6. MMA3: D3^T = W1^T . dz2^T (A = MN-major view of the W1 image, K = 64) ; WG1: DW1^T += z1^T dz2
        if (warp == 0) {                                      // uniform operands; the elected lane issues
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t ah = desc_mn_sw128(sbase + L::W1, kHBlkT), al = desc_mn_sw128(sbase + L::W1 + 2 * kHBlkT, kHBlkT);
                const uint64_t bh = desc_k_sw128(sbase + L::DZ2), bl = desc_k_sw128(sbase + L::DZ2 + kHBlkT);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma3(tD3, ah + ks * 128, al + ks * 128, bh + 2 * ks, bl + 2 * ks, id_a_mn, ks > 0);
                tc_commit(bar);                                // E3 waits for MMA3 only; WG1 runs behind it
                const uint64_t zh = desc_mn_sw128(sbase + L::Z1, kHBlkT), zl = desc_mn_sw128(sbase + L::Z1 + 2 * kHBlkT, kHBlkT);
                const uint64_t dh = desc_mn_sw128(sbase + L::DZ2, kHBlkT), dl = desc_mn_sw128(sbase + L::DZ2 + kHBlkT, kHBlkT);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma3(tDW1, zh + ks * 128, zl + ks * 128, dh + ks * 128, dl + ks * 128, id_wg1, (done | ks) > 0);
                if (proj) tc_commit(bar_wg);                   // (no WG0 in this mode: WG1 alone retires behind the critical path)
            }
            __syncwarp();
        }
        mbar_wait(bar, phase); phase ^= 1;
        tc_fence_after();
        trace_h(done, 9);
        // ---- 7. E3: dz1 = D3^T * [z1 > 0] -> dZ1 tile ; db0, dW0[:,128] in-thread ; demb_p = sum_n dz1 wl (over lanes)
        {
            float v[16];
            tmem_ld16(tD3 + lane_base + col0, v);
            float pd[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                v[j] = ((mask1 >> j) & 1u) ? v[j] : 0.f;
                a_db0 += v[j];
                a_dwl = fmaf(v[j], sEmb[col0 + j], a_dwl);
                pd[j] = v[j] * wl;
            }
            if (!proj) {
                store_split16_paired<FP16>(smem + L::DZ1 + (Ln >> 6) * kHBlkT, smem + L::DZ1 + 2 * kHBlkT + (Ln >> 6) * kHBlkT, col0, Ln & 63, v, lane);
            } else {
                // d/dA[u_p] = d/dA[v_p] = dz1_p: coalesced RED (lanes = hidden units), runs of equal first atom collapsed
                int cur_u = -1;
                float run = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (sValid[col0 + j] != 0.f) {
                        const int uj = sU[col0 + j];
                        atomicAdd(gradA + (int64_t)sV[col0 + j] * 128 + Ln, v[j]);
                        if (uj != cur_u) {
                            if (cur_u >= 0) atomicAdd(gradA + (int64_t)cur_u * 128 + Ln, run);
                            cur_u = uj;
                            run = v[j];
                        } else {
                            run += v[j];
                        }
                    }
                }
                if (cur_u >= 0) atomicAdd(gradA + (int64_t)cur_u * 128 + Ln, run);
            }
            const float tot = warp_reduce16(pd, lane);
            if ((lane & 1) == 0) sRed[q * 64 + col0 + reduce16_index(lane)] = tot;
        }
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
        trace_h(done, 10);
        // ---- 8. MMA4: D4^T = W0h^T . dz1^T (A = MN-major view of the W0 image, K = 128) ; WG0: DW0 += dz1^T feat
        if (!proj && warp == 0) {                             // uniform operands; the elected lane issues
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t ah = desc_mn_sw128(sbase + L::W0, kHBlkW), al = desc_mn_sw128(sbase + L::W0 + 2 * kHBlkW, kHBlkW);
                const uint64_t bh = desc_k_sw128(sbase + L::DZ1), bl = desc_k_sw128(sbase + L::DZ1 + 2 * kHBlkT);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint32_t ob = (ks >> 2) * (kHBlkT >> 4) + 2 * (ks & 3);
                    mma3(tD4, ah + ks * 128, al + ks * 128, bh + ob, bl + ob, id_a_mn, ks > 0);
                }
                tc_commit(bar);                                // E4 waits for MMA4 only
                const uint64_t zh = desc_mn_sw128(sbase + L::DZ1, kHBlkT), zl = desc_mn_sw128(sbase + L::DZ1 + 2 * kHBlkT, kHBlkT);
                const uint64_t fh = desc_mn_sw128(sbase + L::FEAT, kHBlkT), fl = desc_mn_sw128(sbase + L::FEAT + 2 * kHBlkT, kHBlkT);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) mma3(tDW0, zh + ks * 128, zl + ks * 128, fh + ks * 128, fl + ks * 128, id_wg0, (done | ks) > 0);
                tc_commit(bar_wg);                             // covers WG1 and WG0 of this tile
            }
            __syncwarp();
        }
        if (tid < kHP) sDemb[tid] = sRed[tid] + sRed[64 + tid] + sRed[128 + tid] + sRed[192 + tid];
        if (!proj) {
            mbar_wait(bar, phase); phase ^= 1;
            tc_fence_after();
        }
        __syncthreads();
        trace_h(done, 11);
        // ---- 9. E4: scatter dfeat to both endpoints (lanes = features: 128 contiguous bytes per warp instruction)
        if (!proj) {
            float v[16];
            tmem_ld16(tD4 + lane_base + col0, v);
            // consecutive pairs of a molecule share their first atom (itertools order): one RED per run of equal u
            int cur_u = -1;
            float run = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                if (sValid[col0 + j] != 0.f) {
                    const int uj = sU[col0 + j];
                    atomicAdd(grad_h + (int64_t)sV[col0 + j] * 128 + Ln, v[j]);
                    if (uj != cur_u) {
                        if (cur_u >= 0) atomicAdd(grad_h + (int64_t)cur_u * 128 + Ln, run);
                        cur_u = uj;
                        run = v[j];
                    } else {
                        run += v[j];
                    }
                }
            }
            if (cur_u >= 0) atomicAdd(grad_h + (int64_t)cur_u * 128 + Ln, run);
        }
        trace_h(done, 12);
        {   // distance-embedding MLP gradients: thread = (hidden unit, quarter of the tile's pairs)
#pragma unroll 4
            for (int p = pgrp * 16; p < pgrp * 16 + 16; ++p) {
                const float dt = sDt[p], de = sDemb[p];
                const float pa = fmaf(iw0, dt, ib0);
                if (pa > 0.f) {
                    a_iw1 = fmaf(de, pa, a_iw1);
                    const float dpre = de * iw1;
                    a_iw0 = fmaf(dpre, dt, a_iw0);
                    a_ib0 += dpre;
                }
                if (unit == 0) a_ib1 += de;
            }
        }
        tc_fence_before();
        __syncthreads();
        trace_h(done, 13);
    }

    if (!BWD || loss_part != nullptr) {
        __shared__ float red_l[16];
        __shared__ int red_g[16];
        float* lp = BWD ? loss_part : workspace;
        loss_acc = warp_sum(loss_acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gmax = max(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        if (lane == 0) { red_l[warp] = loss_acc; red_g[warp] = gmax; }
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            int g = -1;
            for (int w = 0; w < 16; ++w) { s += red_l[w]; g = max(g, red_g[w]); }
            lp[2 * blockIdx.x] = s;
            lp[2 * blockIdx.x + 1] = (float)g;
        }
    }
    if (BWD) {
        // ---- per-CTA partial gradients in HeadCfg<128>'s layout (reduced by ddm_head_reduce_kernel<128>)
        float* ws = workspace + (int64_t)blockIdx.x * K::kPartial;
        if (done > 0) mbar_wait(bar_wg, phase_wg);             // the last tile's weight-gradient MMAs
        tc_fence_after();
        {   // DW0 [n][k]: lane n, this warp's column group covers k = cg4*32 .. +31
            float v[32];
            if (done > 0 && !proj) tmem_ld32(tDW0 + lane_base + cg4 * 32, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) ws[K::pW0 + Ln * K::LD + cg4 * 32 + j] = (done > 0 && !proj) ? v[j] : 0.f;   // (proj: per-atom wgrad)
        }
        {   // DW1^T [n][n2] -> pW1 [n2][n]
            float v[16];
            if (done > 0) tmem_ld16(tDW1 + lane_base + cg4 * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) ws[K::pW1 + (cg4 * 16 + j) * 128 + Ln] = done > 0 ? v[j] : 0.f;
        }
        tc_fence_before();
        __syncthreads();
        // column-group partials of the in-thread accumulators: reduce over the 4 warps that share a lane quadrant
        float* red = reinterpret_cast<float*>(smem + L::FEAT);      // [4 cg][128] x 4 quantities
        red[(0 * 4 + cg4) * 128 + Ln] = a_db0;
        red[(1 * 4 + cg4) * 128 + Ln] = a_dwl;
        red[(2 * 4 + cg4) * 128 + Ln] = a_db1;
        red[(3 * 4 + cg4) * 128 + Ln] = a_dw2;
        __syncthreads();
        if (tid < 128) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            for (int c = 0; c < 4; ++c) {
                s0 += red[(0 * 4 + c) * 128 + tid]; s1 += red[(1 * 4 + c) * 128 + tid];
                s2 += red[(2 * 4 + c) * 128 + tid]; s3 += red[(3 * 4 + c) * 128 + tid];
            }
            ws[K::pB0 + tid] = s0;
            ws[K::pW0 + tid * K::LD + 128] = s1;
            if (tid < 64) { ws[K::pB1 + tid] = s2; ws[K::pW2 + tid] = s3; }
        }
        __syncthreads();
        red[(0 * 4 + pgrp) * 128 + unit] = a_iw0;                    // the four pair-quarters of every hidden unit
        red[(1 * 4 + pgrp) * 128 + unit] = a_ib0;
        red[(2 * 4 + pgrp) * 128 + unit] = a_iw1;
        if (unit == 0) red[3 * 4 * 128 + pgrp] = a_ib1;
        __syncthreads();
        if (tid < 128) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
            for (int c = 0; c < 4; ++c) {
                s0 += red[(0 * 4 + c) * 128 + tid]; s1 += red[(1 * 4 + c) * 128 + tid]; s2 += red[(2 * 4 + c) * 128 + tid];
            }
            ws[K::pIW0 + tid] = s0; ws[K::pIB0 + tid] = s1; ws[K::pIW1 + tid] = s2;
        }
        if (tid == 0) ws[K::pIB1] = (red[3 * 4 * 128] + red[3 * 4 * 128 + 1]) + (red[3 * 4 * 128 + 2] + red[3 * 4 * 128 + 3]);
        // out_b2: sum over the 64 scalar threads
        if (warp < 2) {
            float s = warp_sum(a_db2);
            if (lane == 0) sRed[warp] = s;
        }
        __syncthreads();
        if (tid == 0) ws[K::pB2] = sRed[0] + sRed[1];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace tc
}  // namespace geossl

extern "C" {

int geossl_debug_set_trace_head(long long* device_buffer) {
    GEOSSL_CUDA(cudaMemcpyToSymbol(tc::g_trace_head, &device_buffer, sizeof(device_buffer)));
    return 0;
}


int geossl_ddm_head_fwd_tc(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                           const float* dist, const float* noise, const int64_t* noise_level,
                           const float* sigmas, int n_levels, float anneal_power, int H,
                           const geossl_ddm_params* params, float* workspace, float* loss, void* stream) {
    GEOSSL_REQUIRE(workspace && loss && n_pairs >= 0, "null workspace/loss");
    GEOSSL_REQUIRE(H == 128, "the tensor-core DDM head is built for emb_dim = 128");
    if (n_pairs == 0) {
        GEOSSL_CUDA(cudaMemsetAsync(loss, 0, 2 * sizeof(float), as_stream(stream)));
        return 0;
    }
    HeadIn in;
    GEOSSL_REQUIRE(fill_head_in(in, h, sei, batch, n_pairs, n_pairs_live, dist, noise, noise_level, sigmas, n_levels, anneal_power, params) == 0,
                   "null input pointer");
    const size_t smem = tc::HeadSmem::kBytes + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::ddm_head_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    const int grid = head_grid(n_pairs);
    const int64_t n_pad = head_pad(n_pairs);
    float* prep = workspace + head_prep_offset();
    GEOSSL_CUDA(launch_pdl(tc::ddm_pair_prep_kernel, dim3((unsigned)((n_pad + 255) / 256)), dim3(256), 0, as_stream(stream), in, n_pad, prep));
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(tc::ddm_head_tc_kernel<true, false>, dim3(grid), dim3(tc::kHThreads), smem, as_stream(stream), in,
                           (const float*)prep, n_pad, (const float*)nullptr, (const float*)nullptr, (float*)nullptr, workspace,
                           (float*)nullptr, (const float*)nullptr, (float*)nullptr));
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(ddm_loss_finalize_kernel, dim3(1), dim3(32), 0, as_stream(stream), (const float*)workspace, grid, loss));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_ddm_head_bwd_tc(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                        int64_t n_atoms,
                           const float* dist, const float* noise, const int64_t* noise_level,
                           const float* sigmas, int n_levels, float anneal_power, int H,
                           const geossl_ddm_params* params, const float* loss_aux, const float* grad_loss, float* workspace,
                           float* grad_h, const geossl_ddm_grads* grads, void* stream) {
    GEOSSL_REQUIRE(workspace && grad_h && grads && loss_aux && grad_loss && n_pairs > 0 && n_atoms >= 0, "null pointer / empty input");
    GEOSSL_REQUIRE(H == 128, "the tensor-core DDM head is built for emb_dim = 128");
    const geossl_ddm_grads& g = *grads;
    GEOSSL_REQUIRE(g.in_w0 && g.in_b0 && g.in_w1 && g.in_b1 && g.out_w0 && g.out_b0 && g.out_w1 && g.out_b1 && g.out_w2 && g.out_b2,
                   "null gradient pointer");
    HeadIn in;
    GEOSSL_REQUIRE(fill_head_in(in, h, sei, batch, n_pairs, n_pairs_live, dist, noise, noise_level, sigmas, n_levels, anneal_power, params) == 0,
                   "null input pointer");
    const size_t smem = tc::HeadSmem::kBytes + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::ddm_head_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    GEOSSL_CUDA(cudaMemsetAsync(grad_h, 0, sizeof(float) * (size_t)n_atoms * 128, as_stream(stream)));
    const int grid = head_grid(n_pairs);
    const int64_t n_pad = head_pad(n_pairs);
    float* prep = workspace + head_prep_offset();
    GEOSSL_CUDA(launch_pdl(tc::ddm_pair_prep_kernel, dim3((unsigned)((n_pad + 255) / 256)), dim3(256), 0, as_stream(stream), in, n_pad, prep));
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(tc::ddm_head_tc_kernel<false, true>, dim3(grid), dim3(tc::kHThreads), smem, as_stream(stream), in,
                           (const float*)prep, n_pad, loss_aux, grad_loss, grad_h, workspace, (float*)nullptr,
                           (const float*)nullptr, (float*)nullptr));
    GEOSSL_LAUNCH_CHECK();
    const int n = HeadCfg<128>::kPartial;
    GEOSSL_CUDA(launch_pdl(ddm_head_reduce_kernel<128>, dim3((n + 255) / 256), dim3(256), 0, as_stream(stream), (const float*)workspace, grid, g));
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_ddm_head_fwd_bwd_tc(const float* h, const int64_t* sei, const int64_t* batch, int64_t n_pairs, const int32_t* n_pairs_live,
                               int64_t n_atoms, const float* dist, const float* noise, const int64_t* noise_level,
                               const float* sigmas, int n_levels, float anneal_power, int H,
                               const geossl_ddm_params* params, float* workspace, float* loss, float* grad_h,
                               const geossl_ddm_grads* grads, void* stream) {
    GEOSSL_REQUIRE(workspace && loss && grad_h && grads && n_pairs > 0 && n_atoms >= 0, "null pointer / empty input");
    GEOSSL_REQUIRE(H == 128, "the tensor-core DDM head is built for emb_dim = 128");
    const geossl_ddm_grads& g = *grads;
    GEOSSL_REQUIRE(g.in_w0 && g.in_b0 && g.in_w1 && g.in_b1 && g.out_w0 && g.out_b0 && g.out_w1 && g.out_b1 && g.out_w2 && g.out_b2,
                   "null gradient pointer");
    HeadIn in;
    GEOSSL_REQUIRE(fill_head_in(in, h, sei, batch, n_pairs, n_pairs_live, dist, noise, noise_level, sigmas, n_levels, anneal_power, params) == 0,
                   "null input pointer");
    const size_t smem = tc::HeadSmem::kBytes + 1024;
    static PerDeviceFlag configured;
    if (!configured.get()) {
        GEOSSL_CUDA(cudaFuncSetAttribute(tc::ddm_head_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured.set();
    }
    cudaStream_t st = as_stream(stream);
    const int grid = head_grid(n_pairs);
    const int64_t n_pad = head_pad(n_pairs);
    float* prep = workspace + head_prep_offset();
    float* loss_part = workspace + head_loss_offset();
    // atom-projected first layer (see the kernel): A = h W0[:, :128]^T per atom, gradA scattered per pair, then per atom
    // grad_h = gradA W0[:, :128] and dW0[:, :128] = gradA^T h
    float* projA = prep + kPrepRowsHost * n_pad + 64;
    float* gradA = projA + (size_t)n_atoms * 128;
    uint8_t* images = reinterpret_cast<uint8_t*>(gradA + (size_t)n_atoms * 128);              // forward (fp16) | transposed (bf16)
    float* wg_ws = reinterpret_cast<float*>(images + 2 * geossl_weight_image_bytes());
    int rc = geossl_pack_weight_pair(params->out_w0, 129, images, stream);
    if (rc) return rc;
    rc = geossl_linear_tc(h, n_atoms, images, nullptr, 0, nullptr, nullptr, projA, 0, stream);
    if (rc) return rc;
    GEOSSL_CUDA(cudaMemsetAsync(gradA, 0, sizeof(float) * (size_t)n_atoms * 128, st));
    GEOSSL_CUDA(launch_pdl(tc::ddm_pair_prep_kernel, dim3((unsigned)((n_pad + 255) / 256)), dim3(256), 0, st, in, n_pad, prep));
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(tc::ddm_head_tc_kernel<false, true>, dim3(grid), dim3(tc::kHThreads), smem, st, in,
                           (const float*)prep, n_pad, (const float*)nullptr, (const float*)nullptr, grad_h, workspace, loss_part,
                           (const float*)projA, gradA));
    GEOSSL_LAUNCH_CHECK();
    GEOSSL_CUDA(launch_pdl(ddm_loss_finalize_kernel, dim3(1), dim3(32), 0, st, (const float*)loss_part, grid, loss));
    GEOSSL_LAUNCH_CHECK();
    const int n = HeadCfg<128>::kPartial;
    GEOSSL_CUDA(launch_pdl(ddm_head_reduce_kernel<128>, dim3((n + 255) / 256), dim3(256), 0, st, (const float*)workspace, grid, g));
    GEOSSL_LAUNCH_CHECK();
    // per-atom tail of the projected first layer (after the reduction has written out_w0, whose [:, :128] block is replaced)
    rc = geossl_linear_tc(gradA, n_atoms, images + geossl_weight_image_bytes(), nullptr, 0, nullptr, nullptr, grad_h, 1, stream);
    if (rc) return rc;
    return geossl_linear_wgrad_tc_block(gradA, 128, h, 128, n_atoms, 0, wg_ws, g.out_w0, 129, nullptr, 128, stream);
}

int64_t geossl_ddm_workspace_fused(int64_t n_pairs, int64_t n_atoms) {
    const int64_t na = n_atoms > 0 ? n_atoms : 0;
    return geossl_ddm_workspace_tc(n_pairs) + 2 * na * 128 + 2 * geossl_weight_image_bytes() / 4 + geossl_linear_wgrad_tc_workspace(na > 0 ? na : 1) + 64;
}

}  // extern "C"
