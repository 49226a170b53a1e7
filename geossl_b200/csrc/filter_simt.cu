// SchNet filter network, fp32 SIMT edition (exact-fp32 parity path and the materialised-W comparison
// point of SURVEY.md 8d).  One persistent CTA per SM walks 64-edge tiles:
//
//   forward   d_e -> rbf (smem) -> Lin1 -> ssp (smem) -> Lin2 -> * cosine cutoff -> W_e (E,F)
//   backward  recompute rbf/Lin1/ssp, dU = x[src]*g[tgt]*C (never materialised), ds = dU W2,
//             da = ds*sigmoid(a); per-CTA register accumulators for dW2 = dU^T s, dW1 = da^T rbf,
//             db2, db1; partials -> workspace -> fixed-order reduction (deterministic, atomic free).
//
// Replaces GaussianSmearing.forward (Geom3D/models/schnet.py:205-207), InteractionBlock.mlp
// (schnet.py:141-145), the cutoff and product in CFConv.forward (schnet.py:186-187) and their autograd.
//
// Thread layout (256 threads): TX = F/4 lanes own features {tx + TX*j, j<4} (stride-TX ownership makes
// both the [k][F] weight reads and the transposed [feature][edge] tile stores bank-conflict minimal),
// TY = 256/TX groups own ME = 64/TY consecutive edges.  All tile operands are k-major in shared memory
// with row stride S = 68 floats (16 B aligned, rows 4 banks apart).
#include "common.cuh"
#include "simt_tile.cuh"

namespace geossl {

// rbf tile: sPhi[g][e] = exp(coeff * (d_e - offset_g)^2), rows G..GP-1 zero; sC[e] = cutoff(d_e); sD[e] = d_e
template <int F>
__device__ __forceinline__ void build_rbf_tile(const float* __restrict__ edge_dist, int64_t e_base, int64_t n_edges,
                                               const float* __restrict__ sOff, float coeff, float cutoff, int G,
                                               float* __restrict__ sPhi, float* __restrict__ sC, float* __restrict__ sD) {
    using C = FCfg<F>;
    const int tid = threadIdx.x;
    if (tid < C::TE) {
        const int64_t e = e_base + tid;
        const float d = (e < n_edges) ? __ldg(edge_dist + e) : 0.f;
        sD[tid] = d;
        sC[tid] = (e < n_edges) ? cosine_cutoff(d, cutoff) : 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < C::GP * C::TE; idx += 256) {
        const int g = idx / C::TE, e = idx % C::TE;
        float v = 0.f;
        if (g < G) {
            const float diff = sD[e] - sOff[g];
            v = expf(__fmul_rn(coeff, __fmul_rn(diff, diff)));   // coeff * pow(diff, 2), schnet.py:207
        }
        sPhi[g * C::S + e] = v;
    }
}

template <int F>
struct FwdSmem {
    using C = FCfg<F>;
    static constexpr int kW1t = 0;                              // [GP][F]   W1^T (rows >= G zero)
    static constexpr int kW2t = kW1t + C::GP * F;               // [F][F]    W2^T : [in][out]
    static constexpr int kPhi = kW2t + F * F;                   // [GP][S]
    static constexpr int kS = kPhi + C::GP * C::S;              // [F][S]
    static constexpr int kB1 = kS + F * C::S;
    static constexpr int kB2 = kB1 + F;
    static constexpr int kOff = kB2 + F;                        // [GP]
    static constexpr int kC = kOff + C::GP;                     // [TE]
    static constexpr int kD = kC + C::TE;                       // [TE]
    static constexpr int kFloats = kD + C::TE;
};

template <int F>
__global__ void __launch_bounds__(256, 1)
filter_fwd_kernel(const float* __restrict__ edge_dist, const int32_t* __restrict__ n_edges_dev, int64_t capacity,
                  const float* __restrict__ offset, float coeff, float cutoff, int G,
                  const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ b2, float* __restrict__ filt) {
    using C = FCfg<F>;
    using L = FwdSmem<F>;
    extern __shared__ __align__(16) float smem[];
    float* sW1t = smem + L::kW1t; float* sW2t = smem + L::kW2t; float* sPhi = smem + L::kPhi; float* sS = smem + L::kS;
    float* sB1 = smem + L::kB1; float* sB2 = smem + L::kB2; float* sOff = smem + L::kOff; float* sC = smem + L::kC;
    float* sD = smem + L::kD;
    const int tid = threadIdx.x, tx = tid % C::TX, ty = tid / C::TX, e0 = ty * C::ME;
    int64_t n_edges = (int64_t)(*n_edges_dev);
    if (n_edges > capacity) n_edges = capacity;

    for (int idx = tid; idx < C::GP * F; idx += 256) {
        const int g = idx / F, f = idx % F;
        sW1t[idx] = (g < G) ? __ldg(w1 + f * G + g) : 0.f;
    }
    for (int idx = tid; idx < F * F; idx += 256) {
        const int i = idx / F, o = idx % F;
        sW2t[idx] = __ldg(w2 + o * F + i);
    }
    if (tid < F) { sB1[tid] = __ldg(b1 + tid); sB2[tid] = __ldg(b2 + tid); }
    if (tid < C::GP) sOff[tid] = (tid < G) ? __ldg(offset + tid) : 0.f;
    __syncthreads();

    const int64_t n_tiles = (n_edges + C::TE - 1) / C::TE;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t e_base = t * C::TE;
        build_rbf_tile<F>(edge_dist, e_base, n_edges, sOff, coeff, cutoff, G, sPhi, sC, sD);
        __syncthreads();
        float acc[C::ME][4];
#pragma unroll
        for (int i = 0; i < C::ME; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        gemm_rows<F>(acc, sPhi, sW1t, F, 1, G, tx, e0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = tx + C::TX * j;
            const float bias = sB1[f];
#pragma unroll
            for (int i = 0; i < C::ME; ++i) sS[f * C::S + e0 + i] = ssp(acc[i][j] + bias);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < C::ME; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        gemm_rows<F>(acc, sS, sW2t, F, 1, F, tx, e0);
#pragma unroll
        for (int i = 0; i < C::ME; ++i) {
            const int64_t e = e_base + e0 + i;
            if (e < n_edges) {
                const float c = sC[e0 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int f = tx + C::TX * j;
                    filt[e * F + f] = (acc[i][j] + sB2[f]) * c;
                }
            }
        }
        __syncthreads();
    }
}

template <int F>
struct BwdSmem {
    using C = FCfg<F>;
    static constexpr int kW1t = 0;                              // [GP][F]
    static constexpr int kW2 = kW1t + C::GP * F;                // [F][F]  row-major W2: [out][in]
    static constexpr int kPhi = kW2 + F * F;                    // [GP][S]
    static constexpr int kS = kPhi + C::GP * C::S;              // [F][S]  s, later da
    static constexpr int kDU = kS + F * C::S;                   // [F][S]  dU^T
    static constexpr int kB1 = kDU + F * C::S;
    static constexpr int kOff = kB1 + F;
    static constexpr int kC = kOff + C::GP;
    static constexpr int kD = kC + C::TE;
    static constexpr int kSrc = kD + C::TE;                     // int [TE]
    static constexpr int kTgt = kSrc + C::TE;                   // int [TE]
    static constexpr int kFloats = kTgt + C::TE;
};

// per-CTA partial layout in the workspace
template <int F>
struct Partial {
    static constexpr int kW2 = 0;                               // [o][i]
    static constexpr int kW1 = F * F;                           // [g][f], g < GP
    static constexpr int kB2 = kW1 + FCfg<F>::GP * F;
    static constexpr int kB1 = kB2 + F;
    static constexpr int kFloats = kB1 + F;
};

template <int F>
__global__ void __launch_bounds__(256, 1)
filter_bwd_kernel(const float* __restrict__ edge_dist, const int32_t* __restrict__ n_edges_dev, int64_t capacity,
                  const float* __restrict__ offset, float coeff, float cutoff, int G,
                  const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                  const float* __restrict__ x, const float* __restrict__ grad_out,
                  const int32_t* __restrict__ src, const int32_t* __restrict__ edge_tgt,
                  const float* __restrict__ grad_filt, float* __restrict__ workspace) {
    using C = FCfg<F>;
    using L = BwdSmem<F>;
    using P = Partial<F>;
    extern __shared__ __align__(16) float smem[];
    float* sW1t = smem + L::kW1t; float* sW2 = smem + L::kW2; float* sPhi = smem + L::kPhi; float* sS = smem + L::kS;
    float* sDU = smem + L::kDU; float* sB1 = smem + L::kB1; float* sOff = smem + L::kOff; float* sC = smem + L::kC;
    float* sD = smem + L::kD;
    int* sSrc = reinterpret_cast<int*>(smem + L::kSrc); int* sTgt = reinterpret_cast<int*>(smem + L::kTgt);
    const int tid = threadIdx.x, tx = tid % C::TX, ty = tid / C::TX, e0 = ty * C::ME;
    int64_t n_edges = (int64_t)(*n_edges_dev);
    if (n_edges > capacity) n_edges = capacity;

    for (int idx = tid; idx < C::GP * F; idx += 256) {
        const int g = idx / F, f = idx % F;
        sW1t[idx] = (g < G) ? __ldg(w1 + f * G + g) : 0.f;
    }
    for (int idx = tid; idx < F * F; idx += 256) sW2[idx] = __ldg(w2 + idx);
    if (tid < F) sB1[tid] = __ldg(b1 + tid);
    if (tid < C::GP) sOff[tid] = (tid < G) ? __ldg(offset + tid) : 0.f;

    float aW2[C::MO][4], aW1[C::MG][4], aB2[4], aB1[4];
#pragma unroll
    for (int i = 0; i < C::MO; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) aW2[i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < C::MG; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) aW1[i][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { aB2[j] = 0.f; aB1[j] = 0.f; }
    __syncthreads();

    const int64_t n_tiles = (n_edges + C::TE - 1) / C::TE;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t e_base = t * C::TE;
        if (!grad_filt && tid < C::TE) {
            const int64_t e = e_base + tid;
            sSrc[tid] = (e < n_edges) ? __ldg(src + e) : 0;
            sTgt[tid] = (e < n_edges) ? __ldg(edge_tgt + e) : 0;
        }
        build_rbf_tile<F>(edge_dist, e_base, n_edges, sOff, coeff, cutoff, G, sPhi, sC, sD);
        __syncthreads();

        // dU^T tile: dU[e][o] = dW_e[o] * C_e,  dW_e = x[src_e] * g[tgt_e]  (or the materialised grad_filt)
#pragma unroll
        for (int i = 0; i < C::ME; ++i) {
            const int el = e0 + i;
            const int64_t e = e_base + el;
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (e < n_edges) {
                const float c = sC[el];
                if (grad_filt) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = __ldg(grad_filt + e * F + tx + C::TX * j) * c;
                } else {
                    const float* xr = x + (int64_t)sSrc[el] * F;
                    const float* gr = grad_out + (int64_t)sTgt[el] * F;
#pragma unroll
                    for (int j = 0; j < 4; ++j) v[j] = __ldg(xr + tx + C::TX * j) * __ldg(gr + tx + C::TX * j) * c;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                sDU[(tx + C::TX * j) * C::S + el] = v[j];
                aB2[j] += v[j];
            }
        }

        // recompute a = rbf W1^T + b1 ; s = ssp(a) -> smem ; sigmoid(a) stays in registers
        float acc[C::ME][4], sig[C::ME][4];
#pragma unroll
        for (int i = 0; i < C::ME; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        gemm_rows<F>(acc, sPhi, sW1t, F, 1, G, tx, e0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int f = tx + C::TX * j;
            const float bias = sB1[f];
#pragma unroll
            for (int i = 0; i < C::ME; ++i) {
                const float a = acc[i][j] + bias;
                sS[f * C::S + e0 + i] = ssp(a);
                sig[i][j] = ssp_grad(a);
            }
        }
        __syncthreads();

        // ds = dU W2  (K = out channel), da = ds * sigmoid(a)
#pragma unroll
        for (int i = 0; i < C::ME; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        gemm_rows<F>(acc, sDU, sW2, F, 1, F, tx, e0);
#pragma unroll
        for (int i = 0; i < C::ME; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[i][j] *= sig[i][j];
                aB1[j] += acc[i][j];
            }

        // dW2[o][i] += sum_e dU[e][o] s[e][i]
        gemm_wgrad<F, C::MO>(aW2, sDU, ty * C::MO, F, sS, tx);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < C::ME; ++i) sS[(tx + C::TX * j) * C::S + e0 + i] = acc[i][j];
        __syncthreads();
        // dW1[f][g] += sum_e da[e][f] rbf[e][g]   (accumulated as [g][f])
        gemm_wgrad<F, C::MG>(aW1, sPhi, ty * C::MG, C::GP, sS, tx);
        __syncthreads();
    }

    // column sums of the bias gradients across the TY edge groups (reuse sS / sDU as scratch)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        sS[ty * F + tx + C::TX * j] = aB2[j];
        sDU[ty * F + tx + C::TX * j] = aB1[j];
    }
    __syncthreads();
    float* ws = workspace + (int64_t)blockIdx.x * P::kFloats;
    if (tid < F) {
        float s2 = 0.f, s1 = 0.f;
        for (int r = 0; r < C::TY; ++r) { s2 += sS[r * F + tid]; s1 += sDU[r * F + tid]; }
        ws[P::kB2 + tid] = s2;
        ws[P::kB1 + tid] = s1;
    }
#pragma unroll
    for (int i = 0; i < C::MO; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ws[P::kW2 + (ty * C::MO + i) * F + tx + C::TX * j] = aW2[i][j];
#pragma unroll
    for (int i = 0; i < C::MG; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ws[P::kW1 + (ty * C::MG + i) * F + tx + C::TX * j] = aW1[i][j];
}

template <int F>
__global__ void filter_bwd_reduce_kernel(const float* __restrict__ workspace, int n_parts, int G,
                                         float* __restrict__ gw1, float* __restrict__ gb1,
                                         float* __restrict__ gw2, float* __restrict__ gb2) {
    using P = Partial<F>;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P::kFloats) return;
    // four independent chains (fixed association order => still deterministic) keep enough loads in flight
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int p = 0;
    for (; p + 3 < n_parts; p += 4) {
        s0 += workspace[(int64_t)p * P::kFloats + idx];
        s1 += workspace[(int64_t)(p + 1) * P::kFloats + idx];
        s2 += workspace[(int64_t)(p + 2) * P::kFloats + idx];
        s3 += workspace[(int64_t)(p + 3) * P::kFloats + idx];
    }
    for (; p < n_parts; ++p) s0 += workspace[(int64_t)p * P::kFloats + idx];
    const float s = (s0 + s1) + (s2 + s3);
    if (idx < P::kW1) {
        gw2[idx] = s;
    } else if (idx < P::kB2) {
        const int r = idx - P::kW1, g = r / F, f = r % F;
        if (g < G) gw1[f * G + g] = s;
    } else if (idx < P::kB1) {
        gb2[idx - P::kB2] = s;
    } else {
        gb1[idx - P::kB1] = s;
    }
}

static int filter_grid(int64_t capacity) {
    int64_t tiles = (capacity + 63) / 64;
    return (int)(tiles < kNumSM ? (tiles > 0 ? tiles : 1) : kNumSM);
}

template <int F>
int launch_filter_fwd(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity, const float* offset,
                      float coeff, float cutoff, int G, const float* w1, const float* b1, const float* w2,
                      const float* b2, float* filt, cudaStream_t st) {
    const size_t smem = FwdSmem<F>::kFloats * sizeof(float);
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(filter_fwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    filter_fwd_kernel<F><<<filter_grid(capacity), 256, smem, st>>>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G,
                                                                    w1, b1, w2, b2, filt);
    return 0;
}

template <int F>
int launch_filter_bwd(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity, const float* offset,
                      float coeff, float cutoff, int G, const float* w1, const float* b1, const float* w2,
                      const float* x, const float* grad_out, const int32_t* src, const int32_t* edge_tgt,
                      const float* grad_filt, float* workspace, float* gw1, float* gb1, float* gw2, float* gb2,
                      cudaStream_t st) {
    const size_t smem = BwdSmem<F>::kFloats * sizeof(float);
    static PerDeviceFlag configured;
    if (!configured.get()) {
        cudaError_t e = cudaFuncSetAttribute(filter_bwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    filter_bwd_kernel<F><<<kNumSM, 256, smem, st>>>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2,
                                                    x, grad_out, src, edge_tgt, grad_filt, workspace);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    count_launch();
    const int n = Partial<F>::kFloats;
    filter_bwd_reduce_kernel<F><<<(n + 255) / 256, 256, 0, st>>>(workspace, kNumSM, G, gw1, gb1, gw2, gb2);
    return 0;
}

}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_filter_fwd(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                      const float* offset, float coeff, float cutoff, int G, int F,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      float* filt, void* stream) {
    if (capacity == 0) return 0;
    GEOSSL_REQUIRE(edge_dist && n_edges_dev && offset && w1 && b1 && w2 && b2 && filt && capacity > 0, "null pointer");
    GEOSSL_REQUIRE(G >= 1 && G <= 64, "num_gaussians must be in [1,64]");
    int rc;
    switch (F) {
        case 32: rc = launch_filter_fwd<32>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, b2, filt, as_stream(stream)); break;
        case 64: rc = launch_filter_fwd<64>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, b2, filt, as_stream(stream)); break;
        case 128: rc = launch_filter_fwd<128>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, b2, filt, as_stream(stream)); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    if (rc) { set_error("%s: %s", __func__, cudaGetErrorString((cudaError_t)rc)); return rc; }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int64_t geossl_filter_bwd_workspace(int G, int F) {
    (void)G;
    switch (F) {
        case 32: return (int64_t)kNumSM * Partial<32>::kFloats;
        case 64: return (int64_t)kNumSM * Partial<64>::kFloats;
        case 128: return (int64_t)kNumSM * Partial<128>::kFloats;
        default: return -1;
    }
}

int geossl_filter_bwd(const float* edge_dist, const int32_t* n_edges_dev, int64_t capacity,
                      const float* offset, float coeff, float cutoff, int G, int F,
                      const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* x, const float* grad_out, const int32_t* src, const int32_t* edge_tgt,
                      const float* grad_filt, float* workspace,
                      float* gw1, float* gb1, float* gw2, float* gb2, void* stream) {
    (void)b2;
    GEOSSL_REQUIRE(edge_dist && n_edges_dev && offset && w1 && b1 && w2 && workspace && gw1 && gb1 && gw2 && gb2, "null pointer");
    GEOSSL_REQUIRE(grad_filt || (x && grad_out && src && edge_tgt), "need grad_filt or x/grad_out/src/edge_tgt");
    GEOSSL_REQUIRE(G >= 1 && G <= 64, "num_gaussians must be in [1,64]");
    int rc;
    switch (F) {
        case 32: rc = launch_filter_bwd<32>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, x, grad_out, src, edge_tgt, grad_filt, workspace, gw1, gb1, gw2, gb2, as_stream(stream)); break;
        case 64: rc = launch_filter_bwd<64>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, x, grad_out, src, edge_tgt, grad_filt, workspace, gw1, gb1, gw2, gb2, as_stream(stream)); break;
        case 128: rc = launch_filter_bwd<128>(edge_dist, n_edges_dev, capacity, offset, coeff, cutoff, G, w1, b1, w2, x, grad_out, src, edge_tgt, grad_filt, workspace, gw1, gb1, gw2, gb2, as_stream(stream)); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    if (rc) { set_error("%s: %s", __func__, cudaGetErrorString((cudaError_t)rc)); return rc; }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
