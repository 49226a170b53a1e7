// PaiNN scalar/vector message block (and its backward) on the precomputed radius_edge_index.
//
// Replaces, per interaction, painn.py:53-64 of the reference (gather x[idx_j], mu[idx_j]; Wij*xj; split;
// two index_add scatters over idx_i) together with the per-edge part of PaiNN.forward
// (painn.py:232-245: r_ij, d_ij, dir_ij, GaussianRBF, CosineCutoff, filter_net * fcut).  The (E,1,3F*n_int)
// filter tensor (4.6 KB/edge at F=128) is never materialised: each warp rebuilds the 3F filter values of
// an edge from its 20 rbf values against the filter_net slice held in shared memory.
//
// forward  : one warp per centre atom i (edges grouped by idx_i), lanes own channels {lane + 32 j};
//            q/mu updates accumulate in registers in edge order and are written once (atomic free).
// backward : one warp per neighbour atom j (edges grouped by idx_j) produces dL/dctx[j], dL/dmu[j] and the
//            per-edge filter gradient (E,3F); a persistent tile kernel contracts that with the rbf values
//            into dL/dfilter_net (3F,R) with per-CTA register accumulators and a fixed-order reduction.
#include "common.cuh"

namespace geossl {

constexpr int kMaxRbf = 32;

// phi_pad (E,128), optional: [rbf_0 .. rbf_{R-1}, 1, 0 ...] -- the K-padded operand of the tensor-core filter GEMM
// (column R carries the bias of filter_net); one warp-coalesced 512-byte row per edge.
__global__ void painn_rbf_pad_kernel(const float* __restrict__ dist, const float* __restrict__ fcut, int64_t n_edges,
                                     const float* __restrict__ offsets, const float* __restrict__ widths, int R, int ld,
                                     float* __restrict__ phi_pad) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e = idx / ld;
    const int c = (int)(idx - e * ld);
    if (e >= n_edges) return;
    float v = 0.f;
    if (c < R) {
        const float wdt = __ldg(widths + c);
        const float coeff = -0.5f / __fmul_rn(wdt, wdt);                     // painn_utils.py:100
        const float diff = __ldg(dist + e) - __ldg(offsets + c);
        v = expf(__fmul_rn(coeff, __fmul_rn(diff, diff)));
    } else if (c == R) {
        v = 1.f;
    }
    phi_pad[idx] = v;
}

__global__ void painn_edge_geom_kernel(const float* __restrict__ pos, const int64_t* __restrict__ rei, int64_t n_edges,
                                       int64_t n_atoms, float cutoff, float* __restrict__ dist, float* __restrict__ dir,
                                       float* __restrict__ fcut) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t i = rei[e], j = rei[n_edges + e];
    if (i < 0 || j < 0 || i >= n_atoms || j >= n_atoms) {        // padding column of a capacity-padded list (idx_j = n_atoms)
        dist[e] = 1.f; dir[3 * e] = 0.f; dir[3 * e + 1] = 0.f; dir[3 * e + 2] = 0.f; fcut[e] = 0.f;
        return;
    }
    const float rx = __fsub_rn(pos[3 * i], pos[3 * j]), ry = __fsub_rn(pos[3 * i + 1], pos[3 * j + 1]),
                rz = __fsub_rn(pos[3 * i + 2], pos[3 * j + 2]);                              // painn.py:232
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz)));
    dist[e] = d;
    dir[3 * e] = rx / d; dir[3 * e + 1] = ry / d; dir[3 * e + 2] = rz / d;                    // painn.py:237
    fcut[e] = (d < cutoff) ? cosine_cutoff(d, cutoff) : 0.f;                                  // painn_utils.py:152-155
}

// 3F filter values of one edge for this lane's channels: w[b][j] = (sum_r phi_r W[b*F + lane + 32 j][r] + bias) * fcut
template <int CPL, int F>
__device__ __forceinline__ void edge_filter(float (&w)[3][CPL], const float* __restrict__ sW, const float* __restrict__ sB,
                                            float phi_lane, int R, float fc, int lane) {
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int j = 0; j < CPL; ++j) w[b][j] = sB[b * F + lane + 32 * j];
    for (int r = 0; r < R; ++r) {
        const float pr = __shfl_sync(0xffffffffu, phi_lane, r);
        const float* wr = sW + r * 3 * F;
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int j = 0; j < CPL; ++j) w[b][j] = fmaf(pr, wr[b * F + lane + 32 * j], w[b][j]);
    }
#pragma unroll
    for (int b = 0; b < 3; ++b)
#pragma unroll
        for (int j = 0; j < CPL; ++j) w[b][j] *= fc;
}

template <int F>
__device__ __forceinline__ void load_filter_slice(const float* __restrict__ wf, const float* __restrict__ bf, int R,
                                                  float* sW, float* sB) {
    // sW[r][c] = wf[c][r]  (wf is the (3F,R) slice of filter_net.weight)
    for (int idx = threadIdx.x; idx < 3 * F * R; idx += blockDim.x) {
        const int r = idx / (3 * F), c = idx % (3 * F);
        sW[idx] = __ldg(wf + c * R + r);
    }
    for (int c = threadIdx.x; c < 3 * F; c += blockDim.x) sB[c] = __ldg(bf + c);
}

__device__ __forceinline__ float rbf_lane(float d, const float* __restrict__ offsets, const float* __restrict__ widths, int R, int lane) {
    if (lane >= R) return 0.f;
    const float wdt = __ldg(widths + lane);
    const float coeff = -0.5f / __fmul_rn(wdt, wdt);                  // -0.5 / pow(widths, 2), painn_utils.py:100
    const float diff = d - __ldg(offsets + lane);
    return expf(__fmul_rn(coeff, __fmul_rn(diff, diff)));
}

template <int F>
__global__ void __launch_bounds__(256)
painn_message_fwd_kernel(const float* __restrict__ q, const float* __restrict__ mu, const float* __restrict__ x,
                         const float* __restrict__ wf, const float* __restrict__ bf, const float* __restrict__ offsets,
                         const float* __restrict__ widths, int R,
                         const float* __restrict__ dist, const float* __restrict__ dir, const float* __restrict__ fcut,
                         const int32_t* __restrict__ i_rowptr, const int32_t* __restrict__ i_eid, const int32_t* __restrict__ i_nbr,
                         int n_atoms, float* __restrict__ q_out, float* __restrict__ mu_out, const float* __restrict__ wpre) {
    // wpre != NULL: the pre-cutoff filter rows (E,3F) were materialised by the tensor-core filter GEMM; each warp streams
    // its edge's 3F values (coalesced 128-byte segments) instead of rebuilding them from 20 rbf values against shared memory
    constexpr int CPL = F / 32;
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* sB = smem + 3 * F * R;
    if (wpre == nullptr) {
        load_filter_slice<F>(wf, bf, R, sW, sB);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    if constexpr (CPL == 4) {
        if (wpre != nullptr) {
            // F = 128 with materialised filter rows: lane owns four CONSECUTIVE features (128-bit loads: 9 per edge instead of
            // 36); the edge records of a row (edge id, neighbour, cutoff, direction) are fetched by the lanes in parallel and
            // broadcast by shuffle, so the row loads of consecutive edges do not wait on per-edge index loads
            const int f4 = lane * 4;
            for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_atoms; i += warps_per_grid) {
                float4 aq = make_float4(0.f, 0.f, 0.f, 0.f), a0 = aq, a1 = aq, a2 = aq;
                const int kb = __ldg(i_rowptr + i), ke = __ldg(i_rowptr + i + 1);
                for (int base = kb; base < ke; base += 32) {
                    const int cnt = min(32, ke - base);
                    int my_e = 0, my_nb = 0;
                    float my_fc = 0.f, my_dx = 0.f, my_dy = 0.f, my_dz = 0.f;
                    if (lane < cnt) {
                        my_e = __ldg(i_eid + base + lane);
                        my_nb = __ldg(i_nbr + base + lane);
                        my_fc = __ldg(fcut + my_e);
                        my_dx = __ldg(dir + 3 * (int64_t)my_e); my_dy = __ldg(dir + 3 * (int64_t)my_e + 1); my_dz = __ldg(dir + 3 * (int64_t)my_e + 2);
                    }
#pragma unroll 2
                    for (int k = 0; k < cnt; ++k) {
                        const int e = __shfl_sync(0xffffffffu, my_e, k), nb = __shfl_sync(0xffffffffu, my_nb, k);
                        const float fc = __shfl_sync(0xffffffffu, my_fc, k);
                        const float dx = __shfl_sync(0xffffffffu, my_dx, k), dy = __shfl_sync(0xffffffffu, my_dy, k),
                                    dz = __shfl_sync(0xffffffffu, my_dz, k);
                        const float* wr = wpre + (int64_t)e * 3 * F + f4;
                        const float* xj = x + (int64_t)nb * 3 * F + f4;
                        const float* mj = mu + (int64_t)nb * 3 * F + f4;
                        const float4 w0 = ldg4(wr), w1 = ldg4(wr + F), w2 = ldg4(wr + 2 * F);
                        const float4 x0 = ldg4(xj), x1 = ldg4(xj + F), x2 = ldg4(xj + 2 * F);
                        const float4 m0 = ldg4(mj), m1 = ldg4(mj + F), m2 = ldg4(mj + 2 * F);
#define GEOSSL_MSG_LANE(c)                                                                                   \
                        {                                                                                    \
                            aq.c = fmaf(w0.c * fc, x0.c, aq.c);                                               \
                            const float dmuR = (w1.c * fc) * x1.c, dmumu = (w2.c * fc) * x2.c;                \
                            a0.c += dmuR * dx + dmumu * m0.c;                                                 \
                            a1.c += dmuR * dy + dmumu * m1.c;                                                 \
                            a2.c += dmuR * dz + dmumu * m2.c;                                                 \
                        }
                        GEOSSL_MSG_LANE(x) GEOSSL_MSG_LANE(y) GEOSSL_MSG_LANE(z) GEOSSL_MSG_LANE(w)
#undef GEOSSL_MSG_LANE
                    }
                }
                const float4 q4 = ldg4(q + (int64_t)i * F + f4);
                *reinterpret_cast<float4*>(q_out + (int64_t)i * F + f4) = make_float4(q4.x + aq.x, q4.y + aq.y, q4.z + aq.z, q4.w + aq.w);
                const float* mi = mu + (int64_t)i * 3 * F + f4;
                float* mo = mu_out + (int64_t)i * 3 * F + f4;
                const float4 u0 = ldg4(mi), u1 = ldg4(mi + F), u2 = ldg4(mi + 2 * F);
                *reinterpret_cast<float4*>(mo) = make_float4(u0.x + a0.x, u0.y + a0.y, u0.z + a0.z, u0.w + a0.w);
                *reinterpret_cast<float4*>(mo + F) = make_float4(u1.x + a1.x, u1.y + a1.y, u1.z + a1.z, u1.w + a1.w);
                *reinterpret_cast<float4*>(mo + 2 * F) = make_float4(u2.x + a2.x, u2.y + a2.y, u2.z + a2.z, u2.w + a2.w);
            }
            return;
        }
    }
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_atoms; i += warps_per_grid) {
        float aq[CPL], amu[3][CPL];
#pragma unroll
        for (int j = 0; j < CPL; ++j) { aq[j] = 0.f; amu[0][j] = 0.f; amu[1][j] = 0.f; amu[2][j] = 0.f; }
        const int kb = __ldg(i_rowptr + i), ke = __ldg(i_rowptr + i + 1);
        for (int k = kb; k < ke; ++k) {
            const int e = __ldg(i_eid + k), nb = __ldg(i_nbr + k);
            const float d = __ldg(dist + e), fc = __ldg(fcut + e);
            const float dx = __ldg(dir + 3 * (int64_t)e), dy = __ldg(dir + 3 * (int64_t)e + 1), dz = __ldg(dir + 3 * (int64_t)e + 2);
            float w[3][CPL];
            if (wpre != nullptr) {
                const float* wr = wpre + (int64_t)e * 3 * F;
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                    for (int j = 0; j < CPL; ++j) w[b][j] = __ldg(wr + b * F + lane + 32 * j) * fc;
            } else {
                edge_filter<CPL, F>(w, sW, sB, rbf_lane(d, offsets, widths, R, lane), R, fc, lane);
            }
            const float* xj = x + (int64_t)nb * 3 * F;
            const float* mj = mu + (int64_t)nb * 3 * F;
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int f = lane + 32 * j;
                aq[j] = fmaf(w[0][j], __ldg(xj + f), aq[j]);
                const float dmuR = w[1][j] * __ldg(xj + F + f);
                const float dmumu = w[2][j] * __ldg(xj + 2 * F + f);
                amu[0][j] += dmuR * dx + dmumu * __ldg(mj + f);
                amu[1][j] += dmuR * dy + dmumu * __ldg(mj + F + f);
                amu[2][j] += dmuR * dz + dmumu * __ldg(mj + 2 * F + f);
            }
        }
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
            const int f = lane + 32 * j;
            q_out[(int64_t)i * F + f] = __ldg(q + (int64_t)i * F + f) + aq[j];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                mu_out[(int64_t)i * 3 * F + c * F + f] = __ldg(mu + (int64_t)i * 3 * F + c * F + f) + amu[c][j];
        }
    }
}

template <int F>
__global__ void __launch_bounds__(256)
painn_message_bwd_kernel(const float* __restrict__ gq_out, const float* __restrict__ gmu_out,
                         const float* __restrict__ mu, const float* __restrict__ x,
                         const float* __restrict__ wf, const float* __restrict__ bf, const float* __restrict__ offsets,
                         const float* __restrict__ widths, int R,
                         const float* __restrict__ dist, const float* __restrict__ dir, const float* __restrict__ fcut,
                         const int32_t* __restrict__ j_rowptr, const int32_t* __restrict__ j_ctr,
                         int n_atoms, float* __restrict__ gx, float* __restrict__ gmu_in, float* __restrict__ gfilt,
                         const float* __restrict__ wpre, int64_t zero_tail_cap) {
    constexpr int CPL = F / 32;
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;
    float* sB = smem + 3 * F * R;
    if (wpre == nullptr) {
        load_filter_slice<F>(wf, bf, R, sW, sB);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    bool fast = false;
    if constexpr (CPL == 4) fast = wpre != nullptr;
    if (fast) {
        // same layout as the forward fast path: four consecutive features per lane, edge records broadcast by shuffle
        const int f4 = lane * 4;
        for (int jn = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; jn < n_atoms; jn += warps_per_grid) {
            const float* xp = x + (int64_t)jn * 3 * F + f4;
            const float* mp = mu + (int64_t)jn * 3 * F + f4;
            const float4 x0 = ldg4(xp), x1 = ldg4(xp + F), x2 = ldg4(xp + 2 * F);
            const float4 m0 = ldg4(mp), m1 = ldg4(mp + F), m2 = ldg4(mp + 2 * F);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float4 gx0 = z4, gx1 = z4, gx2 = z4, gm0 = z4, gm1 = z4, gm2 = z4;
            const int eb = __ldg(j_rowptr + jn), ee = __ldg(j_rowptr + jn + 1);
            for (int base = eb; base < ee; base += 32) {
                const int cnt = min(32, ee - base);
                int my_i = 0;
                float my_fc = 0.f, my_dx = 0.f, my_dy = 0.f, my_dz = 0.f;
                if (lane < cnt) {
                    const int64_t e = base + lane;
                    my_i = __ldg(j_ctr + e);
                    my_fc = __ldg(fcut + e);
                    my_dx = __ldg(dir + 3 * e); my_dy = __ldg(dir + 3 * e + 1); my_dz = __ldg(dir + 3 * e + 2);
                }
#pragma unroll 2
                for (int k = 0; k < cnt; ++k) {
                    const int64_t e = base + k;
                    const int i = __shfl_sync(0xffffffffu, my_i, k);
                    const float fc = __shfl_sync(0xffffffffu, my_fc, k);
                    const float dx = __shfl_sync(0xffffffffu, my_dx, k), dy = __shfl_sync(0xffffffffu, my_dy, k),
                                dz = __shfl_sync(0xffffffffu, my_dz, k);
                    const float* wr = wpre + e * 3 * F + f4;
                    const float4 w0 = ldg4(wr), w1 = ldg4(wr + F), w2 = ldg4(wr + 2 * F);
                    const float4 gq = ldg4(gq_out + (int64_t)i * F + f4);
                    const float* gp = gmu_out + (int64_t)i * 3 * F + f4;
                    const float4 g0 = ldg4(gp), g1 = ldg4(gp + F), g2 = ldg4(gp + 2 * F);
                    float4 o0, o1, o2;
#define GEOSSL_MSG_LANE(c)                                                                                       \
                    {                                                                                            \
                        const float br = g0.c * dx + g1.c * dy + g2.c * dz;                                       \
                        const float bm = g0.c * m0.c + g1.c * m1.c + g2.c * m2.c;                                 \
                        const float wa = w0.c * fc, wb = w1.c * fc, wc = w2.c * fc;                               \
                        gx0.c = fmaf(wa, gq.c, gx0.c);                                                            \
                        gx1.c = fmaf(wb, br, gx1.c);                                                              \
                        gx2.c = fmaf(wc, bm, gx2.c);                                                              \
                        const float dmumu = wc * x2.c;                                                            \
                        gm0.c = fmaf(dmumu, g0.c, gm0.c);                                                         \
                        gm1.c = fmaf(dmumu, g1.c, gm1.c);                                                         \
                        gm2.c = fmaf(dmumu, g2.c, gm2.c);                                                         \
                        o0.c = x0.c * gq.c * fc; o1.c = x1.c * br * fc; o2.c = x2.c * bm * fc;                    \
                    }
                    GEOSSL_MSG_LANE(x) GEOSSL_MSG_LANE(y) GEOSSL_MSG_LANE(z) GEOSSL_MSG_LANE(w)
#undef GEOSSL_MSG_LANE
                    float* go = gfilt + e * 3 * F + f4;
                    *reinterpret_cast<float4*>(go) = o0;
                    *reinterpret_cast<float4*>(go + F) = o1;
                    *reinterpret_cast<float4*>(go + 2 * F) = o2;
                }
            }
            const int64_t o = (int64_t)jn * 3 * F + f4;
            *reinterpret_cast<float4*>(gx + o) = gx0;
            *reinterpret_cast<float4*>(gx + o + F) = gx1;
            *reinterpret_cast<float4*>(gx + o + 2 * F) = gx2;
            const float4 h0 = ldg4(gmu_out + o), h1 = ldg4(gmu_out + o + F), h2 = ldg4(gmu_out + o + 2 * F);
            *reinterpret_cast<float4*>(gmu_in + o) = make_float4(h0.x + gm0.x, h0.y + gm0.y, h0.z + gm0.z, h0.w + gm0.w);
            *reinterpret_cast<float4*>(gmu_in + o + F) = make_float4(h1.x + gm1.x, h1.y + gm1.y, h1.z + gm1.z, h1.w + gm1.w);
            *reinterpret_cast<float4*>(gmu_in + o + 2 * F) = make_float4(h2.x + gm2.x, h2.y + gm2.y, h2.z + gm2.z, h2.w + gm2.w);
        }
    }
    for (int jn = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; jn < n_atoms && !fast; jn += warps_per_grid) {
        float xj[3][CPL], mj[3][CPL], agx[3][CPL], agm[3][CPL];
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                xj[b][j] = __ldg(x + (int64_t)jn * 3 * F + b * F + lane + 32 * j);
                mj[b][j] = __ldg(mu + (int64_t)jn * 3 * F + b * F + lane + 32 * j);
                agx[b][j] = 0.f;
                agm[b][j] = 0.f;
            }
        const int eb = __ldg(j_rowptr + jn), ee = __ldg(j_rowptr + jn + 1);
        for (int e = eb; e < ee; ++e) {      // edges with idx_j == jn are contiguous (radius_edge_index is idx_j sorted)
            const int i = __ldg(j_ctr + e);
            const float d = __ldg(dist + e), fc = __ldg(fcut + e);
            const float dx = __ldg(dir + 3 * (int64_t)e), dy = __ldg(dir + 3 * (int64_t)e + 1), dz = __ldg(dir + 3 * (int64_t)e + 2);
            float w[3][CPL];
            if (wpre != nullptr) {
                const float* wr = wpre + (int64_t)e * 3 * F;
#pragma unroll
                for (int b = 0; b < 3; ++b)
#pragma unroll
                    for (int j = 0; j < CPL; ++j) w[b][j] = __ldg(wr + b * F + lane + 32 * j) * fc;
            } else {
                edge_filter<CPL, F>(w, sW, sB, rbf_lane(d, offsets, widths, R, lane), R, fc, lane);
            }
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int f = lane + 32 * j;
                const float gqi = __ldg(gq_out + (int64_t)i * F + f);
                const float g0 = __ldg(gmu_out + (int64_t)i * 3 * F + f), g1 = __ldg(gmu_out + (int64_t)i * 3 * F + F + f),
                            g2 = __ldg(gmu_out + (int64_t)i * 3 * F + 2 * F + f);
                const float br = g0 * dx + g1 * dy + g2 * dz;                       // d/d(dmuR)
                const float bm = g0 * mj[0][j] + g1 * mj[1][j] + g2 * mj[2][j];     // d/d(dmumu)
                agx[0][j] = fmaf(w[0][j], gqi, agx[0][j]);
                agx[1][j] = fmaf(w[1][j], br, agx[1][j]);
                agx[2][j] = fmaf(w[2][j], bm, agx[2][j]);
                const float dmumu = w[2][j] * xj[2][j];
                agm[0][j] = fmaf(dmumu, g0, agm[0][j]);
                agm[1][j] = fmaf(dmumu, g1, agm[1][j]);
                agm[2][j] = fmaf(dmumu, g2, agm[2][j]);
                // gradient w.r.t. the pre-cutoff filter value (phi W^T + b): x_j * bracket * fcut
                gfilt[(int64_t)e * 3 * F + f] = xj[0][j] * gqi * fc;
                gfilt[(int64_t)e * 3 * F + F + f] = xj[1][j] * br * fc;
                gfilt[(int64_t)e * 3 * F + 2 * F + f] = xj[2][j] * bm * fc;
            }
        }
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int j = 0; j < CPL; ++j) {
                const int64_t o = (int64_t)jn * 3 * F + b * F + lane + 32 * j;
                gx[o] = agx[b][j];
                gmu_in[o] = __ldg(gmu_out + o) + agm[b][j];
            }
    }
    // capacity-padded edge lists: rows [live, n_edges_cap) of the per-edge gradient are padding and never written above; the
    // filter GEMM's weight-gradient kernel contracts over ALL rows, so they must be zeros (cheaper here than a full memset)
    if (zero_tail_cap > 0) {
        const int64_t live = (int64_t)__ldg(j_rowptr + n_atoms);
        const int64_t first = live * 3 * F, last = zero_tail_cap * 3 * F;
        for (int64_t o = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < last; o += (int64_t)gridDim.x * blockDim.x) gfilt[o] = 0.f;
    }
}

// dW[c][r] = sum_e gfilt[e][c] * phi_e[r],  db[c] = sum_e gfilt[e][c]; per-CTA partials in workspace
// layout [c][R+1] (column R = bias).  256 threads: thread owns channels {tid % 128 + 128 b} x half of the rbf range.
template <int F>
__global__ void __launch_bounds__(256)
painn_filter_wgrad_kernel(const float* __restrict__ gfilt, const float* __restrict__ dist, int64_t n_edges_cap,
                          const int32_t* __restrict__ n_edges_live, const float* __restrict__ offsets,
                          const float* __restrict__ widths, int R, float* __restrict__ workspace) {
    // live edge count on the device (= rowptr[n_atoms] of the idx_j-sorted CSR): rows past it are padding / never written
    const int64_t n_edges = min((int64_t)__ldg(n_edges_live), n_edges_cap);
    constexpr int C3 = 3 * F;
    constexpr int NC = (C3 + 127) / 128;            // channels per thread (3 for F=128, 2 for F=64, 1 for F=32)
    constexpr int TE = 16;
    __shared__ float sPhi[TE][kMaxRbf + 1];
    __shared__ float sG[TE][C3 + 1];
    const int tid = threadIdx.x;
    const int c0 = tid % 128, half = tid / 128;
    const int RH = (kMaxRbf + 1) / 2 + 1;           // 17 slots per half: rbf 0..16 | 17..32 (slot R = bias column)
    float acc[NC][17];
#pragma unroll
    for (int b = 0; b < NC; ++b)
#pragma unroll
        for (int r = 0; r < 17; ++r) acc[b][r] = 0.f;
    (void)RH;
    const int64_t n_tiles = (n_edges + TE - 1) / TE;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t e0 = t * TE;
        for (int idx = tid; idx < TE * (kMaxRbf + 1); idx += 256) {
            const int el = idx / (kMaxRbf + 1), r = idx % (kMaxRbf + 1);
            const int64_t e = e0 + el;
            float v = 0.f;
            if (e < n_edges) {
                if (r < R) {
                    const float wdt = __ldg(widths + r);
                    const float coeff = -0.5f / __fmul_rn(wdt, wdt);
                    const float diff = __ldg(dist + e) - __ldg(offsets + r);
                    v = expf(__fmul_rn(coeff, __fmul_rn(diff, diff)));
                } else if (r == R) {
                    v = 1.f;                         // bias column
                }
            }
            sPhi[el][r] = v;
        }
        for (int idx = tid; idx < TE * C3; idx += 256) {
            const int el = idx / C3, c = idx % C3;
            const int64_t e = e0 + el;
            sG[el][c] = (e < n_edges) ? __ldg(gfilt + e * C3 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll 4
        for (int el = 0; el < TE; ++el) {
            float g[NC];
#pragma unroll
            for (int b = 0; b < NC; ++b) g[b] = (c0 + 128 * b < C3) ? sG[el][c0 + 128 * b] : 0.f;
#pragma unroll
            for (int r = 0; r < 17; ++r) {
                const int rr = half * 17 + r;
                const float p = (rr <= kMaxRbf) ? sPhi[el][rr] : 0.f;
#pragma unroll
                for (int b = 0; b < NC; ++b) acc[b][r] = fmaf(g[b], p, acc[b][r]);
            }
        }
        __syncthreads();
    }
    float* ws = workspace + (int64_t)blockIdx.x * C3 * (R + 1);
#pragma unroll
    for (int b = 0; b < NC; ++b) {
        const int c = c0 + 128 * b;
        if (c < C3) {
#pragma unroll
            for (int r = 0; r < 17; ++r) {
                const int rr = half * 17 + r;
                if (rr <= R) ws[c * (R + 1) + rr] = acc[b][r];
            }
        }
    }
}

__global__ void painn_filter_wgrad_reduce_kernel(const float* __restrict__ workspace, int n_parts, int C3, int R,
                                                 float* __restrict__ gw, float* __restrict__ gb) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C3 * (R + 1)) return;
    float s = 0.f;
    for (int p = 0; p < n_parts; ++p) s += workspace[(int64_t)p * C3 * (R + 1) + idx];
    const int c = idx / (R + 1), r = idx % (R + 1);
    if (r < R) gw[c * R + r] = s; else gb[c] = s;
}

template <int F>
int launch_msg_fwd(const float* q, const float* mu, const float* x, const float* wf, const float* bf, const float* offsets,
                   const float* widths, int R, const float* dist, const float* dir, const float* fcut,
                   const int32_t* i_rowptr, const int32_t* i_eid, const int32_t* i_nbr, int64_t n_atoms,
                   float* q_out, float* mu_out, const float* wpre, cudaStream_t st) {
    const size_t smem = wpre ? 0 : (size_t)(3 * F * R + 3 * F) * sizeof(float);
    static PerDeviceFlag configured;
    if (!configured.get() && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(painn_message_fwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    int64_t blocks = (n_atoms + 7) / 8;
    if (blocks > kNumSM * 4) blocks = kNumSM * 4;
    painn_message_fwd_kernel<F><<<(int)blocks, 256, smem, st>>>(q, mu, x, wf, bf, offsets, widths, R, dist, dir, fcut,
                                                                 i_rowptr, i_eid, i_nbr, (int)n_atoms, q_out, mu_out, wpre);
    return 0;
}

template <int F>
int launch_msg_bwd(const float* gq_out, const float* gmu_out, const float* mu, const float* x, const float* wf,
                   const float* bf, const float* offsets, const float* widths, int R, const float* dist, const float* dir,
                   const float* fcut, const int32_t* j_rowptr, const int32_t* j_ctr, int64_t n_atoms, int64_t n_edges,
                   float* gx, float* gmu_in, float* gfilt, float* workspace, float* gw, float* gb, const float* wpre, cudaStream_t st) {
    const size_t smem = wpre ? 0 : (size_t)(3 * F * R + 3 * F) * sizeof(float);
    static PerDeviceFlag configured;
    if (!configured.get() && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(painn_message_bwd_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return (int)e;
        configured.set();
    }
    int64_t blocks = (n_atoms + 7) / 8;
    if (blocks > kNumSM * 4) blocks = kNumSM * 4;
    painn_message_bwd_kernel<F><<<(int)blocks, 256, smem, st>>>(gq_out, gmu_out, mu, x, wf, bf, offsets, widths, R, dist, dir,
                                                                 fcut, j_rowptr, j_ctr, (int)n_atoms, gx, gmu_in, gfilt, wpre, wpre ? n_edges : (int64_t)0);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    if (wpre != nullptr) return 0;     // gfilt IS the gradient of the materialised filter: its GEMM's weight-gradient kernel takes over
    count_launch();
    painn_filter_wgrad_kernel<F><<<kNumSM, 256, 0, st>>>(gfilt, dist, n_edges, j_rowptr + n_atoms, offsets, widths, R, workspace);
    e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    count_launch();
    const int n = 3 * F * (R + 1);
    painn_filter_wgrad_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(workspace, kNumSM, 3 * F, R, gw, gb);
    return 0;
}

}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_painn_edge_geometry(const float* pos, const int64_t* radius_edge_index, int64_t n_edges, int64_t n_atoms, float cutoff,
                               float* dist, float* dir, float* fcut, void* stream) {
    if (n_edges == 0) return 0;
    GEOSSL_REQUIRE(pos && radius_edge_index && dist && dir && fcut && n_edges > 0, "null pointer");
    painn_edge_geom_kernel<<<(int)((n_edges + 255) / 256), 256, 0, as_stream(stream)>>>(pos, radius_edge_index, n_edges, n_atoms,
                                                                                        cutoff, dist, dir, fcut);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_painn_rbf_pad(const float* dist, const float* fcut, int64_t n_edges, const float* offsets, const float* widths, int n_rbf,
                         int ld, float* phi_pad, void* stream) {
    if (n_edges == 0) return 0;
    GEOSSL_REQUIRE(dist && fcut && offsets && widths && phi_pad, "null pointer");
    GEOSSL_REQUIRE((ld == 32 || ld == 64 || ld == 128) && n_rbf >= 1 && n_rbf < ld, "ld must be 32/64/128 and larger than n_rbf");
    const int64_t n = n_edges * ld;
    painn_rbf_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(dist, fcut, n_edges, offsets, widths, n_rbf, ld, phi_pad);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_painn_message_fwd(const float* q, const float* mu, const float* ctx, const float* filter_w, const float* filter_b,
                             const float* offsets, const float* widths, int n_rbf, int F,
                             const float* dist, const float* dir, const float* fcut,
                             const int32_t* i_rowptr, const int32_t* i_eid, const int32_t* i_nbr, int64_t n_atoms,
                             float* q_out, float* mu_out, const float* filter_pre, void* stream) {
    if (n_atoms == 0) return 0;
    GEOSSL_REQUIRE(q && mu && ctx && offsets && widths && i_rowptr && q_out && mu_out, "null pointer");
    GEOSSL_REQUIRE(filter_pre || (filter_w && filter_b), "either the materialised filter or the filter_net slice is required");
    GEOSSL_REQUIRE(n_rbf >= 1 && n_rbf <= kMaxRbf, "n_rbf must be in [1,32]");
    int rc;
    switch (F) {
        case 32: rc = launch_msg_fwd<32>(q, mu, ctx, filter_w, filter_b, offsets, widths, n_rbf, dist, dir, fcut, i_rowptr, i_eid, i_nbr, n_atoms, q_out, mu_out, filter_pre, as_stream(stream)); break;
        case 64: rc = launch_msg_fwd<64>(q, mu, ctx, filter_w, filter_b, offsets, widths, n_rbf, dist, dir, fcut, i_rowptr, i_eid, i_nbr, n_atoms, q_out, mu_out, filter_pre, as_stream(stream)); break;
        case 128: rc = launch_msg_fwd<128>(q, mu, ctx, filter_w, filter_b, offsets, widths, n_rbf, dist, dir, fcut, i_rowptr, i_eid, i_nbr, n_atoms, q_out, mu_out, filter_pre, as_stream(stream)); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    if (rc) { set_error("%s: %s", __func__, cudaGetErrorString((cudaError_t)rc)); return rc; }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int64_t geossl_painn_workspace(int n_rbf, int F) { return (int64_t)kNumSM * 3 * F * (n_rbf + 1); }

int geossl_painn_message_bwd(const float* grad_q_out, const float* grad_mu_out, const float* mu, const float* ctx,
                             const float* filter_w, const float* filter_b, const float* offsets, const float* widths,
                             int n_rbf, int F, const float* dist, const float* dir, const float* fcut,
                             const int32_t* j_rowptr, const int32_t* j_ctr, int64_t n_atoms, int64_t n_edges,
                             float* grad_ctx, float* grad_mu_in, float* edge_scratch, float* workspace,
                             float* grad_filter_w, float* grad_filter_b, const float* filter_pre, void* stream) {
    GEOSSL_REQUIRE(grad_q_out && grad_mu_out && mu && ctx && offsets && widths && j_rowptr && grad_ctx && grad_mu_in, "null pointer");
    GEOSSL_REQUIRE(filter_pre || (filter_w && filter_b && workspace && grad_filter_w && grad_filter_b),
                   "either the materialised filter or the filter_net slice (+ its gradient buffers) is required");
    GEOSSL_REQUIRE(n_edges == 0 || (edge_scratch && dist && dir && fcut && j_ctr), "null edge pointer");
    GEOSSL_REQUIRE(n_rbf >= 1 && n_rbf <= kMaxRbf, "n_rbf must be in [1,32]");
    GEOSSL_REQUIRE(n_atoms > 0, "n_atoms must be > 0");
    int rc;
    switch (F) {
        case 32: rc = launch_msg_bwd<32>(grad_q_out, grad_mu_out, mu, ctx, filter_w, filter_b, offsets, widths, n_rbf, dist, dir, fcut, j_rowptr, j_ctr, n_atoms, n_edges, grad_ctx, grad_mu_in, edge_scratch, workspace, grad_filter_w, grad_filter_b, filter_pre, as_stream(stream)); break;
        case 64: rc = launch_msg_bwd<64>(grad_q_out, grad_mu_out, mu, ctx, filter_w, filter_b, offsets, widths, n_rbf, dist, dir, fcut, j_rowptr, j_ctr, n_atoms, n_edges, grad_ctx, grad_mu_in, edge_scratch, workspace, grad_filter_w, grad_filter_b, filter_pre, as_stream(stream)); break;
        case 128: rc = launch_msg_bwd<128>(grad_q_out, grad_mu_out, mu, ctx, filter_w, filter_b, offsets, widths, n_rbf, dist, dir, fcut, j_rowptr, j_ctr, n_atoms, n_edges, grad_ctx, grad_mu_in, edge_scratch, workspace, grad_filter_w, grad_filter_b, filter_pre, as_stream(stream)); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    if (rc) { set_error("%s: %s", __func__, cudaGetErrorString((cudaError_t)rc)); return rc; }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
