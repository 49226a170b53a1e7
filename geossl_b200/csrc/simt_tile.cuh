// Register-tiled fp32 SIMT GEMM helpers over 64-row tiles held k-major in shared memory.
//
// 256 threads; for an output tile of 64 rows (edges / pairs) x N columns: TX = N/4 lanes own columns
// {tx + TX*j, j<4}, TY = 256/TX groups own ME = 64/TY consecutive rows.  Stride-TX column ownership
// keeps weight reads and transposed [column][row] tile stores at the minimum number of shared-memory
// wavefronts; tile rows have stride S = 68 floats (16 B aligned, 4 banks apart).
#pragma once
#include "common.cuh"

namespace geossl {

template <int N>
struct TileCfg {
    static constexpr int TX = N / 4;
    static constexpr int TY = 256 / TX;
    static constexpr int TE = 64;
    static constexpr int ME = TE / TY;      // 8 / 4 / 2 / 1 for N = 128 / 64 / 32 / 16
    static constexpr int S = TE + 4;
};
template <int F>
struct FCfg : TileCfg<F> {
    static constexpr int GP = 64;                        // padded number of gaussians
    static constexpr int MO = F / TileCfg<F>::TY;        // dW2 rows per thread: 16 / 4 / 1
    static constexpr int MG = GP / TileCfg<F>::TY;       // dW1 gaussians per thread: 8 / 4 / 2
};

// acc[i][j] += sum_{k<K} sA[k*S + r0 + i] * sB[k*bks + (tx + TX*j)*bns]       (outputs: rows x columns)
template <int N>
__device__ __forceinline__ void gemm_rows(float (&acc)[TileCfg<N>::ME][4], const float* __restrict__ sA,
                                          const float* __restrict__ sB, int bks, int bns, int K, int tx, int r0) {
    using C = TileCfg<N>;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float a[C::ME], b[4];
        const float* ap = sA + k * C::S + r0;
        if constexpr (C::ME >= 4) {
#pragma unroll
            for (int i = 0; i < C::ME; i += 4) {
                float4 v = *reinterpret_cast<const float4*>(ap + i);
                a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
            }
        } else if constexpr (C::ME == 2) {
            float2 v = *reinterpret_cast<const float2*>(ap);
            a[0] = v.x; a[1] = v.y;
        } else {
            a[0] = ap[0];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sB[k * bks + (tx + C::TX * j) * bns];
#pragma unroll
        for (int i = 0; i < C::ME; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

// acc[i][j] += sum_{r<64} sA[(m0+i)*S + r] * sB[(tx + TX*j)*S + r]   for m0+i < m_rows   (weight-gradient tiles)
template <int N, int MROWS>
__device__ __forceinline__ void gemm_wgrad(float (&acc)[MROWS][4], const float* __restrict__ sA, int m0, int m_rows,
                                           const float* __restrict__ sB, int tx) {
    using C = TileCfg<N>;
    if (m0 >= m_rows) return;
#pragma unroll 2
    for (int k0 = 0; k0 < C::TE; k0 += 4) {
        float4 b[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(sB + (tx + C::TX * j) * C::S + k0);
#pragma unroll
        for (int i = 0; i < MROWS; ++i) {
            const float4 a = *reinterpret_cast<const float4*>(sA + (m0 + i) * C::S + k0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = acc[i][j];
                t = fmaf(a.x, b[j].x, t); t = fmaf(a.y, b[j].y, t); t = fmaf(a.z, b[j].z, t); t = fmaf(a.w, b[j].w, t);
                acc[i][j] = t;
            }
        }
    }
}

template <int R, int Cc>
__device__ __forceinline__ void zero_acc(float (&acc)[R][Cc]) {
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < Cc; ++j) acc[i][j] = 0.f;
}

}  // namespace geossl
