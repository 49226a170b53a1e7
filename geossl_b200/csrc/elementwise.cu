// Elementwise families that are closed under differentiation, for the composed double-backward path (force training,
// finetune_md17.py:32-54): autograd differentiates the backward pass again, and torch's own derivative formulas of
// softplus / broadcast products expand into chains of 4-8 elementwise launches over (pairs x 128) tensors.
//
//   ssp family    E_k(a, g) = g * s^(k)(a),  s = shifted softplus (schnet.py:215-216): s' = sigmoid, s'' = sigma (1 - sigma),
//                 s''' = sigma (1 - sigma)(1 - 2 sigma);  k = 0 is the activation itself (g ignored).
//                 d/dg E_k(a, g) . v = E_k(a, v),   d/da E_k(a, g) . v = E_{k+1}(a, v * g)
//   row products  RowScale(x, c)[r][f] = x[r][f] * c[r]   (the cutoff product of schnet.py:187),
//                 RowDot(a, b)[r] = sum_f a[r][f] b[r][f]: d RowScale = (RowScale(v, c), RowDot(v, x)),
//                 d RowDot = (RowScale(b, v), RowScale(a, v)).
// fp32 throughout with libdevice expf / log1pf (the same functions as the exact `simt` filter kernels).
#include "common.cuh"

namespace geossl {

__device__ __forceinline__ float ssp_k(float a, int k) {
    if (k == 0) return ssp(a);
    const float sg = a > 20.f ? 1.f : 1.f / (1.f + expf(-a));          // torch: softplus is the identity above threshold 20
    if (k == 1) return sg;
    if (a > 20.f) return 0.f;
    const float d2 = sg * (1.f - sg);
    if (k == 2) return d2;
    return d2 * (1.f - 2.f * sg);
}

__global__ void __launch_bounds__(256)
ssp_family_kernel(const float* __restrict__ a, const float* __restrict__ g, int64_t n, int k, float* __restrict__ out) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        const float4 av = *reinterpret_cast<const float4*>(a + i);
        float4 r = make_float4(ssp_k(av.x, k), ssp_k(av.y, k), ssp_k(av.z, k), ssp_k(av.w, k));
        if (k > 0) {
            const float4 gv = *reinterpret_cast<const float4*>(g + i);
            r.x *= gv.x; r.y *= gv.y; r.z *= gv.z; r.w *= gv.w;
        }
        *reinterpret_cast<float4*>(out + i) = r;
    } else {
        for (int64_t j = i; j < n; ++j) out[j] = ssp_k(a[j], k) * (k > 0 ? g[j] : 1.f);
    }
}

// one warp per row group: LPR lanes cover a row of F floats with float4 each
template <int F>
__global__ void __launch_bounds__(256)
row_scale_kernel(const float* __restrict__ x, const float* __restrict__ c, int64_t n_rows, float* __restrict__ out) {
    constexpr int LPR = F / 4;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = t / LPR;
    if (row >= n_rows) return;
    const int f = (int)(t % LPR) * 4;
    const float s = __ldg(c + row);
    const float4 v = *reinterpret_cast<const float4*>(x + row * F + f);
    *reinterpret_cast<float4*>(out + row * F + f) = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
}

template <int F>
__global__ void __launch_bounds__(256)
row_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n_rows, float* __restrict__ out) {
    constexpr int LPR = F / 4, RPW = 32 / LPR;                         // rows per warp
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t row = warp * RPW + lane / LPR;
    const int f = (lane % LPR) * 4;
    float s = 0.f;
    if (row < n_rows) {
        const float4 av = *reinterpret_cast<const float4*>(a + row * F + f);
        const float4 bv = *reinterpret_cast<const float4*>(b + row * F + f);
        s = fmaf(av.x, bv.x, fmaf(av.y, bv.y, fmaf(av.z, bv.z, av.w * bv.w)));
    }
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // fixed order: deterministic
    if (row < n_rows && (lane % LPR) == 0) out[row] = s;
}

}  // namespace geossl

using namespace geossl;

extern "C" {

int geossl_ssp_family(const float* a, const float* g, int64_t n, int order, float* out, void* stream) {
    if (n == 0) return 0;
    GEOSSL_REQUIRE(a && out && n > 0 && (order == 0 || g), "null pointer");
    GEOSSL_REQUIRE(order >= 0 && order <= 3, "derivative order must be 0..3");
    GEOSSL_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(g)) & 15) == 0,
                   "operands must be 16-byte aligned");
    const int64_t threads = (n + 3) / 4;
    ssp_family_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, as_stream(stream)>>>(a, g, n, order, out);
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_row_scale(const float* x, const float* c, int64_t n_rows, int F, float* out, void* stream) {
    if (n_rows == 0) return 0;
    GEOSSL_REQUIRE(x && c && out && n_rows > 0, "null pointer");
    cudaStream_t st = as_stream(stream);
    switch (F) {
        case 32: row_scale_kernel<32><<<(unsigned)((n_rows * 8 + 255) / 256), 256, 0, st>>>(x, c, n_rows, out); break;
        case 64: row_scale_kernel<64><<<(unsigned)((n_rows * 16 + 255) / 256), 256, 0, st>>>(x, c, n_rows, out); break;
        case 128: row_scale_kernel<128><<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, st>>>(x, c, n_rows, out); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

int geossl_row_dot(const float* a, const float* b, int64_t n_rows, int F, float* out, void* stream) {
    if (n_rows == 0) return 0;
    GEOSSL_REQUIRE(a && b && out && n_rows > 0, "null pointer");
    cudaStream_t st = as_stream(stream);
    switch (F) {
        case 32: row_dot_kernel<32><<<(unsigned)((n_rows * 8 + 255) / 256), 256, 0, st>>>(a, b, n_rows, out); break;
        case 64: row_dot_kernel<64><<<(unsigned)((n_rows * 16 + 255) / 256), 256, 0, st>>>(a, b, n_rows, out); break;
        case 128: row_dot_kernel<128><<<(unsigned)((n_rows * 32 + 255) / 256), 256, 0, st>>>(a, b, n_rows, out); break;
        default: set_error("%s: unsupported width F=%d (32/64/128)", __func__, F); return GEOSSL_EINVAL;
    }
    GEOSSL_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
