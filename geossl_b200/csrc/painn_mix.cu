// PaiNN update ("mixing") block -- the non-GEMM part, fused (painn.py:100-113):
//   pre  : [mu_V | mu_W] = mu_mix;  vn = sqrt(sum_xyz mu_V^2 + eps);  ctx = [q, vn];  dot = sum_xyz mu_V * mu_W
//   post : [a | b | c] = y;  q' = q + a + c * dot;  mu' = mu + b * mu_W
// and their backward kernels.  The three Dense layers around them (mu_channel_mix 128->256 on the 3N vector rows,
// intraatomic_context_net 256->128->384) are library GEMMs.  One thread handles 4 consecutive channels of one atom
// (128-bit loads/stores, fully coalesced); everything is elementwise, so these kernels are HBM bound:
//   pre  reads 4F + 24F B/atom, writes 8F + 4F;  post reads 4F + 12F + 12F + 24F + 4F, writes 4F + 12F.
#include "common.cuh"

namespace geossl {

#define F4(p) (*reinterpret_cast<const float4*>(p))
#define ST4(p, v) (*reinterpret_cast<float4*>(p) = (v))
__device__ __forceinline__ float4 f4mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 f4fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}

__global__ void painn_mix_pre_kernel(const float* __restrict__ q, const float* __restrict__ mm, int64_t n, int F, float eps,
                                     float* __restrict__ ctx, float* __restrict__ dot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = F / 4;
    if (i >= n * per) return;
    const int64_t a = i / per;
    const int f = (int)(i % per) * 4;
    const float* m = mm + a * 6 * F + f;
    const float4 v0 = F4(m), v1 = F4(m + 2 * F), v2 = F4(m + 4 * F);
    const float4 w0 = F4(m + F), w1 = F4(m + 3 * F), w2 = F4(m + 5 * F);
    const float4 ss = f4fma(v2, v2, f4fma(v1, v1, f4mul(v0, v0)));
    ST4(ctx + a * 2 * F + f, F4(q + a * F + f));
    ST4(ctx + a * 2 * F + F + f, make_float4(sqrtf(ss.x + eps), sqrtf(ss.y + eps), sqrtf(ss.z + eps), sqrtf(ss.w + eps)));
    ST4(dot + a * F + f, f4fma(v2, w2, f4fma(v1, w1, f4mul(v0, w0))));
}

// g_mm[V_k] = g_vn * V_k / vn + g_dot * W_k ; g_mm[W_k] = g_dot * V_k ; g_q = g_ctx[:, :F]
__global__ void painn_mix_pre_bwd_kernel(const float* __restrict__ mm, const float* __restrict__ ctx, const float* __restrict__ g_ctx,
                                         const float* __restrict__ g_dot, int64_t n, int F, float* __restrict__ g_q,
                                         float* __restrict__ g_mm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = F / 4;
    if (i >= n * per) return;
    const int64_t a = i / per;
    const int f = (int)(i % per) * 4;
    const float* m = mm + a * 6 * F + f;
    float* gm = g_mm + a * 6 * F + f;
    const float4 vn = F4(ctx + a * 2 * F + F + f), gvn = F4(g_ctx + a * 2 * F + F + f);
    const float4 gd = g_dot ? F4(g_dot + a * F + f) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 s = make_float4(gvn.x / vn.x, gvn.y / vn.y, gvn.z / vn.z, gvn.w / vn.w);
    ST4(g_q + a * F + f, F4(g_ctx + a * 2 * F + f));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 v = F4(m + 2 * k * F), w = F4(m + (2 * k + 1) * F);
        ST4(gm + 2 * k * F, f4fma(s, v, f4mul(gd, w)));
        ST4(gm + (2 * k + 1) * F, f4mul(gd, v));
    }
}

__global__ void painn_mix_post_kernel(const float* __restrict__ q, const float* __restrict__ mu, const float* __restrict__ y,
                                      const float* __restrict__ mm, const float* __restrict__ dot, int64_t n, int F,
                                      float* __restrict__ q_out, float* __restrict__ mu_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = F / 4;
    if (i >= n * per) return;
    const int64_t a = i / per;
    const int f = (int)(i % per) * 4;
    const float4 ya = F4(y + a * 3 * F + f), yb = F4(y + a * 3 * F + F + f), yc = F4(y + a * 3 * F + 2 * F + f);
    ST4(q_out + a * F + f, f4fma(yc, F4(dot + a * F + f), f4add(F4(q + a * F + f), ya)));
#pragma unroll
    for (int k = 0; k < 3; ++k)
        ST4(mu_out + (a * 3 + k) * F + f, f4fma(yb, F4(mm + a * 6 * F + (2 * k + 1) * F + f), F4(mu + (a * 3 + k) * F + f)));
}

// g_y = [g_q', sum_k g_mu'[k] * W_k, g_q' * dot] ; g_mm[W_k] = g_mu'[k] * b (V part 0) ; g_dot = g_q' * c
__global__ void painn_mix_post_bwd_kernel(const float* __restrict__ gq, const float* __restrict__ gmu, const float* __restrict__ y,
                                          const float* __restrict__ mm, const float* __restrict__ dot, int64_t n, int F,
                                          float* __restrict__ g_y, float* __restrict__ g_mm, float* __restrict__ g_dot) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = F / 4;
    if (i >= n * per) return;
    const int64_t a = i / per;
    const int f = (int)(i % per) * 4;
    const float4 g = F4(gq + a * F + f), yb = F4(y + a * 3 * F + F + f), yc = F4(y + a * 3 * F + 2 * F + f);
    float4 gb = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 zero = gb;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 gm = F4(gmu + (a * 3 + k) * F + f);
        gb = f4fma(gm, F4(mm + a * 6 * F + (2 * k + 1) * F + f), gb);
        ST4(g_mm + a * 6 * F + 2 * k * F + f, zero);
        ST4(g_mm + a * 6 * F + (2 * k + 1) * F + f, f4mul(gm, yb));
    }
    ST4(g_y + a * 3 * F + f, g);
    ST4(g_y + a * 3 * F + F + f, gb);
    ST4(g_y + a * 3 * F + 2 * F + f, f4mul(g, F4(dot + a * F + f)));
    ST4(g_dot + a * F + f, f4mul(g, yc));
}

}  // namespace geossl

using namespace geossl;

#define MIX_LAUNCH(kernel, ...)                                                         \
    do {                                                                                \
        if (n_atoms == 0) return 0;                                                     \
        GEOSSL_REQUIRE(F > 0 && F % 4 == 0, "F must be a positive multiple of 4");      \
        const int64_t total = n_atoms * (F / 4);                                        \
        kernel<<<(int)((total + 255) / 256), 256, 0, as_stream(stream)>>>(__VA_ARGS__); \
        GEOSSL_LAUNCH_CHECK();                                                          \
        return 0;                                                                       \
    } while (0)

extern "C" {

int geossl_painn_mix_pre(const float* q, const float* mu_mix, int64_t n_atoms, int F, float epsilon, float* ctx, float* dot, void* stream) {
    GEOSSL_REQUIRE(n_atoms == 0 || (q && mu_mix && ctx && dot), "null pointer");
    MIX_LAUNCH(painn_mix_pre_kernel, q, mu_mix, n_atoms, F, epsilon, ctx, dot);
}
int geossl_painn_mix_pre_bwd(const float* mu_mix, const float* ctx, const float* grad_ctx, const float* grad_dot, int64_t n_atoms, int F,
                             float* grad_q, float* grad_mu_mix, void* stream) {
    GEOSSL_REQUIRE(n_atoms == 0 || (mu_mix && ctx && grad_ctx && grad_q && grad_mu_mix), "null pointer");
    MIX_LAUNCH(painn_mix_pre_bwd_kernel, mu_mix, ctx, grad_ctx, grad_dot, n_atoms, F, grad_q, grad_mu_mix);
}
int geossl_painn_mix_post(const float* q, const float* mu, const float* y, const float* mu_mix, const float* dot, int64_t n_atoms, int F,
                          float* q_out, float* mu_out, void* stream) {
    GEOSSL_REQUIRE(n_atoms == 0 || (q && mu && y && mu_mix && dot && q_out && mu_out), "null pointer");
    MIX_LAUNCH(painn_mix_post_kernel, q, mu, y, mu_mix, dot, n_atoms, F, q_out, mu_out);
}
int geossl_painn_mix_post_bwd(const float* grad_q_out, const float* grad_mu_out, const float* y, const float* mu_mix, const float* dot,
                              int64_t n_atoms, int F, float* grad_y, float* grad_mu_mix, float* grad_dot, void* stream) {
    GEOSSL_REQUIRE(n_atoms == 0 || (grad_q_out && grad_mu_out && y && mu_mix && dot && grad_y && grad_mu_mix && grad_dot), "null pointer");
    MIX_LAUNCH(painn_mix_post_bwd_kernel, grad_q_out, grad_mu_out, y, mu_mix, dot, n_atoms, F, grad_y, grad_mu_mix, grad_dot);
}

}  // extern "C"
