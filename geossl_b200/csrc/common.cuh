// Shared helpers of libgeossl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/geossl_b200.h"

namespace geossl {

constexpr int kWarp = 32;
constexpr int kNumSM = 148;            // B200: 2 dies x 74 SMs; persistent grids are sized in multiples of it
constexpr float kPi = 3.14159265358979323846f;
constexpr float kLog2 = 0.6931471824645996f;   // fp32 log(2), ShiftedSoftplus.shift (schnet.py:213)

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a PER-DEVICE property of a kernel: every launcher remembers per
// (call site, device) whether it has opted in, so a process that touches several GPUs configures each of them.
struct PerDeviceFlag {
    bool done[64] = {};
    static int device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < 64) ? d : 0; }
    bool get() const { return done[device()]; }
    void set() { done[device()] = true; }
};

#define GEOSSL_REQUIRE(cond, msg)                                            \
    do {                                                                     \
        if (!(cond)) {                                                       \
            ::geossl::set_error("%s: %s", __func__, msg);                    \
            return GEOSSL_EINVAL;                                            \
        }                                                                    \
    } while (0)

#define GEOSSL_LAUNCH_CHECK()                                                \
    do {                                                                     \
        cudaError_t e__ = cudaGetLastError();                                \
        if (e__ != cudaSuccess) {                                            \
            ::geossl::set_error("%s: %s", __func__, cudaGetErrorString(e__)); \
            return (int)e__;                                                 \
        }                                                                    \
        ::geossl::count_launch();                                            \
    } while (0)

#define GEOSSL_CUDA(call)                                                    \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) {                                            \
            ::geossl::set_error("%s: %s", __func__, cudaGetErrorString(e__)); \
            return (int)e__;                                                 \
        }                                                                    \
    } while (0)

// Programmatic dependent launch: a kernel launched with launch_pdl() may begin while its predecessor on the stream is
// still draining; pdl_wait() blocks until the predecessor has completed and its writes are visible, so everything
// before it (barrier / TMEM set-up, index math) overlaps the predecessor's tail and the launch latency.  EVERY kernel
// launched this way calls pdl_wait() before its first global-memory access (which also makes the ordering transitive).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    // Off by default: measured on the full DDM step it costs 3 % (83.0 k vs 85.6 k molecules/s) although an isolated chain of
    // dense-layer kernels gains 0.85 us per launch -- early-resident dependents take SM slots from the predecessor's tail.
    const char* e = getenv("GEOSSL_PDL");                       // GEOSSL_PDL=1 turns programmatic dependent launch on
    return e && e[0] == '1';
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    static const bool enabled = pdl_enabled();
    cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// torch softplus (beta=1, threshold=20) minus log 2  (schnet.py:215-216)
__device__ __forceinline__ float ssp(float x) {
    return (x > 20.f ? x : log1pf(expf(x))) - kLog2;
}
// d/dx of the above: sigmoid(x), exactly 1 above the threshold
__device__ __forceinline__ float ssp_grad(float x) {
    return x > 20.f ? 1.f : 1.f / (1.f + expf(-x));
}
// 0.5*(cos(d*pi/cutoff)+1) with the reference's op order (schnet.py:186)
__device__ __forceinline__ float cosine_cutoff(float d, float cutoff) {
    return 0.5f * (cosf(__fdiv_rn(__fmul_rn(d, kPi), cutoff)) + 1.0f);
}
__device__ __forceinline__ float cosine_cutoff_grad(float d, float cutoff) {
    return -0.5f * (kPi / cutoff) * sinf(__fdiv_rn(__fmul_rn(d, kPi), cutoff));
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// streaming 128-bit load that does not allocate in L1 (data read exactly once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace geossl
