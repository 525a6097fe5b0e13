// Shared helpers for the sm_100a kernels of the ConvNet-Approximater hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/cna_b200.h"

namespace cna {

// ---- error plumbing (thread-local message, integer status over the C ABI) -------------------
void set_error(const std::string& msg);
void count_launch(int n = 1);

#define CNA_FAIL(msg)                                                                   \
    do {                                                                                \
        ::cna::set_error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
        return 1;                                                                       \
    } while (0)

#define CNA_CHECK(cond, msg) \
    do { if (!(cond)) CNA_FAIL(std::string("check failed: " #cond " — ") + (msg)); } while (0)

#define CNA_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) CNA_FAIL(std::string(#expr " -> ") + cudaGetErrorString(_e)); \
    } while (0)

// After a launch: catch configuration errors without synchronising.
#define CNA_LAUNCH_OK()                                                                 \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) CNA_FAIL(std::string("kernel launch: ") + cudaGetErrorString(_e)); \
        ::cna::count_launch();                                                          \
    } while (0)

// Same check without counting (helpers whose caller counts the launch).
#define CNA_LAUNCH_CHECK()                                                              \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) CNA_FAIL(std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    } while (0)

// dwconv_tile.cu: register-tiled depthwise kernels for the MSCA shapes; *handled = false -> use the generic kernel
int dw_tile_try(const void* x, const void* w_packed, const float* bias, const float* res, int res_p, const void* add, void* y,
                int B, int H, int W, int Cin, int mult, int kh, int kw, int sh, int sw, int ph, int pw, int dtype, cudaStream_t st,
                bool* handled);

// dwstrip_tma.cu: TMA-pipelined bf16 strips (1 x K / K x 1) without fused extras; 0 = launched, -1 = not taken, 1 = error
int dw_strip_tma_try(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, int kh, int kw, cudaStream_t st);

int dw_square_tma_try(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, int k, int gelu, int force,
                      cudaStream_t st);

// ---- per-device state ----------------------------------------------------------------------------
// One process may drive several GPUs (one host thread per device or cudaSetDevice round-robin): SM count, the opt-in
// shared-memory limit and every cudaFuncSetAttribute are per DEVICE, so nothing below is cached process-wide.
struct DevInfo { int dev, num_sms, max_smem; };
// Properties of the CURRENT device (cached per ordinal, thread-safe).  Returns 0 / non-zero like the C ABI.
int dev_info(DevInfo* out);

// cudaFuncAttributeMaxDynamicSharedMemorySize, applied once per (kernel instantiation, device): `done` is a per-call-site
// bit mask over device ordinals (function-local static std::atomic<uint64_t>).
template <typename KernelT>
inline cudaError_t ensure_dyn_smem(KernelT kernel, int bytes, std::atomic<unsigned long long>& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
    return e;
}

// Stream-ordered scratch that is released on every return path (the error macros return early).
struct AsyncScratch {
    void* ptr = nullptr;
    cudaStream_t st = nullptr;
    explicit AsyncScratch(cudaStream_t s) : st(s) {}
    AsyncScratch(const AsyncScratch&) = delete;
    AsyncScratch& operator=(const AsyncScratch&) = delete;
    cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&ptr, bytes, st); }
    ~AsyncScratch() { if (ptr) cudaFreeAsync(ptr, st); }
};

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// GELU (erf form, nn.GELU's default) with erf from Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7 + the two MUFU approximations, i.e.
// fp32-faithful for a result that is rounded to bf16 next): 2 MUFU + ~10 FMA-pipe instructions per value, where libdevice's erff is
// ~35 with a divergent branch - the depthwise 3x3 + GELU pass of the FFN was 283 us with erff against an 82 us HBM floor.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float erf_abs = fmaf(-p * t, __expf(-z * z), 1.f);      // erf(|x| / sqrt 2)
    return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// ---- element type helpers ---------------------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ __nv_bfloat16 ldg<__nv_bfloat16>(const __nv_bfloat16* p) {
    unsigned short u = __ldg(reinterpret_cast<const unsigned short*>(p));
    return __ushort_as_bfloat16(u);
}

}  // namespace cna
