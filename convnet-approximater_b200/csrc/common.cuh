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

// GELU (erf form, nn.GELU's default) on a pair of fp32 values with erf from Abramowitz & Stegun 7.1.28:
//   erf(z) = 1 - (1 + a1 z + ... + a6 z^6)^-16,  z >= 0,  |error| <= 3e-7  (measured against F.gelu: 1.1e-6 absolute in fp32),
// fp32-faithful for a result that is rounded to bf16 next.  Everything but the reciprocal runs as packed f32x2 instructions:
// ~15 FMA-pipe instructions + 2 MUFU per PAIR, where libdevice's erff costs ~35 per value with a divergent branch (the depthwise
// 3x3 + GELU pass of the FFN was 283 us with erff against an 82 us HBM floor).
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
    return (unsigned long long)__float_as_uint(a) | ((unsigned long long)__float_as_uint(b) << 32);
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float2 gelu_erf_fast2(float x0, float x1) {
    const unsigned long long x = f2_pack(x0, x1);
    const unsigned long long z = f2_pack(fabsf(x0) * 0.70710678118654752f, fabsf(x1) * 0.70710678118654752f);
    unsigned long long p = f2_fma(f2_pack(0.0000430638f, 0.0000430638f), z, f2_pack(0.0002765672f, 0.0002765672f));
    p = f2_fma(p, z, f2_pack(0.0001520143f, 0.0001520143f));
    p = f2_fma(p, z, f2_pack(0.0092705272f, 0.0092705272f));
    p = f2_fma(p, z, f2_pack(0.0422820123f, 0.0422820123f));
    p = f2_fma(p, z, f2_pack(0.0705230784f, 0.0705230784f));
    p = f2_fma(p, z, f2_pack(1.f, 1.f));
    unsigned long long r = f2_pack(__fdividef(1.f, __uint_as_float((uint32_t)p)), __fdividef(1.f, __uint_as_float((uint32_t)(p >> 32))));
    r = f2_mul(r, r); r = f2_mul(r, r); r = f2_mul(r, r); r = f2_mul(r, r);                    // ^16
    // 0.5 x (1 + sign(x) (1 - r)) = 0.5 x + 0.5 |x| (1 - r) = 0.5 (x + |x|) - 0.5 |x| r
    const unsigned long long ax = f2_pack(fabsf(x0), fabsf(x1));
    const unsigned long long h = f2_pack(0.5f, 0.5f);
    const unsigned long long s = f2_fma(ax, h, f2_mul(x, h));                                   // 0.5 (x + |x|)
    const unsigned long long g = f2_fma(f2_mul(ax, f2_pack(-0.5f, -0.5f)), r, s);
    return make_float2(__uint_as_float((uint32_t)g), __uint_as_float((uint32_t)(g >> 32)));
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
