// NCHW <-> NHWC boundary conversions (the reference's tensors are logical NCHW;
// every kernel in this library works on NHWC).  Tiled 32x32 shared-memory transposes.
#include "common.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cstdint>
#include <type_traits>

namespace cna {

// src [B][C][HW] -> dst [B][HW][c_pad]
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const TI* __restrict__ src, TO* __restrict__ dst,
                                                          int C, int HW, int c_pad) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const TI* s = src + (size_t)b * C * HW;
    TO* d = dst + (size_t)b * HW * c_pad;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = c0 + ty + i * 8, p = p0 + tx;
        tile[ty + i * 8][tx] = (c < C && p < HW) ? to_f32(ldg(s + (size_t)c * HW + p)) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int p = p0 + ty + i * 8, c = c0 + tx;
        if (p < HW && c < c_pad) d[(size_t)p * c_pad + c] = from_f32<TO>(tile[tx][ty + i * 8]);
    }
}

// Few-channel images (C <= 8: the network input) -> bf16 NHWC zero-padded to CPAD channels: a thread owns one pixel, reads its C
// planes (coalesced across the warp) and writes CPAD * 2 bytes with 16-byte stores.  The 32 x 32 tile transpose above leaves
// 29 of 32 tile rows empty for C = 3 and issues 2-byte stores; this is the e2e path's first kernel (bench.py, approx/pipeline.py).
template <typename TI, int CPAD>
__global__ void __launch_bounds__(256) nchw_to_nhwc_smallc_kernel(const TI* __restrict__ src, __nv_bfloat16* __restrict__ dst, int C,
                                                                  int HW, long long total) {
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / HW;
        const int p = (int)(idx - b * HW);
        const TI* s = src + (size_t)b * C * HW + p;
        alignas(16) __nv_bfloat16 v[CPAD];
#pragma unroll
        for (int c = 0; c < CPAD; ++c) v[c] = c < C ? from_f32<__nv_bfloat16>(to_f32(ldg(s + (size_t)c * HW))) : __float2bfloat16_rn(0.f);
        __nv_bfloat16* d = dst + (size_t)idx * CPAD;
        if constexpr (CPAD == 16) {
            if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) ptx::st_global_v8(d, *reinterpret_cast<const uint4*>(&v[0]), *reinterpret_cast<const uint4*>(&v[8]));
            else { reinterpret_cast<uint4*>(d)[0] = *reinterpret_cast<const uint4*>(&v[0]); reinterpret_cast<uint4*>(d)[1] = *reinterpret_cast<const uint4*>(&v[8]); }
        } else {
            reinterpret_cast<uint4*>(d)[0] = *reinterpret_cast<const uint4*>(&v[0]);
        }
    }
}

// uint8 images (what an image loader holds on the host: timm's prefetch loader ships uint8 NCHW batches to the GPU and normalises
// there, `input.float().sub_(mean * 255).div_(std * 255)`) -> normalised bf16 NHWC zero-padded to CPAD channels, in one pass.
// A thread owns 4 consecutive pixels: one 4-byte load per plane, CPAD * 8 bytes of stores.  HW % 4 == 0.
template <int CPAD>
__global__ void __launch_bounds__(256) u8_nchw_to_nhwc_norm_kernel(const uint8_t* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                                   const float* __restrict__ mean255, const float* __restrict__ inv_std255,
                                                                   int C, int HW, long long total4) {
    float mu[8], is[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { mu[c] = c < C ? mean255[c] : 0.f; is[c] = c < C ? inv_std255[c] : 0.f; }
    const int HW4 = HW >> 2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / HW4;
        const int p4 = (int)(idx - b * HW4);
        const uint8_t* s = src + (size_t)b * C * HW + (size_t)p4 * 4;
        uint32_t plane[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) plane[c] = c < C ? __ldg(reinterpret_cast<const uint32_t*>(s + (size_t)c * HW)) : 0u;
        __nv_bfloat16* d = dst + ((size_t)b * HW + (size_t)p4 * 4) * CPAD;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            alignas(16) __nv_bfloat16 v[CPAD];
#pragma unroll
            for (int c = 0; c < CPAD; ++c)
                v[c] = (c < 8 && c < C) ? __float2bfloat16_rn(((float)((plane[c < 8 ? c : 0] >> (8 * q)) & 0xffu) - mu[c < 8 ? c : 0]) * is[c < 8 ? c : 0])
                                        : __float2bfloat16_rn(0.f);
#pragma unroll
            for (int w = 0; w < CPAD / 8; ++w) reinterpret_cast<uint4*>(d + q * CPAD)[w] = *reinterpret_cast<const uint4*>(&v[8 * w]);
        }
    }
}

// src [B][HW][c_pad] -> dst [B][C][HW]
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const TI* __restrict__ src, TO* __restrict__ dst,
                                                          int C, int HW, int c_pad) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const TI* s = src + (size_t)b * HW * c_pad;
    TO* d = dst + (size_t)b * C * HW;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int p = p0 + ty + i * 8, c = c0 + tx;
        tile[ty + i * 8][tx] = (p < HW && c < C) ? to_f32(ldg(s + (size_t)p * c_pad + c)) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = c0 + ty + i * 8, p = p0 + tx;
        if (c < C && p < HW) d[(size_t)c * HW + p] = from_f32<TO>(tile[tx][ty + i * 8]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                        int B, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)B * Ho * Wo * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long pix = i / C;
        int wo = (int)(pix % Wo);
        int ho = (int)((pix / Wo) % Ho);
        int b = (int)(pix / ((long long)Wo * Ho));
        const T* p = x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C + c;
        float v = fmaxf(fmaxf(to_f32(ldg(p)), to_f32(ldg(p + C))),
                        fmaxf(to_f32(ldg(p + (size_t)W * C)), to_f32(ldg(p + (size_t)W * C + C))));
        y[i] = from_f32<T>(v);
    }
}

// 8 bf16 channels per thread (128-bit loads/stores); C % 8 == 0.
__global__ void __launch_bounds__(256) maxpool2x2_bf16x8_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                               int B, int H, int W, int C8) {
    const int Ho = H / 2, Wo = W / 2;
    const long long total = (long long)B * Ho * Wo * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C8);
        long long pix = i / C8;
        int wo = (int)(pix % Wo);
        int ho = (int)((pix / Wo) % Ho);
        int b = (int)(pix / ((long long)Wo * Ho));
        const uint4* p = x + (((size_t)b * H + 2 * ho) * W + 2 * wo) * C8 + c;
        uint4 a = __ldg(p), bq = __ldg(p + C8), cq = __ldg(p + (size_t)W * C8), dq = __ldg(p + (size_t)W * C8 + C8);
        uint4 r;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&bq);
        const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&cq);
        const __nv_bfloat162* pd = reinterpret_cast<const __nv_bfloat162*>(&dq);
        __nv_bfloat162* pr = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
        for (int k = 0; k < 4; ++k) pr[k] = __hmax2(__hmax2(pa[k], pb[k]), __hmax2(pc[k], pd[k]));
        y[i] = r;
    }
}

template <typename TI, typename TO>
static int launch_to_nhwc(const void* src, void* dst, int B, int C, int H, int W, int c_pad, cudaStream_t st) {
    if constexpr (std::is_same<TO, __nv_bfloat16>::value) {
        if (C <= 8 && (c_pad == 8 || c_pad == 16) && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            const long long total = (long long)B * H * W;
            const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 16);
            if (c_pad == 8) nchw_to_nhwc_smallc_kernel<TI, 8><<<grid, 256, 0, st>>>((const TI*)src, (__nv_bfloat16*)dst, C, H * W, total);
            else nchw_to_nhwc_smallc_kernel<TI, 16><<<grid, 256, 0, st>>>((const TI*)src, (__nv_bfloat16*)dst, C, H * W, total);
            CNA_LAUNCH_OK();
            return 0;
        }
    }
    dim3 grid(ceil_div(H * W, 32), ceil_div(c_pad, 32), B);
    nchw_to_nhwc_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)src, (TO*)dst, C, H * W, c_pad);
    CNA_LAUNCH_OK();
    return 0;
}
template <typename TI, typename TO>
static int launch_to_nchw(const void* src, void* dst, int B, int C, int H, int W, int c_pad, cudaStream_t st) {
    dim3 grid(ceil_div(H * W, 32), ceil_div(C, 32), B);
    nhwc_to_nchw_kernel<TI, TO><<<grid, 256, 0, st>>>((const TI*)src, (TO*)dst, C, H * W, c_pad);
    CNA_LAUNCH_OK();
    return 0;
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_nchw_to_nhwc(const void* src, void* dst, int B, int C, int H, int W, int c_pad,
                     int src_dtype, int dst_dtype, void* stream) {
    CNA_CHECK(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && c_pad >= C, "bad arguments");
    CNA_CHECK(B <= 65535, "batch too large for grid.z");
    cudaStream_t st = as_stream(stream);
    if (src_dtype == CNA_F32 && dst_dtype == CNA_F32) return launch_to_nhwc<float, float>(src, dst, B, C, H, W, c_pad, st);
    if (src_dtype == CNA_F32 && dst_dtype == CNA_BF16) return launch_to_nhwc<float, __nv_bfloat16>(src, dst, B, C, H, W, c_pad, st);
    if (src_dtype == CNA_BF16 && dst_dtype == CNA_BF16) return launch_to_nhwc<__nv_bfloat16, __nv_bfloat16>(src, dst, B, C, H, W, c_pad, st);
    if (src_dtype == CNA_BF16 && dst_dtype == CNA_F32) return launch_to_nhwc<__nv_bfloat16, float>(src, dst, B, C, H, W, c_pad, st);
    CNA_FAIL("unsupported dtype pair");
}

int cna_u8_nchw_to_nhwc_norm(const void* src_u8, void* dst_bf16, const float* mean255, const float* inv_std255, int B, int C, int H,
                             int W, int c_pad, void* stream) {
    CNA_CHECK(src_u8 && dst_bf16 && mean255 && inv_std255 && B > 0 && H > 0 && W > 0, "bad arguments");
    CNA_CHECK(C >= 1 && C <= 8 && (c_pad == 8 || c_pad == 16) && c_pad >= C, "1..8 image channels, padded to 8 or 16");
    CNA_CHECK((H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(src_u8) & 3) == 0 && (reinterpret_cast<uintptr_t>(dst_bf16) & 15) == 0,
              "H*W must be a multiple of 4 and the buffers 4- / 16-byte aligned");
    const long long total4 = (long long)B * H * W / 4;
    const int grid = (int)std::min<long long>(ceil_div_ll(total4, 256), 148LL * 16);
    cudaStream_t st = as_stream(stream);
    if (c_pad == 8) u8_nchw_to_nhwc_norm_kernel<8><<<grid, 256, 0, st>>>((const uint8_t*)src_u8, (__nv_bfloat16*)dst_bf16, mean255, inv_std255, C, H * W, total4);
    else u8_nchw_to_nhwc_norm_kernel<16><<<grid, 256, 0, st>>>((const uint8_t*)src_u8, (__nv_bfloat16*)dst_bf16, mean255, inv_std255, C, H * W, total4);
    CNA_LAUNCH_OK();
    return 0;
}

int cna_nhwc_to_nchw(const void* src, void* dst, int B, int C, int H, int W, int c_pad,
                     int src_dtype, int dst_dtype, void* stream) {
    CNA_CHECK(src && dst && B > 0 && C > 0 && H > 0 && W > 0 && c_pad >= C, "bad arguments");
    CNA_CHECK(B <= 65535, "batch too large for grid.z");
    cudaStream_t st = as_stream(stream);
    if (src_dtype == CNA_F32 && dst_dtype == CNA_F32) return launch_to_nchw<float, float>(src, dst, B, C, H, W, c_pad, st);
    if (src_dtype == CNA_F32 && dst_dtype == CNA_BF16) return launch_to_nchw<float, __nv_bfloat16>(src, dst, B, C, H, W, c_pad, st);
    if (src_dtype == CNA_BF16 && dst_dtype == CNA_BF16) return launch_to_nchw<__nv_bfloat16, __nv_bfloat16>(src, dst, B, C, H, W, c_pad, st);
    if (src_dtype == CNA_BF16 && dst_dtype == CNA_F32) return launch_to_nchw<__nv_bfloat16, float>(src, dst, B, C, H, W, c_pad, st);
    CNA_FAIL("unsupported dtype pair");
}

int cna_maxpool2x2(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
    CNA_CHECK(x && y && B > 0 && C > 0 && H >= 2 && W >= 2, "bad arguments");
    cudaStream_t st = as_stream(stream);
    long long total = (long long)B * (H / 2) * (W / 2) * C;
    if (dtype == CNA_BF16 && C % 8 == 0) {
        total /= 8;
        int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 32);
        maxpool2x2_bf16x8_kernel<<<grid, 256, 0, st>>>((const uint4*)x, (uint4*)y, B, H, W, C / 8);
        CNA_LAUNCH_OK();
        return 0;
    }
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 32);
    if (dtype == CNA_F32) maxpool2x2_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, B, H, W, C);
    else if (dtype == CNA_BF16) maxpool2x2_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, H, W, C);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
