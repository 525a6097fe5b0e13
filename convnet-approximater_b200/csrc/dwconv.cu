// Generic depthwise convolution (groups = Cin, channel multiplier `mult`) on NHWC tensors,
// with the reference's follow-up elementwise ops fused: bias, ParallelConv's running sum (`add`)
// and FixPaddingBias' border-row residues (`res`).  One thread per (pixel, output channel);
// channels are the fastest index so global accesses coalesce.  This is the general path (any
// kernel size / stride / padding); the MSCA-specific fused kernels live in msca.cu.
#include "common.cuh"

#include <stdlib.h>

namespace cna {

template <typename T>
__global__ void __launch_bounds__(256)
dwconv_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                  const float* __restrict__ res, int res_p, const T* __restrict__ add, T* __restrict__ y,
                  int B, int H, int W, int Cin, int mult, int Ho, int Wo, int kh, int kw,
                  int sh, int sw, int ph, int pw) {
    const int Cout = Cin * mult;
    const long long total = (long long)B * Ho * Wo * Cout;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int co = (int)(idx % Cout);
        long long pix = idx / Cout;
        int wo = (int)(pix % Wo);
        int ho = (int)((pix / Wo) % Ho);
        int b = (int)(pix / ((long long)Wo * Ho));
        int ci = co / mult;
        float acc = 0.f;
        for (int i = 0; i < kh; ++i) {
            int hi = ho * sh - ph + i;
            if (hi < 0 || hi >= H) continue;
            const T* xr = x + ((size_t)b * H + hi) * W * Cin + ci;
            const T* wr = w + (size_t)i * kw * Cout + co;
            for (int j = 0; j < kw; ++j) {
                int wi = wo * sw - pw + j;
                if (wi < 0 || wi >= W) continue;
                acc = fmaf(to_f32(ldg(xr + (size_t)wi * Cin)), to_f32(ldg(wr + (size_t)j * Cout)), acc);
            }
        }
        if (bias) acc += __ldg(bias + co);
        // Rounding points follow the reference's tensor materialisation (no-ops in fp32).
        if (res) {
            acc = to_f32(from_f32<T>(acc));
            int p2 = min(Ho, res_p);
            if (ho < p2) acc += __ldg(res + (size_t)co * res_p + ho);
            acc = to_f32(from_f32<T>(acc));
            if (ho >= Ho - p2) acc += __ldg(res + ((size_t)Cout + co) * res_p + (res_p - (Ho - ho)));
        }
        if (add) acc = to_f32(from_f32<T>(acc)) + to_f32(ldg(add + idx));
        y[idx] = from_f32<T>(acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
fix_padding_bias_kernel(T* __restrict__ x, const float* __restrict__ res, int B, int H, int W, int C, int p) {
    // rows touched: top [0,p2) and bottom [H-p2,H); when H < 2p they overlap and a row gets both residues,
    // top first then bottom, each add rounded like the reference's two in-place += statements.
    const int p2 = min(H, p);
    const int bot0 = max(p2, H - p2);          // first bottom row that is not also a top row
    const int nrows = p2 + (H - bot0);
    const long long total = (long long)B * nrows * W * C;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int c = (int)(idx % C);
        long long r = idx / C;
        int wv = (int)(r % W);
        int i = (int)((r / W) % nrows);
        int b = (int)(r / ((long long)W * nrows));
        int h = i < p2 ? i : bot0 + (i - p2);
        size_t o = (((size_t)b * H + h) * W + wv) * C + c;
        float v = to_f32(x[o]);
        if (h < p2) v = to_f32(from_f32<T>(v + __ldg(res + (size_t)c * p + h)));
        if (h >= H - p2) v += __ldg(res + ((size_t)C + c) * p + (p - (H - h)));
        x[o] = from_f32<T>(v);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o,
                                                 long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        o[i] = from_f32<T>(to_f32(ldg(a + i)) + to_f32(ldg(b + i)));
}

// (Cout,1,kh,kw) fp32 -> [kh*kw][Cout]
template <typename T>
__global__ void pack_dw_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int taps) {
    const long long total = (long long)taps * Cout;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int co = (int)(idx % Cout);
        int t = (int)(idx / Cout);
        out[idx] = from_f32<T>(w[(size_t)co * taps + t]);
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_pack_dw_weight(const float* w, void* packed, int Cout, int kh, int kw, int dtype, void* stream) {
    CNA_CHECK(w && packed && Cout > 0 && kh > 0 && kw > 0, "bad arguments");
    long long total = (long long)Cout * kh * kw;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 8);
    if (dtype == CNA_F32) pack_dw_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w, (float*)packed, Cout, kh * kw);
    else if (dtype == CNA_BF16) pack_dw_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w, (__nv_bfloat16*)packed, Cout, kh * kw);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_dwconv_forward(const void* x, const void* w_packed, const float* bias, const float* res, int res_p,
                       const void* add, void* y, int B, int H, int W, int Cin, int mult, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w, int dtype, void* stream) {
    CNA_CHECK(x && w_packed && y, "null pointer");
    CNA_CHECK(B > 0 && H > 0 && W > 0 && Cin > 0 && mult > 0 && kh > 0 && kw > 0, "bad dims");
    CNA_CHECK(stride_h > 0 && stride_w > 0 && pad_h >= 0 && pad_w >= 0, "bad stride/pad");
    CNA_CHECK(res == nullptr || res_p > 0, "res given without res_p");
    int Ho = (H + 2 * pad_h - kh) / stride_h + 1;
    int Wo = (W + 2 * pad_w - kw) / stride_w + 1;
    CNA_CHECK(Ho > 0 && Wo > 0, "empty output");
    long long total = (long long)B * Ho * Wo * Cin * mult;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 64);
    cudaStream_t st = as_stream(stream);
    if (getenv("CNA_DW_GENERIC") == nullptr) {      // (diagnostic switch: force the generic kernel)
        bool handled = false;
        if (dw_tile_try(x, w_packed, bias, res, res_p, add, y, B, H, W, Cin, mult, kh, kw, stride_h, stride_w, pad_h, pad_w, dtype, st,
                        &handled))
            return 1;
        if (handled) {
            CNA_LAUNCH_OK();
            return 0;
        }
    }
    if (dtype == CNA_F32)
        dwconv_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)w_packed, bias, res, res_p, (const float*)add,
                                                      (float*)y, B, H, W, Cin, mult, Ho, Wo, kh, kw, stride_h, stride_w, pad_h, pad_w);
    else if (dtype == CNA_BF16)
        dwconv_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed, bias, res, res_p,
                                                              (const __nv_bfloat16*)add, (__nv_bfloat16*)y, B, H, W, Cin, mult, Ho, Wo,
                                                              kh, kw, stride_h, stride_w, pad_h, pad_w);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_add(const void* a, const void* b, void* out, long long count, int dtype, void* stream) {
    CNA_CHECK(a && b && out && count > 0, "bad arguments");
    int grid = (int)std::min<long long>(ceil_div_ll(count, 256), 148LL * 32);
    if (dtype == CNA_F32) add_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)a, (const float*)b, (float*)out, count);
    else if (dtype == CNA_BF16) add_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, count);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_fix_padding_bias(void* x, const float* res, int B, int H, int W, int C, int p, int dtype, void* stream) {
    CNA_CHECK(x && res && B > 0 && H > 0 && W > 0 && C > 0 && p > 0, "bad arguments");
    int p2 = std::min(H, p);
    int nrows = p2 + (H - std::max(p2, H - p2));
    long long total = (long long)B * nrows * W * C;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 32);
    cudaStream_t st = as_stream(stream);
    if (dtype == CNA_F32) fix_padding_bias_kernel<float><<<grid, 256, 0, st>>>((float*)x, res, B, H, W, C, p);
    else if (dtype == CNA_BF16) fix_padding_bias_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((__nv_bfloat16*)x, res, B, H, W, C, p);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
