// Backward pieces for the depthwise layers of MSCA (reference: autograd through nn.Conv2d(groups=C) / FixPaddingBias / the
// `x *` gate, reached from total_loss.backward() in src/approx/hooks/finetune.py:329-330 when an MSCA block's sd_convs are the
// trainable replaced layer).  Correctness-first fp32-accumulating kernels:
//   * data gradient of a stride-1 'same' depthwise conv = the forward kernels on the flipped filter (no new kernel),
//   * weight gradient   dw[c][i][j] = sum_{b,h,w} x[b, h+i-ph, w+j-pw, c] * g[b,h,w,c]      (cna_dwconv_wgrad)
//   * FixPaddingBias    dres[0][c][h] / dres[1][c][p-(H-h)] = sum_{b,w} g[b,h,w,c] on the border rows   (cna_fix_padding_bias_grad)
//   * elementwise product for the gate                                                        (cna_mul)
#include "common.cuh"

namespace cna {

constexpr int DWG_CH = 32;      // channels per CTA
constexpr int DWG_LANES = 8;    // pixel lanes per channel

// grid (ceil(C/32), slabs, taps): one tap, 32 channels, a slab of pixels; fp32 atomics into dw[c][tap]
template <typename T>
__global__ void __launch_bounds__(DWG_CH * DWG_LANES)
dw_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ g, float* __restrict__ dw, int B, int H, int W, int C, int mult, int kh,
                int kw, int ph, int pw, long long pix_per_cta) {   // C = OUTPUT channels (= Cin * mult); output channel c reads input channel c / mult
    const int Cin = C / mult;
    const int c = blockIdx.x * DWG_CH + (threadIdx.x & (DWG_CH - 1));
    const int lane = threadIdx.x / DWG_CH;
    const int tap = blockIdx.z, ti = tap / kw, tj = tap - ti * kw;
    const long long P = (long long)B * H * W;
    const long long p_beg = (long long)blockIdx.y * pix_per_cta, p_end = min(P, p_beg + pix_per_cta);
    float acc = 0.f;
    if (c < C) {
        for (long long p = p_beg + lane; p < p_end; p += DWG_LANES) {
            const int w = (int)(p % W);
            const int h = (int)((p / W) % H);
            const int hi = h + ti - ph, wi = w + tj - pw;
            if (hi < 0 || hi >= H || wi < 0 || wi >= W) continue;
            const long long q = p + (long long)(ti - ph) * W + (tj - pw);
            acc = fmaf(to_f32(ldg(x + q * Cin + c / mult)), to_f32(ldg(g + p * C + c)), acc);
        }
    }
    __shared__ float red[DWG_LANES][DWG_CH + 1];
    red[lane][threadIdx.x & (DWG_CH - 1)] = acc;
    __syncthreads();
    if (lane == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < DWG_LANES; ++l) s += red[l][threadIdx.x];
        atomicAdd(dw + (size_t)c * kh * kw + tap, s);
    }
}

// grid (ceil(C/32), rows touched): row sums over (b, w) of g on one border row
template <typename T>
__global__ void __launch_bounds__(DWG_CH * DWG_LANES)
fix_bias_grad_kernel(const T* __restrict__ g, float* __restrict__ dres, int B, int H, int W, int C, int p) {
    const int p2 = min(H, p);
    const int c = blockIdx.x * DWG_CH + (threadIdx.x & (DWG_CH - 1));
    const int lane = threadIdx.x / DWG_CH;
    const int r = blockIdx.y;                       // 0..p2-1: top row r; p2..2*p2-1: bottom row H - p2 + (r - p2)
    const bool top = r < p2;
    const int h = top ? r : H - p2 + (r - p2);
    float acc = 0.f;
    if (c < C)
        for (long long i = lane; i < (long long)B * W; i += DWG_LANES) {
            const int b = (int)(i / W), w = (int)(i - (long long)b * W);
            acc += to_f32(ldg(g + (((size_t)b * H + h) * W + w) * C + c));
        }
    __shared__ float red[DWG_LANES][DWG_CH + 1];
    red[lane][threadIdx.x & (DWG_CH - 1)] = acc;
    __syncthreads();
    if (lane == 0 && c < C) {
        float s = 0.f;
#pragma unroll
        for (int l = 0; l < DWG_LANES; ++l) s += red[l][threadIdx.x];
        if (top) dres[(size_t)c * p + h] = s;
        else dres[((size_t)C + c) * p + (p - (H - h))] = s;
    }
}

// Data gradient of a stride-1 'same' depthwise conv with channel multiplier `mult`:
//   gx[b,h,w,ci] = sum_m sum_{i,j} g[b, h-(i-ph), w-(j-pw), ci*mult+m] * w[ci*mult+m][i][j]        (w packed [kh*kw][Cout])
template <typename T>
__global__ void __launch_bounds__(256)
dw_dgrad_mult_kernel(const T* __restrict__ g, const T* __restrict__ w, T* __restrict__ gx, int B, int H, int W, int Cin, int mult, int kh,
                     int kw, int ph, int pw) {
    const int Cout = Cin * mult;
    const long long total = (long long)B * H * W * Cin;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(idx % Cin);
        const long long pix = idx / Cin;
        const int wq = (int)(pix % W), hq = (int)((pix / W) % H);
        const int b = (int)(pix / ((long long)W * H));
        float acc = 0.f;
        for (int i = 0; i < kh; ++i) {
            const int ho = hq - (i - ph);
            if (ho < 0 || ho >= H) continue;
            for (int j = 0; j < kw; ++j) {
                const int wo = wq - (j - pw);
                if (wo < 0 || wo >= W) continue;
                const T* gp = g + (((size_t)b * H + ho) * W + wo) * Cout + (size_t)ci * mult;
                const T* wp = w + (size_t)(i * kw + j) * Cout + (size_t)ci * mult;
                for (int m = 0; m < mult; ++m) acc = fmaf(to_f32(ldg(gp + m)), to_f32(ldg(wp + m)), acc);
            }
        }
        gx[idx] = from_f32<T>(acc);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) mul_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        o[i] = from_f32<T>(to_f32(ldg(a + i)) * to_f32(ldg(b + i)));
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_dwconv_wgrad(const void* x, const void* g, float* dw, int B, int H, int W, int Cin, int mult, int kh, int kw, int pad_h, int pad_w,
                     int dtype, void* stream) {
    CNA_CHECK(x && g && dw && B > 0 && H > 0 && W > 0 && Cin > 0 && mult > 0 && kh > 0 && kw > 0, "bad arguments");
    const int C = Cin * mult;
    CNA_CHECK(2 * pad_h == kh - 1 && 2 * pad_w == kw - 1, "implemented for stride-1 'same' depthwise convolutions");
    cudaStream_t st = as_stream(stream);
    const int taps = kh * kw;
    CNA_CHECK(taps <= 65535, "too many taps");
    CNA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)C * taps, st));
    const long long P = (long long)B * H * W;
    const int cblocks = ceil_div(C, DWG_CH);
    int slabs = (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(P, 512), (148LL * 8) / std::max(1, cblocks * std::min(taps, 32))));
    const long long ppc = ceil_div_ll(P, slabs);
    slabs = (int)ceil_div_ll(P, ppc);
    dim3 grid(cblocks, slabs, taps);
    if (dtype == CNA_F32) dw_wgrad_kernel<float><<<grid, DWG_CH * DWG_LANES, 0, st>>>((const float*)x, (const float*)g, dw, B, H, W, C, mult, kh, kw, pad_h, pad_w, ppc);
    else if (dtype == CNA_BF16) dw_wgrad_kernel<__nv_bfloat16><<<grid, DWG_CH * DWG_LANES, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)g, dw, B, H, W, C, mult, kh, kw, pad_h, pad_w, ppc);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_dwconv_dgrad(const void* g, const void* w_packed, void* gx, int B, int H, int W, int Cin, int mult, int kh, int kw, int pad_h,
                     int pad_w, int dtype, void* stream) {
    CNA_CHECK(g && w_packed && gx && B > 0 && H > 0 && W > 0 && Cin > 0 && mult > 0 && kh > 0 && kw > 0, "bad arguments");
    CNA_CHECK(2 * pad_h == kh - 1 && 2 * pad_w == kw - 1, "implemented for stride-1 'same' depthwise convolutions");
    const long long total = (long long)B * H * W * Cin;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 64);
    cudaStream_t st = as_stream(stream);
    if (dtype == CNA_F32) dw_dgrad_mult_kernel<float><<<grid, 256, 0, st>>>((const float*)g, (const float*)w_packed, (float*)gx, B, H, W, Cin, mult, kh, kw, pad_h, pad_w);
    else if (dtype == CNA_BF16) dw_dgrad_mult_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)w_packed, (__nv_bfloat16*)gx, B, H, W, Cin, mult, kh, kw, pad_h, pad_w);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_fix_padding_bias_grad(const void* g, float* dres, int B, int H, int W, int C, int p, int dtype, void* stream) {
    CNA_CHECK(g && dres && B > 0 && H > 0 && W > 0 && C > 0 && p > 0, "bad arguments");
    cudaStream_t st = as_stream(stream);
    CNA_CUDA(cudaMemsetAsync(dres, 0, sizeof(float) * 2 * (size_t)C * p, st));
    const int p2 = std::min(H, p);
    dim3 grid(ceil_div(C, DWG_CH), 2 * p2);
    if (dtype == CNA_F32) fix_bias_grad_kernel<float><<<grid, DWG_CH * DWG_LANES, 0, st>>>((const float*)g, dres, B, H, W, C, p);
    else if (dtype == CNA_BF16) fix_bias_grad_kernel<__nv_bfloat16><<<grid, DWG_CH * DWG_LANES, 0, st>>>((const __nv_bfloat16*)g, dres, B, H, W, C, p);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_mul(const void* a, const void* b, void* out, long long count, int dtype, void* stream) {
    CNA_CHECK(a && b && out && count > 0, "bad arguments");
    int grid = (int)std::min<long long>(ceil_div_ll(count, 256), 148LL * 32);
    if (dtype == CNA_F32) mul_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)a, (const float*)b, (float*)out, count);
    else if (dtype == CNA_BF16) mul_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (__nv_bfloat16*)out, count);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
