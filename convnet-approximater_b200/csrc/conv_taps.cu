// Generic dense "tap" convolution on CUDA cores (FFMA, fp32 accumulate).
//
// One kernel covers every dense conv on the hot path whose filter is a run of taps along one
// axis: the (d,1) v_conv and dense (1,d) h_conv of scheme 2 (any stride / padding), and all 1x1
// convs (scheme-1 d_conv, MSCA channel_mix with its `x *` gate fused).  It is the exact-fp32 path
// (the reference's 1e-5 tolerance cannot be met by TF32 tensor-core math) and the fallback for
// shapes the fused tcgen05 kernel does not take.  Shared-memory tiled SGEMM: 64 pixels x 64
// outputs per CTA, K-step 16, 4x4 register micro-tile.
#include "common.cuh"

namespace cna {

constexpr int BM = 64, BN = 64, BK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
conv_taps_fwd_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                     const T* __restrict__ gate, T* __restrict__ y,
                     int B, int H, int W, int C, int Ho, int Wo, int N, int taps, int axis,
                     int sh, int sw, int pad, int relu) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const long long Mtot = (long long)B * Ho * Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-load role: pixel a_p (0..63), channel quad a_c (0..3)
    const int a_p = tid >> 2, a_c = (tid & 3) * 4;
    long long am = m0 + a_p;
    const bool a_valid = am < Mtot;
    int a_b = 0, a_ho = 0, a_wo = 0;
    if (a_valid) {
        a_wo = (int)(am % Wo);
        a_ho = (int)((am / Wo) % Ho);
        a_b = (int)(am / ((long long)Wo * Ho));
    }
    // B-load role: row b_r (0..15), column quad b_c
    const int b_r = tid >> 4, b_c = (tid & 15) * 4;

    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < taps; ++t) {
        int hi = a_ho * sh + (axis == 0 ? t - pad : 0);
        int wi = a_wo * sw + (axis == 1 ? t - pad : 0);
        const bool inb = a_valid && hi >= 0 && hi < H && wi >= 0 && wi < W;
        const T* xrow = x + (((size_t)a_b * H + (inb ? hi : 0)) * W + (inb ? wi : 0)) * C;
        const T* wt = w + (size_t)t * C * N;
        for (int c0 = 0; c0 < C; c0 += BK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = c0 + a_c + j;
                As[a_c + j][a_p] = (inb && c < C) ? to_f32(ldg(xrow + c)) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = c0 + b_r, n = n0 + b_c + j;
                Bs[b_r][b_c + j] = (c < C && n < N) ? to_f32(ldg(wt + (size_t)c * N + n)) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        long long m = m0 + ty * 4 + i;
        if (m >= Mtot) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? __ldg(bias + n) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            if (gate) {
                // the reference materialises channel_mix's output before the multiply (msca.py:26)
                v = to_f32(from_f32<T>(v)) * to_f32(ldg(gate + (size_t)m * N + n));
            }
            y[(size_t)m * N + n] = from_f32<T>(v);
        }
    }
}

// w[o][i][t] fp32 -> packed[t][i_pad][o_pad]
template <typename T>
__global__ void pack_taps_kernel(const float* __restrict__ w, T* __restrict__ out, int O, int I, int taps,
                                 int o_pad, int i_pad) {
    const long long total = (long long)taps * i_pad * o_pad;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int o = (int)(idx % o_pad);
        int i = (int)((idx / o_pad) % i_pad);
        int t = (int)(idx / ((long long)o_pad * i_pad));
        float v = (o < O && i < I) ? w[((size_t)o * I + i) * taps + t] : 0.f;
        out[idx] = from_f32<T>(v);
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_pack_taps_weight(const float* w_oit, void* packed, int O, int I, int taps, int o_pad, int i_pad,
                         int dtype, void* stream) {
    CNA_CHECK(w_oit && packed && O > 0 && I > 0 && taps > 0 && o_pad >= O && i_pad >= I, "bad arguments");
    long long total = (long long)taps * i_pad * o_pad;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 8);
    if (dtype == CNA_F32)
        pack_taps_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w_oit, (float*)packed, O, I, taps, o_pad, i_pad);
    else if (dtype == CNA_BF16)
        pack_taps_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w_oit, (__nv_bfloat16*)packed, O, I, taps, o_pad, i_pad);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_conv_taps_forward(const void* x, const void* w_packed, const float* bias, const void* gate, void* y,
                          int B, int H, int W, int C, int N, int taps, int axis,
                          int stride_h, int stride_w, int pad, int relu, int dtype, void* stream) {
    CNA_CHECK(x && w_packed && y, "null pointer");
    CNA_CHECK(B > 0 && H > 0 && W > 0 && C > 0 && N > 0 && taps > 0, "bad dims");
    CNA_CHECK(axis == 0 || axis == 1, "axis must be 0 (H) or 1 (W)");
    CNA_CHECK(stride_h > 0 && stride_w > 0 && pad >= 0, "bad stride/pad");
    int Ho = axis == 0 ? (H + 2 * pad - taps) / stride_h + 1 : (H - 1) / stride_h + 1;
    int Wo = axis == 1 ? (W + 2 * pad - taps) / stride_w + 1 : (W - 1) / stride_w + 1;
    CNA_CHECK(Ho > 0 && Wo > 0, "empty output");
    long long Mtot = (long long)B * Ho * Wo;
    dim3 grid((unsigned)ceil_div_ll(Mtot, BM), (unsigned)ceil_div(N, BN));
    CNA_CHECK(grid.y <= 65535, "N too large");
    cudaStream_t st = as_stream(stream);
    if (dtype == CNA_F32)
        conv_taps_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)w_packed, bias, (const float*)gate,
                                                         (float*)y, B, H, W, C, Ho, Wo, N, taps, axis, stride_h, stride_w,
                                                         pad, relu);
    else if (dtype == CNA_BF16)
        conv_taps_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w_packed, bias,
                                                                 (const __nv_bfloat16*)gate, (__nv_bfloat16*)y, B, H, W, C, Ho,
                                                                 Wo, N, taps, axis, stride_h, stride_w, pad, relu);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
