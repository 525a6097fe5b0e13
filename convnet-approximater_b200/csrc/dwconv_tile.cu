// Register-tiled depthwise convolution for the shapes MSCA uses (reference layers/msca.py:18-19 conv0 5x5,
// layers/depth_seperable_conv.py:12-16 strip convs 1xk / kx1, core/msca_rep.py:44-45 the merged 21x21 kernel):
// stride 1, 'same' padding, channel multiplier 1, C % 32 == 0, NHWC, bf16 or fp32 storage, fp32 accumulation.
//
// One CTA = 16 x 16 output pixels x 32 channels.  The input tile with its halo is staged once in shared memory
// (128-bit global loads, zero fill outside the image = the conv padding); a thread owns one channel PAIR and a run
// of 16 outputs along the longer kernel axis, so each staged input value is loaded from shared memory once per
// kernel column (row) and reused for up to 16 FMAs per channel, and each weight once per 16 outputs: ~10 FMAs per
// shared-memory load for the 21x21 kernel (FP32-pipe bound: 441 MAC/element), while the 1xk / kx1 strips and the
// 5x5 stay what they are on paper — one HBM read + one write per element.  Lanes of a warp run over
// the 16 channel pairs of two adjacent pixels: every shared-memory access is one conflict-free 128-byte wavefront.
// Bias, ParallelConv's running sum (`add`) and FixPaddingBias' border residues (`res`) are fused like in the
// generic kernel (dwconv.cu), with the same rounding points.
#include "common.cuh"

namespace cna {

constexpr int kDT = 16;        // tile edge (outputs) and run length per thread
constexpr int kDC = 32;        // channels per CTA

template <typename T> struct Pair;
template <> struct Pair<__nv_bfloat16> {
    using Raw = uint32_t;
    static __device__ __forceinline__ float2 unpack(uint32_t v) {
        return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
    }
    static __device__ __forceinline__ uint32_t pack(float a, float b) {
        __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    static __device__ __forceinline__ float round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
};
template <> struct Pair<float> {
    using Raw = float2;
    static __device__ __forceinline__ float2 unpack(float2 v) { return v; }
    static __device__ __forceinline__ float2 pack(float a, float b) { return make_float2(a, b); }
    static __device__ __forceinline__ float round(float a) { return a; }
};

// ALONG_W = false: thread = (channel pair, column), 16 outputs down the rows (used when KH >= KW)
// ALONG_W = true : thread = (channel pair, row),    16 outputs along the columns (1 x k strips)
template <typename T, int KH, int KW, bool ALONG_W>
__global__ void __launch_bounds__(256)
dw_tile_kernel(const T* __restrict__ x, const T* __restrict__ w /*[KH*KW][C]*/, const float* __restrict__ bias,
               const float* __restrict__ res, int res_p, const T* __restrict__ add, T* __restrict__ y,
               int B, int H, int W, int C, int tiles_h, int tiles_w, long long ntiles) {
    using P = Pair<T>;
    using Raw = typename P::Raw;
    constexpr int XH = kDT + KH - 1, XW = kDT + KW - 1;        // staged tile incl. halo
    constexpr int PH = KH / 2, PW = KW / 2;
    constexpr int CP = kDC / 2;                                // channel pairs per CTA
    extern __shared__ __align__(16) uint8_t dsm[];
    Raw* sx = reinterpret_cast<Raw*>(dsm);                     // [XH][XW][CP]
    Raw* sw = sx + XH * XW * CP;                               // [KH*KW][CP]
    const int cp = threadIdx.x & (CP - 1);
    const int pos = threadIdx.x >> 4;                          // 0..15: column (ALONG_W=false) or row (true)
    const int ncb = C / kDC;
    int w_cb = -1;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        // tile order: channel block slowest, so a CTA re-stages its weights only when the block changes
        const long long per_cb = (long long)B * tiles_h * tiles_w;
        const int cb = (int)(tile / per_cb);
        long long rem = tile - cb * per_cb;
        const int b = (int)(rem / (tiles_h * tiles_w));
        rem -= (long long)b * tiles_h * tiles_w;
        const int th = (int)(rem / tiles_w), tw = (int)(rem - (long long)th * tiles_w);
        const int h0 = th * kDT, w0 = tw * kDT, c0 = cb * kDC;
        __syncthreads();                                       // previous tile's readers are done with sx / sw
        if (cb != w_cb) {
            constexpr int WV = KH * KW * CP;
            for (int i = threadIdx.x; i < WV; i += 256) {
                const int t = i / CP, c = i - t * CP;
                sw[i] = *reinterpret_cast<const Raw*>(w + (size_t)t * C + c0 + 2 * c);
            }
            w_cb = cb;
        }
        {   // stage the input tile: 16-byte chunks, zero outside the image
            constexpr int CHUNK = 16 / sizeof(T);              // channels per 16-byte chunk
            constexpr int NCH = kDC / CHUNK;                   // chunks per pixel
            constexpr int TOTAL = XH * XW * NCH;
            for (int i = threadIdx.x; i < TOTAL; i += 256) {
                const int pix = i / NCH, ch = i - pix * NCH;
                const int r = pix / XW, c = pix - r * XW;
                const int hi = h0 - PH + r, wi = w0 - PW + c;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (hi >= 0 && hi < H && wi >= 0 && wi < W)
                    v = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + hi) * W + wi) * C + c0 + ch * CHUNK));
                *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(sx) + ((size_t)pix * NCH + ch) * 16) = v;
            }
        }
        __syncthreads();

        float2 acc[kDT];
#pragma unroll
        for (int r = 0; r < kDT; ++r) acc[r] = make_float2(0.f, 0.f);
        if constexpr (!ALONG_W) {
            for (int j = 0; j < KW; ++j) {
                float2 wj[KH];
#pragma unroll
                for (int t = 0; t < KH; ++t) wj[t] = P::unpack(sw[(t * KW + j) * CP + cp]);
                const Raw* col = sx + (pos + j) * CP + cp;
#pragma unroll
                for (int i = 0; i < XH; ++i) {
                    const float2 xv = P::unpack(col[i * XW * CP]);
#pragma unroll
                    for (int r = 0; r < kDT; ++r) {
                        if (i - r >= 0 && i - r < KH) {
                            acc[r].x = fmaf(xv.x, wj[i - r].x, acc[r].x);
                            acc[r].y = fmaf(xv.y, wj[i - r].y, acc[r].y);
                        }
                    }
                }
            }
        } else {
            for (int i = 0; i < KH; ++i) {
                float2 wi_[KW];
#pragma unroll
                for (int t = 0; t < KW; ++t) wi_[t] = P::unpack(sw[(i * KW + t) * CP + cp]);
                const Raw* row = sx + ((pos + i) * XW) * CP + cp;
#pragma unroll
                for (int j = 0; j < XW; ++j) {
                    const float2 xv = P::unpack(row[j * CP]);
#pragma unroll
                    for (int r = 0; r < kDT; ++r) {
                        if (j - r >= 0 && j - r < KW) {
                            acc[r].x = fmaf(xv.x, wi_[j - r].x, acc[r].x);
                            acc[r].y = fmaf(xv.y, wi_[j - r].y, acc[r].y);
                        }
                    }
                }
            }
        }

        // epilogue: bias, border residues, running sum — rounding points as in dwconv.cu / the reference's tensors
        const int c = c0 + 2 * cp;
        const float2 bz = bias ? make_float2(__ldg(bias + c), __ldg(bias + c + 1)) : make_float2(0.f, 0.f);
        const int p2 = res ? min(H, res_p) : 0;
#pragma unroll
        for (int r = 0; r < kDT; ++r) {
            const int ho = ALONG_W ? h0 + pos : h0 + r;
            const int wo = ALONG_W ? w0 + r : w0 + pos;
            if (ho >= H || wo >= W) continue;
            float a0 = acc[r].x + bz.x, a1 = acc[r].y + bz.y;
            if (res) {
                a0 = P::round(a0); a1 = P::round(a1);
                if (ho < p2) { a0 += __ldg(res + (size_t)c * res_p + ho); a1 += __ldg(res + (size_t)(c + 1) * res_p + ho); }
                a0 = P::round(a0); a1 = P::round(a1);
                if (ho >= H - p2) {
                    a0 += __ldg(res + ((size_t)C + c) * res_p + (res_p - (H - ho)));
                    a1 += __ldg(res + ((size_t)C + c + 1) * res_p + (res_p - (H - ho)));
                }
            }
            const size_t o = (((size_t)b * H + ho) * W + wo) * C + c;
            if (add) {
                const float2 av = P::unpack(*reinterpret_cast<const Raw*>(add + o));
                a0 = P::round(a0) + av.x; a1 = P::round(a1) + av.y;
            }
            *reinterpret_cast<Raw*>(y + o) = P::pack(a0, a1);
        }
    }
}

template <typename T, int KH, int KW>
static int dw_tile_launch_t(const void* x, const void* w, const float* bias, const float* res, int res_p, const void* add, void* y,
                            int B, int H, int W, int C, cudaStream_t st) {
    constexpr bool ALONG_W = KW > KH;
    constexpr int XH = kDT + KH - 1, XW = kDT + KW - 1;
    const size_t smem = ((size_t)XH * XW + (size_t)KH * KW) * (kDC / 2) * 2 * sizeof(T);
    auto kern = dw_tile_kernel<T, KH, KW, ALONG_W>;
    static bool attr_set = false;
    if (!attr_set) {
        CNA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int tiles_h = ceil_div(H, kDT), tiles_w = ceil_div(W, kDT);
    const long long ntiles = (long long)B * tiles_h * tiles_w * (C / kDC);
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (216 * 1024) / (smem + 1024)));
    const int grid = (int)std::min<long long>(ntiles, 148LL * per_sm);
    kern<<<grid, 256, smem, st>>>((const T*)x, (const T*)w, bias, res, res_p, (const T*)add, (T*)y, B, H, W, C, tiles_h, tiles_w, ntiles);
    return 0;
}

template <typename T>
static int dw_tile_dispatch(int kh, int kw, const void* x, const void* w, const float* bias, const float* res, int res_p, const void* add,
                            void* y, int B, int H, int W, int C, cudaStream_t st, bool* handled) {
    *handled = true;
#define CNA_DW_CASE(KHV, KWV) \
    if (kh == KHV && kw == KWV) return dw_tile_launch_t<T, KHV, KWV>(x, w, bias, res, res_p, add, y, B, H, W, C, st)
    CNA_DW_CASE(5, 5);
    CNA_DW_CASE(3, 3);
    CNA_DW_CASE(21, 21);
    CNA_DW_CASE(1, 7);
    CNA_DW_CASE(7, 1);
    CNA_DW_CASE(1, 11);
    CNA_DW_CASE(11, 1);
    CNA_DW_CASE(1, 21);
    CNA_DW_CASE(21, 1);
    CNA_DW_CASE(1, 3);
    CNA_DW_CASE(3, 1);
    CNA_DW_CASE(1, 5);
    CNA_DW_CASE(5, 1);
#undef CNA_DW_CASE
    *handled = false;
    return 0;
}

// Returns 0 and sets *handled when this file has a kernel for the configuration (launch already enqueued).
int dw_tile_try(const void* x, const void* w_packed, const float* bias, const float* res, int res_p, const void* add, void* y,
                int B, int H, int W, int Cin, int mult, int kh, int kw, int sh, int sw, int ph, int pw, int dtype, cudaStream_t st,
                bool* handled) {
    *handled = false;
    if (mult != 1 || sh != 1 || sw != 1 || ph != kh / 2 || pw != kw / 2 || (kh & 1) == 0 || (kw & 1) == 0 || Cin % kDC != 0) return 0;
    if (dtype == CNA_BF16) return dw_tile_dispatch<__nv_bfloat16>(kh, kw, x, w_packed, bias, res, res_p, add, y, B, H, W, Cin, st, handled);
    if (dtype == CNA_F32) return dw_tile_dispatch<float>(kh, kw, x, w_packed, bias, res, res_p, add, y, B, H, W, Cin, st, handled);
    return 0;
}

}  // namespace cna
