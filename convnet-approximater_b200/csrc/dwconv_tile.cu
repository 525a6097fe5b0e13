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

#include <stdlib.h>
#include <type_traits>

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


// Blackwell's packed fp32 FMA (SASS FFMA2): two independent fp32 FMAs per issued instruction — the same FLOPs per FMA-pipe
// cycle, half the issue slots, which is what an instruction-issue-bound kernel is short of.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

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
                        if (i - r >= 0 && i - r < KH) acc[r] = ffma2(xv, wj[i - r], acc[r]);
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
                        if (j - r >= 0 && j - r < KW) acc[r] = ffma2(xv, wi_[j - r], acc[r]);
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


// ------------------------------------------------------------------------------------------------
// 1-D strips (1 x K along W, K x 1 along H): HBM-bound for K = 7 / 11 (7-11 MAC per element against 4 bytes), so the
// design goal is few INSTRUCTIONS per element (a B200 SM can issue ~22 thread-instructions per element at the HBM rate):
// no shared memory at all — a thread owns 4 channels (one 8-byte bf16 / 16-byte fp32 vector) and a run of 16 (8 for K > 7) outputs
// along the strip axis, keeps the K + 15 inputs it needs in a register window (each loaded and unpacked once, straight
// from global memory; neighbouring runs' halo re-reads hit L1/L2) and its 4 x K weights in registers:
// ~11 (K = 7) to ~28 (K = 21) instructions per element.  Lanes run over channel groups, then pixels, so loads coalesce.
// ------------------------------------------------------------------------------------------------
// Outputs per thread along the strip axis and resident CTAs per SM.  The strips are HBM-bound, so what matters is how many loads are in
// flight per SM, i.e. OCCUPANCY: with 16-output runs the K = 7 kernel needed 170-210 registers — one 8-warp CTA per SM, whose load
// phases and FMA phases never overlapped (43 % of the HBM peak by ncu DRAM bytes, profiles/r02_dwconv_b256_ncu_summary.csv).
// 8-output runs fit 2 CTAs = 16 warps per SM without spilling: 1x7 / 7x1 0.51, 1x11 / 11x1 0.45 of the HBM peak at B = 256
// (4-output runs at 3 CTAs per SM were measured too: 0.47 / 0.42 — more loads per output than the extra warps hide).  ncu on the
// 1x7 kernel: long-scoreboard stalls dominate (3.9 per issue), 43 % issue-active: the next step for these is a TMA-fed shared-memory
// pipeline, not more tuning of this register-window form.
template <int K> struct StripRun { static constexpr int value = 8; };
template <int K> struct StripCtas { static constexpr int value = 2; };


template <typename T> struct Vec4;
template <> struct Vec4<__nv_bfloat16> {
    using Raw = uint2;
    static __device__ __forceinline__ void unpack(uint2 v, float (&f)[4]) {
        f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
        f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    }
    static __device__ __forceinline__ uint2 pack(const float (&f)[4]) {
        __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
        return make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
    }
    static __device__ __forceinline__ float round(float a) { return __bfloat162float(__float2bfloat16_rn(a)); }
    static __device__ __forceinline__ uint2 zero() { return make_uint2(0u, 0u); }
};
template <> struct Vec4<float> {
    using Raw = float4;
    static __device__ __forceinline__ void unpack(float4 v, float (&f)[4]) { f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w; }
    static __device__ __forceinline__ float4 pack(const float (&f)[4]) { return make_float4(f[0], f[1], f[2], f[3]); }
    static __device__ __forceinline__ float round(float a) { return a; }
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};

template <typename T, int K, bool ALONG_W, bool EXTRAS>
__global__ void __launch_bounds__(256, StripCtas<K>::value)
dw_strip_kernel(const T* __restrict__ x, const T* __restrict__ w /*[K][C]*/, const float* __restrict__ bias,
                const float* __restrict__ res, int res_p, const T* __restrict__ add, T* __restrict__ y,
                int B, int H, int W, int C, long long nitems) {
    constexpr int kSR = StripRun<K>::value;
    using V = Vec4<T>;
    using Raw = typename V::Raw;
    constexpr int P = K / 2;
    const int ncg = C >> 2;                                   // channel groups of 4
    const int L = ALONG_W ? W : H;                            // length of the strip axis
    const int lines = ALONG_W ? H : W;                        // lines across it
    const int runs = (L + kSR - 1) / kSR;
    const size_t step = ALONG_W ? (size_t)C : (size_t)W * C;  // element stride along the strip axis
    // 32-bit index arithmetic (the launcher guarantees nitems < 2^31): 64-bit div/mod would cost more than the convolution
    const uint32_t n32 = (uint32_t)nitems, stride32 = gridDim.x * blockDim.x;
    for (uint32_t it = blockIdx.x * blockDim.x + threadIdx.x; it < n32; it += stride32) {
        // item order: channel group fastest, then (ALONG_W: run, line | ALONG_H: line, run), then image
        const uint32_t cg = it % (uint32_t)ncg;
        uint32_t r = it / (uint32_t)ncg;
        uint32_t run, line;
        if (ALONG_W) { run = r % (uint32_t)runs; r /= (uint32_t)runs; line = r % (uint32_t)lines; r /= (uint32_t)lines; }
        else         { line = r % (uint32_t)lines; r /= (uint32_t)lines; run = r % (uint32_t)runs; r /= (uint32_t)runs; }
        const int b = (int)r;
        const int c = (int)cg * 4;
        const int s0 = (int)run * kSR;                        // first output along the strip axis
        const size_t base = ALONG_W ? (((size_t)b * H + line) * W) * C + c : (((size_t)b * H) * W + line) * C + c;
        float2 wt[K][2];
#pragma unroll
        for (int t = 0; t < K; ++t) {
            float f[4];
            V::unpack(__ldg(reinterpret_cast<const Raw*>(w + (size_t)t * C + c)), f);
            wt[t][0] = make_float2(f[0], f[1]); wt[t][1] = make_float2(f[2], f[3]);
        }
        float2 bz0 = make_float2(0.f, 0.f), bz1 = bz0;
        if (bias) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c)); bz0 = make_float2(b4.x, b4.y); bz1 = make_float2(b4.z, b4.w); }
        float2 acc[kSR][2];                                   // accumulators start at the bias
#pragma unroll
        for (int o = 0; o < kSR; ++o) { acc[o][0] = bz0; acc[o][1] = bz1; }
#pragma unroll
        for (int i = 0; i < kSR + K - 1; ++i) {
            const int s = s0 - P + i;
            float xv[4];
            Raw raw = V::zero();
            if (s >= 0 && s < L) raw = __ldg(reinterpret_cast<const Raw*>(x + base + (size_t)s * step));
            V::unpack(raw, xv);
            const float2 x0 = make_float2(xv[0], xv[1]), x1 = make_float2(xv[2], xv[3]);
#pragma unroll
            for (int o = 0; o < kSR; ++o) {
                if (i - o >= 0 && i - o < K) {
                    acc[o][0] = ffma2(x0, wt[i - o][0], acc[o][0]);
                    acc[o][1] = ffma2(x1, wt[i - o][1], acc[o][1]);
                }
            }
        }
        const int p2 = (EXTRAS && res) ? min(H, res_p) : 0;
#pragma unroll
        for (int o = 0; o < kSR; ++o) {
            const int s = s0 + o;
            if (s >= L) break;
            const int ho = ALONG_W ? line : s;
            float a[4] = {acc[o][0].x, acc[o][0].y, acc[o][1].x, acc[o][1].y};
            if (EXTRAS && res) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    a[q] = V::round(a[q]);
                    if (ho < p2) a[q] += __ldg(res + (size_t)(c + q) * res_p + ho);
                    a[q] = V::round(a[q]);
                    if (ho >= H - p2) a[q] += __ldg(res + ((size_t)C + c + q) * res_p + (res_p - (H - ho)));
                }
            }
            const size_t off = base + (size_t)s * step;
            if (EXTRAS && add) {
                float av[4];
                V::unpack(*reinterpret_cast<const Raw*>(add + off), av);
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q] = V::round(a[q]) + av[q];
            }
            *reinterpret_cast<Raw*>(y + off) = V::pack(a);
        }
    }
}

template <typename T, int K, bool ALONG_W>
static int dw_strip_launch_t(const void* x, const void* w, const float* bias, const float* res, int res_p, const void* add, void* y,
                             int B, int H, int W, int C, cudaStream_t st) {
    constexpr int kSR = StripRun<K>::value;
    const int L = ALONG_W ? W : H, lines = ALONG_W ? H : W;
    const long long nitems = (long long)B * lines * ((L + kSR - 1) / kSR) * (C / 4);
    if (nitems >= (1LL << 31)) return -1;                     // caller falls back to the tiled kernel
    const int grid = (int)std::min<long long>((nitems + 255) / 256, 148LL * 16);
    if (res || add)
        dw_strip_kernel<T, K, ALONG_W, true><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, bias, res, res_p, (const T*)add, (T*)y, B, H, W, C, nitems);
    else        // the plain conv: no border-residue / running-sum code at all (predicated-off instructions still cost issue slots)
        dw_strip_kernel<T, K, ALONG_W, false><<<grid, 256, 0, st>>>((const T*)x, (const T*)w, bias, res, res_p, (const T*)add, (T*)y, B, H, W, C, nitems);
    CNA_LAUNCH_CHECK();
    return 0;
}

template <typename T, int KH, int KW>
static int dw_tile_launch_t(const void* x, const void* w, const float* bias, const float* res, int res_p, const void* add, void* y,
                            int B, int H, int W, int C, cudaStream_t st) {
    constexpr bool ALONG_W = KW > KH;
    constexpr int XH = kDT + KH - 1, XW = kDT + KW - 1;
    const size_t smem = ((size_t)XH * XW + (size_t)KH * KW) * (kDC / 2) * 2 * sizeof(T);
    auto kern = dw_tile_kernel<T, KH, KW, ALONG_W>;
    static std::atomic<unsigned long long> attr_done{0};
    CNA_CUDA(ensure_dyn_smem(kern, (int)smem, attr_done));
    const int tiles_h = ceil_div(H, kDT), tiles_w = ceil_div(W, kDT);
    const long long ntiles = (long long)B * tiles_h * tiles_w * (C / kDC);
    int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (216 * 1024) / (smem + 1024)));
    const int grid = (int)std::min<long long>(ntiles, 148LL * per_sm);
    kern<<<grid, 256, smem, st>>>((const T*)x, (const T*)w, bias, res, res_p, (const T*)add, (T*)y, B, H, W, C, tiles_h, tiles_w, ntiles);
    CNA_LAUNCH_CHECK();
    return 0;
}

template <typename T>
static int dw_tile_dispatch(int kh, int kw, const void* x, const void* w, const float* bias, const float* res, int res_p, const void* add,
                            void* y, int B, int H, int W, int C, cudaStream_t st, bool* handled) {
    *handled = true;
    if (std::is_same<T, __nv_bfloat16>::value && res == nullptr && add == nullptr && (kh == 1 || kw == 1)) {
        const int rc = dw_strip_tma_try(x, w, bias, y, B, H, W, C, kh, kw, st);      // TMA-fed shared-memory pipeline (dwstrip_tma.cu)
        if (rc >= 0) return rc;
    }
    if (std::is_same<T, __nv_bfloat16>::value && res == nullptr && add == nullptr && kh == kw && (kh == 3 || kh == 5) &&
        getenv("CNA_DW_NO_TMA_SQUARE") == nullptr) {
        const int rc = dw_square_tma_try(x, w, bias, y, B, H, W, C, kh, 0, 0, st);       // same pipeline, 3 / 5 box rows per input column
        if (rc >= 0) return rc;
    }
    if (getenv("CNA_DW_NO_STRIP") == nullptr) {      // (diagnostic switch: use the smem-tiled kernel for strips too)
#define CNA_STRIP_CASE(KV) \
        if (kh == 1 && kw == KV) { const int rc = dw_strip_launch_t<T, KV, true>(x, w, bias, res, res_p, add, y, B, H, W, C, st); if (rc >= 0) return rc; } \
        if (kw == 1 && kh == KV) { const int rc = dw_strip_launch_t<T, KV, false>(x, w, bias, res, res_p, add, y, B, H, W, C, st); if (rc >= 0) return rc; }
        CNA_STRIP_CASE(3)
        CNA_STRIP_CASE(5)
        CNA_STRIP_CASE(7)
        CNA_STRIP_CASE(11)
        CNA_STRIP_CASE(21)
#undef CNA_STRIP_CASE
    }
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

using namespace cna;

extern "C" {

int cna_msca_chain_supported(int C, int k1, int nb, const int* ks_host, int dtype) {
    if (C % kDC != 0 || k1 != 5 || nb < 1 || nb > 3 || (dtype != CNA_BF16 && dtype != CNA_F32)) return 0;
    for (int i = 0; i < nb; ++i)
        if (ks_host[i] != 3 && ks_host[i] != 5 && ks_host[i] != 7 && ks_host[i] != 11 && ks_host[i] != 21) return 0;
    return 1;
}

int cna_msca_chain_forward(const void* x, void* a0, void* t, void* s, const void* w0, const float* b0, const void* const* w1,
                           const float* const* b1, const void* const* w2, const float* const* b2, const int* ks_host, int nb,
                           int identity, const float* res, int res_p, int B, int H, int W, int C, int dtype, void* stream) {
    CNA_CHECK(x && a0 && t && s && w0 && w1 && w2 && b1 && b2 && ks_host, "null pointer");
    CNA_CHECK(cna_msca_chain_supported(C, 5, nb, ks_host, dtype), "configuration not supported (C % 32 == 0, k1 = 5, strips in {3,5,7,11,21})");
    cudaStream_t st = as_stream(stream);
    bool handled = false;
    auto run = [&](const void* in, const void* w, const float* bias, const float* rs, int rp, const void* add, void* out, int kh, int kw) -> int {
        handled = false;
        if (dw_tile_try(in, w, bias, rs, rp, add, out, B, H, W, C, 1, kh, kw, 1, 1, kh / 2, kw / 2, dtype, st, &handled)) return 1;
        CNA_CHECK(handled, "no tiled kernel for this depthwise shape");
        count_launch();
        return 0;
    };
    if (run(x, w0, b0, nullptr, 0, nullptr, a0, 5, 5)) return 1;                                   // conv0
    const bool fuse_res = res != nullptr && nb == 1 && !identity;                                 // CascadeConv + FixPaddingBias
    for (int i = 0; i < nb; ++i) {
        const int k = ks_host[i];
        if (run(a0, w1[i], b1[i], nullptr, 0, nullptr, t, 1, k)) return 1;                        // 1 x k
        if (run(t, w2[i], b2[i], fuse_res ? res : nullptr, fuse_res ? res_p : 0, i > 0 ? s : nullptr, s, k, 1)) return 1;   // k x 1 (+ running sum)
    }
    if (identity && cna_add(s, a0, s, (long long)B * H * W * C, dtype, stream)) return 1;         // ParallelConv's identity branch
    if (res != nullptr && !fuse_res && cna_fix_padding_bias(s, res, B, H, W, C, res_p, dtype, stream)) return 1;
    return 0;
}

}  // extern "C"
