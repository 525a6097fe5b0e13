// TMA-fed depthwise strip convolution (1 x K / K x 1, stride 1, 'same' padding, channel multiplier 1) for bf16 NHWC on sm_100a:
// the stand-alone form of the strips of CascadeConv / ParallelConv (reference layers/depth_seperable_conv.py:12-16, 45-47).
//
// A strip is 4 bytes of HBM traffic and K multiply-adds per element: HBM-bound on paper, but the register-window kernel in
// dwconv_tile.cu issues every load from a thread that then waits for it (long-scoreboard stalls 3.9 per issue, 0.51 of the HBM peak
// at best).  Here the loads are TMA boxes into a shared-memory ring, kept in flight by one producer warp, the conv padding is the
// tensor map's out-of-bounds zero fill, and the outputs leave through double-buffered TMA stores that clip at the image edge.
// (Storing the bf16 pairs straight from registers instead - 4 bytes per lane, two 64-byte pixel rows per warp store - was measured:
// 1x7 0.85 -> 0.66 of the HBM peak, 7x1 0.86 -> 0.56; half-line writes cost the L2 more than the staging costs the SM.)
//
//   tile      = one image, 32 channels, 8 rows x 32 columns (1 x K) or 32 rows x 8 columns (K x 1)
//   stage     = the tile plus its K-1 halo pixels along the strip axis (a 4-D box: channels, columns, rows, image)
//   thread    = one channel pair x one line of the tile; it slides along the strip axis with K accumulators in registers
//               (K packed FFMA2 per input value), rounds once and writes the bf16 pair into the store buffer
//   warp      = 16 channel pairs x 2 adjacent lines: a shared-memory access is two 64-byte pixel rows, in different bank halves
//               for the loads (the 1 x K box is padded to an odd number of columns)
//   4 compute warps + 1 producer warp, persistent CTAs (two per SM), tiles ordered channel chunk -> strip axis -> ... so that
//   the CTAs running side by side share their halos through L2.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <algorithm>
#include <mutex>

namespace cna {
using namespace ptx;

constexpr int kSC = 32;            // channels per tile
constexpr int kSLong = 32;         // outputs along the strip axis per tile
constexpr int kSMaxStages = 4;
// LINES = lines across the strip axis per tile = 2 per compute warp.  8 lines (4 compute warps + producer, two CTAs per SM) is what
// ships: with one 16-line CTA per SM all compute warps sit in the same phase, so the FP32 pipe and the memory system took turns
// (1x21: FMA pipe 48 % active, HBM 49 %, ncu profiles/r02_strip_tma_ncu.txt); two independent CTAs overlap each other's phases.

template <int LINES> struct StripCfg {
    static constexpr int kWarps = LINES / 2;                       // compute warps
    static constexpr int kThreads = (kWarps + 1) * 32;             // + producer warp
    static constexpr int kCtasPerSm = LINES == 8 ? 2 : 1;
    static constexpr uint32_t kOutBytes = kSLong * LINES * kSC * 2;   // store buffer
};

struct StripParams {
    int B, H, W, C;
    int t_long, t_lines, n_cc;     // tiles along the strip axis, across it, channel chunks
    int ntiles;
    int stages;
    uint32_t stage_bytes;
};

__device__ __forceinline__ float2 s_unpack(uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); }
__device__ __forceinline__ float2 s_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ uint32_t s_lds(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void s_sts(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// tile index -> (image, line tile, long tile, channel chunk); channel chunk fastest, then the strip axis
struct StripTile { int b, tl, tg, cc; };
__device__ __forceinline__ StripTile strip_decode(int tile, const StripParams& p) {
    StripTile t;
    t.cc = tile % p.n_cc; tile /= p.n_cc;
    t.tg = tile % p.t_long; tile /= p.t_long;
    t.tl = tile % p.t_lines; t.b = tile / p.t_lines;
    return t;
}

// K = taps along the axis a thread slides on, KX = taps across it (1 for the strips; 3 / 5 for the square 3x3 / 5x5 kernels, which
// slide along W and read KX box rows per input column), GELU = exact erf-GELU on the fp32 sum before the rounding (the depthwise conv +
// activation of MSCAN's FFN, reference models/mscan.py:58-59).
template <int K, int KX, bool ALONG_W, int LINES, bool GELU>
__global__ void __launch_bounds__(StripCfg<LINES>::kThreads, StripCfg<LINES>::kCtasPerSm)
dw_strip_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                    const __nv_bfloat16* __restrict__ w /*[KH*KW][C]*/, const float* __restrict__ bias, const StripParams p) {
    static_assert(KX == 1 || ALONG_W, "square kernels slide along W");
    using Cfg = StripCfg<LINES>;
    constexpr int kSLines = LINES;
    constexpr uint32_t kSOutBytes = Cfg::kOutBytes;
    constexpr int P = K / 2;
    constexpr int NIN = kSLong + K - 1;                          // input positions along the strip axis
    constexpr int PX = KX / 2;
    constexpr int BW = ALONG_W ? (kSLong + K) : kSLines;          // box columns (sliding along W: one spare column makes the count odd)
    // byte strides inside a stage / the store buffer ([rows][cols][32 ch] bf16, dense = TMA box order)
    constexpr uint32_t kPix = kSC * 2;                            // 64 bytes per pixel
    constexpr uint32_t in_long = ALONG_W ? kPix : BW * kPix;      // step along the strip axis
    constexpr uint32_t in_line = ALONG_W ? BW * kPix : kPix;      // step to the next line
    constexpr int OW = ALONG_W ? kSLong : kSLines;                // store-buffer columns
    constexpr uint32_t out_long = ALONG_W ? kPix : OW * kPix;
    constexpr uint32_t out_line = ALONG_W ? OW * kPix : kPix;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t out0 = base;                                   // 2 store buffers
    const uint32_t in0 = base + 2 * kSOutBytes;                   // p.stages input stages
    const uint32_t bars = in0 + p.stages * p.stage_bytes;         // full[stages], empty[stages]
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(kSMaxStages + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_x); prefetch_tmap(&tm_y);
        for (int s = 0; s < p.stages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), Cfg::kWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    const int first = blockIdx.x, step = gridDim.x;
    const int my_tiles = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;

    if (warp == Cfg::kWarps) {
        // ---------------- producer: one TMA box per tile ----------------
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const StripTile t = strip_decode(first + it * step, p);
            mbar_wait(empty(s), ph ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(full(s), p.stage_bytes);
                const int w0 = ALONG_W ? t.tg * kSLong - P : t.tl * kSLines;
                const int h0 = ALONG_W ? t.tl * kSLines - PX : t.tg * kSLong - P;
                tma_load_4d(in0 + s * p.stage_bytes, &tm_x, full(s), t.cc * kSC, w0, h0, t.b);
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else {
        // ---------------- compute warps ----------------
        const int pair = lane & 15;
        const int line = warp * 2 + (lane >> 4);                  // line of the tile this thread slides along
        const uint32_t in_thr = (uint32_t)line * in_line + (uint32_t)pair * 4u;
        const uint32_t out_thr = (uint32_t)line * out_line + (uint32_t)pair * 4u;
        int s = 0; uint32_t ph = 0;
        int cc_have = -1;
        float2 wt[KX * K];
        float2 bz = make_float2(0.f, 0.f);
        for (int it = 0; it < my_tiles; ++it) {
            const StripTile t = strip_decode(first + it * step, p);
            if (t.cc != cc_have) {                                // taps and bias of this thread's channel pair
                cc_have = t.cc;
                const int c = t.cc * kSC + pair * 2;
#pragma unroll
                for (int k = 0; k < KX * K; ++k) wt[k] = s_unpack(__ldg(reinterpret_cast<const uint32_t*>(w + (size_t)k * p.C + c)));
                if (bias) bz = __ldg(reinterpret_cast<const float2*>(bias + c));
            }
            const uint32_t ob = out0 + (uint32_t)(it & 1) * kSOutBytes;
            // the store buffer was handed to a TMA store two tiles ago
            if (threadIdx.x == 0) tma_store_wait_read<1>();
            named_bar_sync(1, Cfg::kWarps * 32);
            mbar_wait(full(s), ph);
            const uint32_t src = in0 + s * p.stage_bytes + in_thr;
            float2 acc[K];
#pragma unroll
            for (int k = 0; k < K; ++k) acc[k] = bz;
            // the shared-memory loads run kAhead input positions ahead of their use (they are volatile asm, i.e. kept in program order
            // against the stores and barriers: issued right before their first use each one would expose its ~30-cycle latency)
#ifndef CNA_STRIP_AHEAD_SMALLK
#define CNA_STRIP_AHEAD_SMALLK 0
#endif
            constexpr int kAhead = (KX * K >= 9) ? 3 : CNA_STRIP_AHEAD_SMALLK;     // short strips (K <= 7): measured per K, see profiles/r02_strip_tma.txt
            uint32_t q[kAhead + 1][KX];
#pragma unroll
            for (int i = 0; i < (kAhead > 0 ? kAhead : 1) && i < NIN; ++i)
#pragma unroll
                for (int rx = 0; rx < KX; ++rx) q[i][rx] = s_lds(src + (uint32_t)rx * in_line + (uint32_t)i * in_long);
#pragma unroll
            for (int i = 0; i < NIN; ++i) {
                if (i + kAhead < NIN && (kAhead > 0 || i > 0)) {
#pragma unroll
                    for (int rx = 0; rx < KX; ++rx)
                        q[(i + kAhead) % (kAhead + 1)][rx] = s_lds(src + (uint32_t)rx * in_line + (uint32_t)(i + kAhead) * in_long);
                }
#pragma unroll
                for (int rx = 0; rx < KX; ++rx) {                 // the KX box rows this line's window covers
                    const float2 xv = s_unpack(q[i % (kAhead + 1)][rx]);
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int o = i - k;                      // output fed by tap (rx, k) of input i
                        if (o >= 0 && o < kSLong) acc[o % K] = s_ffma2(xv, wt[rx * K + k], acc[o % K]);
                    }
                }
                if (i >= K - 1) {                                 // output i-K+1 has all its taps
                    const int o = i - (K - 1);
                    if constexpr (GELU) {
                        acc[o % K] = gelu_erf_fast2(acc[o % K].x, acc[o % K].y);
                    }
                    __nv_bfloat162 r = __floats2bfloat162_rn(acc[o % K].x, acc[o % K].y);
                    s_sts(ob + out_thr + (uint32_t)o * out_long, *reinterpret_cast<uint32_t*>(&r));
                    acc[o % K] = bz;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(s));                 // stage consumed
            if (++s == p.stages) { s = 0; ph ^= 1u; }
            fence_proxy_async_smem();
            named_bar_sync(2, Cfg::kWarps * 32);
            if (threadIdx.x == 0) {
                const int w0 = ALONG_W ? t.tg * kSLong : t.tl * kSLines;
                const int h0 = ALONG_W ? t.tl * kSLines : t.tg * kSLong;
                tma_store_4d(&tm_y, ob, t.cc * kSC, w0, h0, t.b);
                tma_store_commit();
            }
        }
        if (threadIdx.x == 0) tma_store_wait_all<0>();
    }
}

typedef CUresult (*EncodeTiledFnS)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnS strip_get_encode() {
    static EncodeTiledFnS fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnS>(f);
    });
    return fn;
}

static int strip_encode(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, uint32_t box_w, uint32_t box_h) {
    EncodeTiledFnS fn = strip_get_encode();
    CNA_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kSC, box_w, box_h, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CNA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

template <int K, int KX, bool ALONG_W, int LINES, bool GELU>
static int strip_tma_launch(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, cudaStream_t st) {
    DevInfo di;
    if (dev_info(&di)) return 1;
    using Cfg = StripCfg<LINES>;
    constexpr int kSLines = LINES;
    constexpr uint32_t kSOutBytes = Cfg::kOutBytes;
    StripParams p{};
    p.B = B; p.H = H; p.W = W; p.C = C;
    p.n_cc = C / kSC;
    p.t_long = ceil_div(ALONG_W ? W : H, kSLong);
    p.t_lines = ceil_div(ALONG_W ? H : W, kSLines);
    const long long ntiles = (long long)B * p.t_long * p.t_lines * p.n_cc;
    if (ntiles >= (1LL << 31)) return -1;
    p.ntiles = (int)ntiles;
    const uint32_t box_w = ALONG_W ? kSLong + K : kSLines, box_h = ALONG_W ? kSLines + KX - 1 : kSLong + K - 1;
    p.stage_bytes = (box_w * box_h * kSC * 2 + 127u) & ~127u;
    const size_t fixed = 2 * (size_t)kSOutBytes + 1024 /*barriers*/ + 1024 /*alignment slack*/;
    const size_t budget = Cfg::kCtasPerSm == 1 ? (size_t)di.max_smem : ((size_t)di.max_smem - 2048) / 2;   // 1 KB per CTA is reserved by the system
    p.stages = (int)std::min<size_t>(kSMaxStages, (budget - fixed) / p.stage_bytes);
    if (p.stages < 2) return -1;
    const size_t smem = fixed + (size_t)p.stages * p.stage_bytes;
    CUtensorMap tm_x, tm_y;
    if (strip_encode(&tm_x, x, B, H, W, C, box_w, box_h)) return 1;
    if (strip_encode(&tm_y, y, B, H, W, C, ALONG_W ? kSLong : kSLines, ALONG_W ? kSLines : kSLong)) return 1;
    auto kern = dw_strip_tma_kernel<K, KX, ALONG_W, LINES, GELU>;
    static std::atomic<unsigned long long> attr_done{0};
    CNA_CUDA(ensure_dyn_smem(kern, di.max_smem, attr_done));
    const int grid = std::min(p.ntiles, di.num_sms * Cfg::kCtasPerSm);
    kern<<<grid, Cfg::kThreads, smem, st>>>(tm_x, tm_y, (const __nv_bfloat16*)w, bias, p);
    CNA_LAUNCH_CHECK();
    return 0;
}

// Square 3x3 / 5x5 depthwise kernels on the same pipeline, optionally with GELU.  Same return convention as dw_strip_tma_try.
int dw_square_tma_try(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, int k, int gelu, int force,
                      cudaStream_t st) {
    if (C % kSC != 0 || getenv("CNA_DW_NO_TMA") != nullptr) return -1;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) return -1;
    if (W <= 16 && !force && getenv("CNA_DW_FORCE_TMA") == nullptr) return -1;        // half-empty tiles: the smem-tiled kernel is faster (16 x 16 images)
    if (k == 3) return gelu ? strip_tma_launch<3, 3, true, 8, true>(x, w, bias, y, B, H, W, C, st) : strip_tma_launch<3, 3, true, 8, false>(x, w, bias, y, B, H, W, C, st);
    if (k == 5 && !gelu) return strip_tma_launch<5, 5, true, 8, false>(x, w, bias, y, B, H, W, C, st);
    return -1;
}

// bf16 strips without fused extras.  Returns 0 when launched, -1 when this kernel does not take the configuration (the caller
// falls back to the register-window kernel), 1 on error.
int dw_strip_tma_try(const void* x, const void* w, const float* bias, void* y, int B, int H, int W, int C, int kh, int kw, cudaStream_t st) {
    if (C % kSC != 0 || getenv("CNA_DW_NO_TMA") != nullptr) return -1;
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) return -1;
    // a strip axis of <= 16 positions leaves half of every 32-long tile empty: the register-window kernel is faster there
    // (16 x 16 images at C = 256: 0.47 against 0.38 of the HBM peak, profiles/r02_strip_tma.txt)
    if ((kh == 1 ? W : H) <= 16 && getenv("CNA_DW_FORCE_TMA") == nullptr) return -1;
    static const bool wide = getenv("CNA_STRIP_LINES16") != nullptr;      // diagnostic: the 16-line, one-CTA-per-SM configuration
#define CNA_STRIP_TMA_CASE(KV) \
    if (kh == 1 && kw == KV) return wide ? strip_tma_launch<KV, 1, true, 16, false>(x, w, bias, y, B, H, W, C, st) : strip_tma_launch<KV, 1, true, 8, false>(x, w, bias, y, B, H, W, C, st); \
    if (kw == 1 && kh == KV) return wide ? strip_tma_launch<KV, 1, false, 16, false>(x, w, bias, y, B, H, W, C, st) : strip_tma_launch<KV, 1, false, 8, false>(x, w, bias, y, B, H, W, C, st);
    CNA_STRIP_TMA_CASE(3)
    CNA_STRIP_TMA_CASE(5)
    CNA_STRIP_TMA_CASE(7)
    CNA_STRIP_TMA_CASE(11)
    CNA_STRIP_TMA_CASE(21)
#undef CNA_STRIP_TMA_CASE
    return -1;
}

}  // namespace cna

using namespace cna;

extern "C" int cna_dwconv3x3_act_forward(const void* x, const void* w_packed, const float* bias, void* y, int B, int H, int W, int C, int act,
                                         void* stream) {
    CNA_CHECK(x && w_packed && y && B > 0 && H > 0 && W > 0, "bad arguments");
    CNA_CHECK(act == 0 || act == 2, "activation must be 0 (none) or 2 (GELU)");
    CNA_CHECK(C % kSC == 0, "needs C % 32 == 0");
    const int rc = dw_square_tma_try(x, w_packed, bias, y, B, H, W, C, 3, act == 2, 1, as_stream(stream));
    CNA_CHECK(rc >= 0, "configuration not taken by the TMA depthwise kernel (16-byte aligned tensors, C % 32 == 0)");
    if (rc) return rc;
    count_launch();
    return 0;
}

