// Fused scheme-2 low-rank convolution for sm_100a:  y = h_conv(v_conv(x)) + bias  in ONE kernel.
//
//   T[b,h,w,k] = sum_dy sum_c x[b,h+dy-1,w,c] * Wv[dy][k][c]          GEMM1  (implicit GEMM, K-dim = 3*C)
//   y[b,h,w,n] = sum_dx sum_k T[b,h,w+dx-1,k] * Wh[dx][n][k] + b[n]   GEMM2  (implicit GEMM, K-dim = 3*K)
//
// Replaces LowRankExpConvV2.forward (reference src/approx/layers/low_rank_conv.py:83-84), where the
// K-channel intermediate T is a full HBM round trip between two cuDNN launches.  Here T never leaves
// the SM: GEMM1 accumulates in TMEM, the epilogue warps round it to bf16 (exactly where the reference's
// bf16 modules materialise the v_conv output) into shared memory, and GEMM2 reads it from there.
//
// Tiling.  One tile = 8 output rows x 14 output columns of one image, handled as a 128-slot UMMA
// M-tile of 8 rows x 16 columns (one halo column each side, recomputed by the neighbouring tile).
//   * x tile (TMA, 4-D NHWC map, OOB zero fill = conv padding): 10 rows x 16 cols x CK channels,
//     slot order (row, col), col fastest.  The three vertical taps of GEMM1 are the SAME smem tile
//     read through UMMA descriptors whose start address is shifted by 16 slots (one row): a
//     multiple of the 8-row swizzle atom, so no data movement and no re-load per tap.
//   * T tile (written by the epilogue warps): slot order (col, row), ROW fastest, 18 cols x 8 rows.
//     The three horizontal taps of GEMM2 are then descriptor shifts of 8 slots (one column) — again
//     swizzle-atom aligned.  The transposition of slot order costs nothing: it is just the address
//     each epilogue thread stores its TMEM lane to.
//   * y tile: TMEM -> registers (+bias, ReLU) -> swizzled smem staging -> TMA store (clips at edges).
//
// Pipeline (warp specialised, persistent CTAs, 192 threads):
//   warp 0     TMA producer   x/Wv chunk ring (S stages), Wh ring or resident weights
//   warp 1     MMA issuer     order G1(0) G1(1) G2(0) G1(2) G2(1) ...  (GEMM1 of tile i+1 is in flight
//                              while the epilogue converts tile i's T, GEMM2 of tile i while it stores i-1)
//   warps 2-5  epilogue       T-epilogue(i), Y-epilogue(i-1), ...   TMEM lanes (warp%4)*32..+31
// TMEM: 2 x 64-col T accumulators + 2 x 64-col Y accumulators = 256 columns.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <map>
#include <mutex>
#include <tuple>

namespace cna {
using namespace ptx;

constexpr int kThreads = 480;        // warp 0 TMA, 1 GEMM1 issuer, 2-5 T-epilogue, 6-9 Y-epilogue group 0, 10 GEMM2 issuer, 11-14 Y-epilogue group 1
constexpr int kTileRows = 8;           // output rows per tile
constexpr int kTileCols = 16;          // computed columns per tile (14 valid + 2 halo)
constexpr int kValidCols = kTileCols - 2;
constexpr int kXRows = kTileRows + 2;  // x rows per tile (d = 3)
constexpr int kTCols = kTileCols + 2;  // T columns kept in smem
constexpr int kNT = 64;                // GEMM2 N-tile (output channels per accumulator)
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 6;
constexpr int kMaxNTiles = 16;         // N <= 1024
// T lives in the NO-swizzle K-major UMMA layout: 8-slot groups of 16-byte chunks, group pitch kTSbo, chunk-plane
// pitch kTLbo.  The 144-byte group pitch (instead of 128) makes the transposing epilogue stores — 8 lanes that
// hit 8 consecutive groups at the same in-group row — fall into 8 different bank quads.
constexpr int kTraceEv = 10;
constexpr uint32_t kTSbo = 144;
constexpr uint32_t kTLbo = kTCols * kTSbo;   // 2592

struct LR2Params {
    int B, H, W, N;
    int nch;               // number of channel chunks = c_pad / CK
    int nnt;               // N / 64
    int tiles_w, tiles_h;  // tile grid per image
    int ntiles;
    int stages;            // x ring depth
    int resident;          // 1: all weights stay in smem for the whole kernel; 0: streamed per tile
    int relu;
    int pool;              // fuse a 2x2/stride-2 max-pool into the epilogue (y is then [B, H/2, W/2, N])
    // smem byte offsets from the 1024-aligned base
    uint32_t off_x, x_stage;      // x ring
    uint32_t off_v, v_chunk;      // Wv chunks (resident: nch of them; streamed: one per x stage)
    uint32_t off_h, h_tile;       // Wh N-tiles (resident: nnt; streamed: 2)
    uint32_t off_t, t_buf;        // T buffers (2)
    uint32_t off_y, y_buf;        // y staging (2)
    uint32_t off_p;               // pooled staging (2 x 4 KB), only when pool
    uint32_t off_bias, off_bar;
    int debug;             // diagnostic bit mask (CNA_LR2_DEBUG): see lr2_launch; 0 in production
    long long* trace;      // diagnostic: per-tile clock64() stamps of CTA 0 (NULL = off), kTraceEv slots per tile
    int trace_tiles, trace_skip;
};

__host__ __device__ constexpr uint32_t swz_type(int row_bytes) { return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// barrier indices (each 8 bytes) inside the barrier block
enum : int {
    BAR_XFULL = 0,                         // [kMaxStages]
    BAR_XEMPTY = BAR_XFULL + kMaxStages,   // [kMaxStages]
    BAR_WFULL = BAR_XEMPTY + kMaxStages,   // resident weights landed
    BAR_HFULL = BAR_WFULL + 1,             // [2] streamed Wh ring
    BAR_HEMPTY = BAR_HFULL + 2,            // [2]
    BAR_TACC_FULL = BAR_HEMPTY + 2,        // [2]
    BAR_TACC_EMPTY = BAR_TACC_FULL + 2,    // [2]
    BAR_TSM_FULL = BAR_TACC_EMPTY + 2,     // [2]
    BAR_TSM_EMPTY = BAR_TSM_FULL + 2,      // [2]
    BAR_YACC_FULL = BAR_TSM_EMPTY + 2,     // [2]
    BAR_YACC_EMPTY = BAR_YACC_FULL + 2,    // [2]
    BAR_COUNT = BAR_YACC_EMPTY + 2
};

// 64-bit UMMA shared-memory descriptor with the start-address field left 0; add (addr >> 4) per operand.
__host__ __device__ constexpr uint64_t desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7) << 61);
}

struct TileCoord { int b, h0, w0; };
__device__ __forceinline__ TileCoord decode_tile(int tile, const LR2Params& p) {
    int per_img = p.tiles_w * p.tiles_h;
    int b = tile / per_img, rem = tile - b * per_img;
    int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
    return {b, th * kTileRows, tw * kValidCols};
}

template <int CK, int KP>
__global__ void __launch_bounds__(kThreads, 1)
lowrank2_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wv,
                      const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_y,
                      const float* __restrict__ bias, const LR2Params p) {   // tm_y describes the pooled tensor when p.pool
    constexpr int XROW = CK * 2;            // bytes per x / Wv smem row (== swizzle span)
    constexpr int HROW = KP * 2;            // bytes per Wh smem row
    constexpr int TCH = KP / 8;             // 16-byte chunks per T slot
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + p.off_bar;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    float* sbias = reinterpret_cast<float*>(base_ptr + p.off_bias);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + p.off_bar + 8 * BAR_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---------------- one-time setup ----------------
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_x); prefetch_tmap(&tm_wv); prefetch_tmap(&tm_wh); prefetch_tmap(&tm_y);
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(bar(BAR_XFULL + i), 1); mbar_init(bar(BAR_XEMPTY + i), 1); }
        mbar_init(bar(BAR_WFULL), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(BAR_HFULL + i), 1);      mbar_init(bar(BAR_HEMPTY + i), 1);
            mbar_init(bar(BAR_TACC_FULL + i), 1);  mbar_init(bar(BAR_TACC_EMPTY + i), 128);
            mbar_init(bar(BAR_TSM_FULL + i), 128); mbar_init(bar(BAR_TSM_EMPTY + i), 1);
            mbar_init(bar(BAR_YACC_FULL + i), 1);  mbar_init(bar(BAR_YACC_EMPTY + i), 128);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    for (int i = threadIdx.x; i < p.N; i += kThreads) sbias[i] = bias ? bias[i] : 0.f;
    // T column groups 0 and 17 are never produced (they only feed the two discarded halo outputs): keep them finite.
    for (int i = threadIdx.x; i < 2 * 2 * TCH * 8; i += kThreads) {
        int row = i & 7, ch = (i >> 3) % TCH, g = ((i >> 3) / TCH) & 1, bsel = (i >> 3) / TCH / 2;
        st_shared_v4(base + p.off_t + bsel * p.t_buf + ch * kTLbo + (g ? (kTCols - 1) : 0) * kTSbo + row * 16, 0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_T = tmem;            // + buf*64
    const uint32_t tmem_Y = tmem + 128;      // + buf*64

    // diagnostic: row `trace_tiles` of the trace holds {clock64, globaltimer} at CTA 0's start and end (effective SM clock)
    auto stamp_wall = [&](int slot) {
        if (p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            p.trace[p.trace_tiles * kTraceEv + slot] = clock64();
            p.trace[p.trace_tiles * kTraceEv + slot + 1] = (long long)gt;
        }
    };
    stamp_wall(0);
    const int first = blockIdx.x, step = gridDim.x;
    const int my_tiles = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;
    auto stamp = [&](int it, int ev) {
        if (p.trace != nullptr && blockIdx.x == 0 && it >= p.trace_skip && it < p.trace_skip + p.trace_tiles)
            p.trace[(it - p.trace_skip) * kTraceEv + ev] = clock64();
    };

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (my_tiles > 0) {
            constexpr uint32_t x_bytes = kXRows * kTileCols * XROW;
            constexpr uint32_t v_bytes = 3 * KP * XROW;
            constexpr uint32_t h_bytes = 3 * kNT * HROW;
            if (p.resident && elect_one()) {
                mbar_arrive_expect_tx(bar(BAR_WFULL), v_bytes * p.nch + h_bytes * p.nnt);
                for (int ch = 0; ch < p.nch; ++ch)
                    tma_load_2d(base + p.off_v + ch * p.v_chunk, &tm_wv, bar(BAR_WFULL), ch * CK, 0);
                for (int nt = 0; nt < p.nnt; ++nt)
                    for (int dx = 0; dx < 3; ++dx)
                        tma_load_2d(base + p.off_h + nt * p.h_tile + dx * (kNT * HROW), &tm_wh, bar(BAR_WFULL), 0,
                                    dx * p.N + nt * kNT);
            }
            __syncwarp();
            int xs = 0; uint32_t xph = 0;   // x ring stage / phase
            int hs = 0; uint32_t hph = 0;   // Wh ring stage / phase
            auto load_x = [&](int it) {
                TileCoord t = decode_tile(first + it * step, p);
                for (int ch = 0; ch < p.nch; ++ch) {
                    if ((p.debug & 16) && it * p.nch + ch >= p.stages) continue;
                    mbar_wait(bar(BAR_XEMPTY + xs), xph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(BAR_XFULL + xs), x_bytes + (p.resident ? 0u : v_bytes));
                        tma_load_4d(base + p.off_x + xs * p.x_stage, &tm_x, bar(BAR_XFULL + xs), ch * CK, t.w0 - 1, t.h0 - 1, t.b);
                        if (!p.resident)
                            tma_load_2d(base + p.off_v + xs * p.v_chunk, &tm_wv, bar(BAR_XFULL + xs), ch * CK, 0);
                    }
                    __syncwarp();
                    if (++xs == p.stages) { xs = 0; xph ^= 1u; }
                }
                if (lane == 0) stamp(it, 0);
            };
            auto load_h = [&]() {
                if (p.resident) return;
                for (int nt = 0; nt < p.nnt; ++nt) {
                    mbar_wait(bar(BAR_HEMPTY + hs), hph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(BAR_HFULL + hs), h_bytes);
                        for (int dx = 0; dx < 3; ++dx)
                            tma_load_2d(base + p.off_h + hs * p.h_tile + dx * (kNT * HROW), &tm_wh, bar(BAR_HFULL + hs), 0,
                                        dx * p.N + nt * kNT);
                    }
                    __syncwarp();
                    if (++hs == 2) { hs = 0; hph ^= 1u; }
                }
            };
            // same order as the MMA warp consumes: x(0) | x(1) h(0) | x(2) h(1) | ...
            load_x(0);
            for (int it = 0; it < my_tiles; ++it) {
                if (it + 1 < my_tiles) load_x(it + 1);
                load_h();
            }
        }
    } else if (warp == 1) {
        // =========================== GEMM1 issuer ===========================
        // Two issuing warps (GEMM1 here, GEMM2 in warp 10): UTCHMMA issue blocks while the tensor pipe's short queue
        // is full, so a single issuer would expose every mbarrier wait (~100-200 cycles each, ~1350 per tile measured
        // with benchmarks/trace_lowrank2.py) as pipe idle time; with two, one warp's waits overlap the other's MMAs.
        if (my_tiles > 0) {
            constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);
            constexpr uint64_t xdesc = desc_hi(16, 8 * XROW, swz_type(XROW));     // x tile and Wv chunk (swizzled, K-major)
            int xs = 0; uint32_t xph = 0;
            if (p.resident) { mbar_wait(bar(BAR_WFULL), 0); tc_fence_after_sync(); }
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                if (lane == 0) stamp(it, 1);
                mbar_wait(bar(BAR_TACC_EMPTY + buf), (uint32_t)((it >> 1) & 1) ^ 1u);
                tc_fence_after_sync();
                const uint32_t d = tmem_T + buf * 64;
                for (int ch = 0; ch < p.nch; ++ch) {
                    if (!((p.debug & 16) && it * p.nch + ch >= p.stages)) mbar_wait(bar(BAR_XFULL + xs), xph);
                    tc_fence_after_sync();
                    if (lane == 0 && ch == 0) stamp(it, 2);
                    const uint64_t a0 = xdesc + ((base + p.off_x + xs * p.x_stage) >> 4);
                    const uint64_t b0 = xdesc + ((base + p.off_v + (p.resident ? ch : xs) * p.v_chunk) >> 4);
                    if (elect_one()) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int kk = 0; kk < CK / 16; ++kk)
                                umma_bf16(d, a0 + ((dy * kTileCols * XROW + kk * 32) >> 4), b0 + ((dy * KP * XROW + kk * 32) >> 4),
                                          idesc1, (uint32_t)((ch | dy | kk) != 0));
                        if (!(p.debug & 16)) umma_commit(bar(BAR_XEMPTY + xs));
                        if (ch == p.nch - 1) umma_commit(bar(BAR_TACC_FULL + buf));
                    }
                    __syncwarp();
                    if (++xs == p.stages) { xs = 0; xph ^= 1u; }
                }
                if (lane == 0) stamp(it, 3);
            }
        }
    } else if (warp == 10) {
        // =========================== GEMM2 issuer ===========================
        if (my_tiles > 0) {
            constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kNT >> 3) << 17) | (8u << 24);
            constexpr uint64_t hdesc = desc_hi(16, 8 * HROW, swz_type(HROW));     // Wh tile
            constexpr uint64_t tdesc = desc_hi(kTLbo, kTSbo, 0);                  // T: no-swizzle K-major
            int q = 0;   // N-tile counter (Y accumulator / Wh ring index)
            if (p.resident) { mbar_wait(bar(BAR_WFULL), 0); tc_fence_after_sync(); }
            for (int j = 0; j < my_tiles; ++j) {
                const int buf = j & 1;
                if (lane == 0) stamp(j, 4);
                mbar_wait(bar(BAR_TSM_FULL + buf), (uint32_t)((j >> 1) & 1));
                tc_fence_after_sync();
                if (lane == 0) stamp(j, 5);
                const uint64_t a0 = tdesc + ((base + p.off_t + buf * p.t_buf) >> 4);
                for (int nt = 0; nt < p.nnt; ++nt, ++q) {
                    const int yb = q & 1;
                    const uint32_t ph = (uint32_t)((q >> 1) & 1);
                    mbar_wait(bar(BAR_YACC_EMPTY + yb), ph ^ 1u);
                    uint32_t hb;
                    if (p.resident) {
                        hb = base + p.off_h + nt * p.h_tile;
                    } else {
                        mbar_wait(bar(BAR_HFULL + yb), ph);
                        hb = base + p.off_h + yb * p.h_tile;
                    }
                    tc_fence_after_sync();
                    const uint64_t b0 = hdesc + (hb >> 4);
                    const uint32_t d = tmem_Y + yb * 64;
                    if (elect_one()) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                            for (int kk = 0; kk < KP / 16; ++kk)
                                umma_bf16(d, a0 + ((dx * kTSbo + kk * 2 * kTLbo) >> 4), b0 + ((dx * (kNT * HROW) + kk * 32) >> 4),
                                          idesc2, (uint32_t)((dx | kk) != 0));
                        if (!p.resident) umma_commit(bar(BAR_HEMPTY + yb));
                        umma_commit(bar(BAR_YACC_FULL + yb));
                        if (nt == p.nnt - 1) umma_commit(bar(BAR_TSM_EMPTY + buf));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp < 6) {
        // =========================== T-epilogue warps (2..5) ===========================
        // TMEM accumulator of GEMM1 -> bf16 -> shared memory, transposing slot order (row,col) -> (col,row).
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const int m = lg * 32 + lane;            // accumulator row: m = r*16 + xc
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const int r = m >> 4, xc = m & 15;
        const uint32_t slot_off = (uint32_t)(xc + 1) * kTSbo + (uint32_t)r * 16;
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            mbar_wait(bar(BAR_TACC_FULL + buf), ph);
            tc_fence_after_sync();
            if (threadIdx.x == 64) stamp(it, 6);
            uint32_t v[KP];
#pragma unroll
            for (int c16 = 0; c16 < KP / 16; ++c16) tmem_ld16(tmem_T + buf * 64 + lane_addr + c16 * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c16 * 16]));
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(bar(BAR_TACC_EMPTY + buf));          // accumulator drained: GEMM1 of tile it+2 may start
            mbar_wait(bar(BAR_TSM_EMPTY + buf), ph ^ 1u);     // GEMM2 of tile it-2 has finished reading this T buffer
            const uint32_t tb = base + p.off_t + buf * p.t_buf + slot_off;
            if (!(p.debug & 4))
#pragma unroll
            for (int ch = 0; ch < TCH; ++ch)
                st_shared_v4(tb + ch * kTLbo,
                             pack_bf16(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1])),
                             pack_bf16(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3])),
                             pack_bf16(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5])),
                             pack_bf16(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7])));
            if (!(p.debug & 8)) fence_proxy_async_smem();
            mbar_arrive(bar(BAR_TSM_FULL + buf));
            if (threadIdx.x == 64) stamp(it, 7);
        }
    } else if (warp < 10 || warp >= 11) {
        // =========================== Y-epilogue warps: group 0 = warps 6..9, group 1 = warps 11..14 ===========================
        // TMEM accumulator of GEMM2 -> +bias, ReLU -> bf16 -> swizzled staging -> TMA store.  Group g owns Y accumulator g,
        // staging buffer g and every N-tile q with (q & 1) == g, so two N-tiles are in the epilogue at any time: one
        // group's ld -> math -> st -> fence -> store chain (~1300 cycles) would otherwise bound the whole pipeline.
        const int g = warp >= 11 ? 1 : 0;
        const int lg = warp & 3;
        const int m = lg * 32 + lane;            // accumulator row: m = yc*8 + r
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const int et = (warp - (g ? 11 : 6)) * 32 + lane;   // 0..127 inside the group
        const uint32_t bar_a = 1 + 3 * g, bar_b = 2 + 3 * g, bar_c = 3 + 3 * g;
        const int yc = m >> 3, r = m & 7;
        const bool valid = yc >= 1 && yc <= kValidCols;
        // staging rows in TMA box order (col, row) — the y tensor map lists H before W — so the 8 lanes of a store
        // phase hit 8 consecutive 128-byte rows: with the 128B swizzle (chunk ^= row & 7) that is conflict free.
        const uint32_t srow = (uint32_t)((yc - 1) * kTileRows + r);
        const uint32_t row_off = srow * 128u, sw = (uint32_t)r;
        const float lo = p.relu ? 0.f : -INFINITY;
        const uint32_t sb = base + p.off_y + g * p.y_buf;
        const uint32_t pb = base + p.off_p + g * 4096;
        const uint32_t tacc = tmem_Y + g * 64 + lane_addr;
        const int nq = my_tiles * p.nnt;
        int j = 0, nt = g;                       // q = j * nnt + nt
        while (nt >= p.nnt) { nt -= p.nnt; ++j; }
        for (int q = g; q < nq; q += 2) {
            TileCoord t = decode_tile(first + j * step, p);
            mbar_wait(bar(BAR_YACC_FULL + g), (uint32_t)((q >> 1) & 1));
            tc_fence_after_sync();
            if (et == 0 && nt == 0) stamp(j, 8);
            uint32_t v[kNT];
#pragma unroll
            for (int c16 = 0; c16 < kNT / 16; ++c16) tmem_ld16(tacc + c16 * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c16 * 16]));
            tmem_ld_wait();
            tc_fence_before_sync();
            mbar_arrive(bar(BAR_YACC_EMPTY + g));        // accumulator drained: GEMM2 of N-tile q+2 may start
            uint32_t w[kNT / 2];
            {
                const float* bz = sbias + nt * kNT;
#pragma unroll
                for (int ch = 0; ch < kNT / 8; ++ch) {
                    float4 b0 = ld_shared_f4(bz + ch * 8), b1 = ld_shared_f4(bz + ch * 8 + 4);
                    w[ch * 4 + 0] = pack_bf16(fmaxf(__uint_as_float(v[ch * 8 + 0]) + b0.x, lo), fmaxf(__uint_as_float(v[ch * 8 + 1]) + b0.y, lo));
                    w[ch * 4 + 1] = pack_bf16(fmaxf(__uint_as_float(v[ch * 8 + 2]) + b0.z, lo), fmaxf(__uint_as_float(v[ch * 8 + 3]) + b0.w, lo));
                    w[ch * 4 + 2] = pack_bf16(fmaxf(__uint_as_float(v[ch * 8 + 4]) + b1.x, lo), fmaxf(__uint_as_float(v[ch * 8 + 5]) + b1.y, lo));
                    w[ch * 4 + 3] = pack_bf16(fmaxf(__uint_as_float(v[ch * 8 + 6]) + b1.z, lo), fmaxf(__uint_as_float(v[ch * 8 + 7]) + b1.w, lo));
                }
            }
            // this group's staging buffer was handed to a TMA store one item ago: wait until it has been read
            if (et == 0) tma_store_wait_read<0>();
            named_bar_sync(bar_a, 128);
            if (valid && !(p.debug & 2)) {
#pragma unroll
                for (int ch = 0; ch < kNT / 8; ++ch)
                    st_shared_v4(sb + row_off + (((uint32_t)ch ^ sw) << 4), w[ch * 4], w[ch * 4 + 1], w[ch * 4 + 2], w[ch * 4 + 3]);
            }
            if (!p.pool) {
                if (!(p.debug & 8)) fence_proxy_async_smem();
                named_bar_sync(bar_b, 128);
                if (et == 0 && !(p.debug & 1)) {
                    tma_store_4d(&tm_y, sb, nt * kNT, t.h0, t.w0, t.b);
                    tma_store_commit();
                    if (nt == p.nnt - 1) stamp(j, 9);
                }
            } else {
                // 2x2 max-pool of the staged 8 x 14 tile -> 4 x 7 pooled tile (bf16 max commutes with the rounding)
                named_bar_sync(bar_b, 128);
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int id = et + rep * 128;            // 28 pooled pixels x 8 chunks of 16 bytes
                    if (id < 28 * 8) {
                        const int pp = id >> 3, ch = id & 7, pc = pp >> 2, pr = pp & 3;     // pooled row index pp = pc*4 + pr
                        uint4 m4 = make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int bq = 0; bq < 2; ++bq) {
                                const uint32_t row = (uint32_t)((2 * pc + bq) * kTileRows + 2 * pr + a);
                                uint4 v4;
                                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(v4.x), "=r"(v4.y), "=r"(v4.z), "=r"(v4.w)
                                             : "r"(sb + row * 128u + (((uint32_t)ch ^ (row & 7u)) << 4)));
                                if (a == 0 && bq == 0) m4 = v4;
                                else {
                                    __nv_bfloat162* mm = reinterpret_cast<__nv_bfloat162*>(&m4);
                                    const __nv_bfloat162* vv = reinterpret_cast<const __nv_bfloat162*>(&v4);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) mm[e] = __hmax2(mm[e], vv[e]);
                                }
                            }
                        st_shared_v4(pb + (uint32_t)pp * 128u + (((uint32_t)ch ^ ((uint32_t)pp & 7u)) << 4), m4.x, m4.y, m4.z, m4.w);
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_c, 128);
                if (et == 0 && !(p.debug & 1)) {
                    tma_store_4d(&tm_y, pb, nt * kNT, t.h0 >> 1, t.w0 >> 1, t.b);
                    tma_store_commit();
                    if (nt == p.nnt - 1) stamp(j, 9);
                }
            }
            nt += 2;
            while (nt >= p.nnt) { nt -= p.nnt; ++j; }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    // ---------------- teardown ----------------
    tc_fence_before_sync();
    __syncthreads();
    stamp_wall(2);
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<kTmemCols>(tmem);
    }
}

// v_w (K,C,d,1) fp32 -> wv[d][k_pad][c_pad] bf16 ; h_w (N,K,1,d) fp32 -> wh[d][N][k_pad] bf16
__global__ void lowrank2_pack_kernel(const float* __restrict__ v_w, const float* __restrict__ h_w,
                                     __nv_bfloat16* __restrict__ wv, __nv_bfloat16* __restrict__ wh,
                                     int C, int K, int N, int d, int c_pad, int k_pad) {
    const long long nv = (long long)d * k_pad * c_pad, nh = (long long)d * N * k_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nv + nh; i += (long long)gridDim.x * blockDim.x) {
        if (i < nv) {
            int c = (int)(i % c_pad), k = (int)((i / c_pad) % k_pad), t = (int)(i / ((long long)c_pad * k_pad));
            float v = (c < C && k < K) ? v_w[((size_t)k * C + c) * d + t] : 0.f;
            wv[i] = __float2bfloat16_rn(v);
        } else {
            long long j = i - nv;
            int k = (int)(j % k_pad), n = (int)((j / k_pad) % N), t = (int)(j / ((long long)k_pad * N));
            float v = (k < K) ? h_w[((size_t)n * K + k) * d + t] : 0.f;
            wh[j] = __float2bfloat16_rn(v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
    });
    return fn;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
    return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int encode(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, CUtensorMapSwizzle swz) {
    EncodeTiledFn fn = get_encode();
    CNA_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box,
                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CNA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

static int pick_ck(int c_pad) { return c_pad % 64 == 0 ? 64 : (c_pad % 32 == 0 ? 32 : 16); }
static uint32_t up1k(uint32_t v) { return (v + 1023u) & ~1023u; }

// Fills the smem plan; returns total dynamic smem bytes (incl. 1 KB alignment slack) or 0 if it cannot fit.
static size_t plan_smem(LR2Params& p, int CK, int KP, size_t limit) {
    const uint32_t pool_bytes = p.pool ? 2 * 4096 : 0;
    const uint32_t xrow = CK * 2, trow = KP * 2;
    p.x_stage = up1k(kXRows * kTileCols * xrow);
    p.v_chunk = up1k(3 * KP * xrow);
    p.h_tile = up1k(3 * kNT * trow);
    p.t_buf = up1k((KP / 8) * kTLbo);
    p.y_buf = 16384;
    const uint32_t fixed = 2 * p.t_buf + 2 * p.y_buf + pool_bytes + up1k(p.N * 4) + 1024 /*barriers*/ + 1024 /*align slack*/;
    for (int resident = 1; resident >= 0; --resident) {
        for (int stages = kMaxStages; stages >= 2; --stages) {
            uint32_t vbytes = (resident ? p.nch : stages) * p.v_chunk;
            uint32_t hbytes = (resident ? p.nnt : 2) * p.h_tile;
            size_t total = (size_t)fixed + (size_t)stages * p.x_stage + vbytes + hbytes;
            if (total <= limit) {
                p.resident = resident;
                p.stages = stages;
                uint32_t o = 0;
                p.off_x = o; o += stages * p.x_stage;
                p.off_v = o; o += vbytes;
                p.off_h = o; o += hbytes;
                p.off_t = o; o += 2 * p.t_buf;
                p.off_y = o; o += 2 * p.y_buf;
                p.off_p = o; o += pool_bytes;
                p.off_bias = o; o += up1k(p.N * 4);
                p.off_bar = o;
                return total;
            }
        }
    }
    return 0;
}

static bool shape_ok(int H, int W, int c_pad, int k_pad, int N, int d) {
    return d == 3 && H >= 1 && W >= 1 && c_pad >= 16 && c_pad % 16 == 0 && (k_pad == 16 || k_pad == 32 || k_pad == 64) &&
           N >= 64 && N % 64 == 0 && N / 64 <= kMaxNTiles;
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_lowrank2_fused_supported(int H, int W, int c_pad, int k_pad, int N, int d) {
    if (!shape_ok(H, W, c_pad, k_pad, N, d)) return 0;
    LR2Params p{};
    int CK = pick_ck(c_pad);
    p.N = N; p.nch = c_pad / CK; p.nnt = N / kNT;
    return plan_smem(p, CK, k_pad, 227 * 1024) != 0 ? 1 : 0;
}

int cna_lowrank2_pack(const float* v_w, const float* h_w, void* wv_packed, void* wh_packed, int C, int K, int N, int d,
                      int c_pad, int k_pad, void* stream) {
    CNA_CHECK(v_w && h_w && wv_packed && wh_packed, "null pointer");
    CNA_CHECK(C > 0 && K > 0 && N > 0 && d > 0 && c_pad >= C && k_pad >= K, "bad dims");
    long long total = (long long)d * k_pad * c_pad + (long long)d * N * k_pad;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 8);
    lowrank2_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(v_w, h_w, (__nv_bfloat16*)wv_packed, (__nv_bfloat16*)wh_packed,
                                                             C, K, N, d, c_pad, k_pad);
    CNA_LAUNCH_OK();
    return 0;
}

static int lr2_launch(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y,
                      int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, int pool, void* stream,
                      long long* trace, int trace_tiles) {
    CNA_CHECK(x && wv_packed && wh_packed && y, "null pointer");
    CNA_CHECK(B > 0, "bad batch");
    CNA_CHECK(shape_ok(H, W, c_pad, k_pad, N, d), "configuration not supported by the fused kernel "
              "(needs d=3, c_pad%16==0, k_pad in {16,32,64}, N%64==0)");
    static int num_sms = 0, max_smem = 0;
    if (num_sms == 0) {
        int dev = 0;
        CNA_CUDA(cudaGetDevice(&dev));
        CNA_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        CNA_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    const int CK = pick_ck(c_pad);
    LR2Params p{};
    CNA_CHECK(!pool || (H % 2 == 0 && W % 2 == 0), "fused 2x2 max-pool needs even H and W");
    p.B = B; p.H = H; p.W = W; p.N = N; p.relu = relu; p.pool = pool ? 1 : 0; p.trace = trace; p.trace_tiles = trace_tiles;
    { const char* dbg = getenv("CNA_LR2_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    { const char* sk = getenv("CNA_LR2_TRACE_SKIP"); p.trace_skip = sk ? atoi(sk) : 0; }
    p.nch = c_pad / CK; p.nnt = N / kNT;
    p.tiles_w = ceil_div(W, kValidCols); p.tiles_h = ceil_div(H, kTileRows);
    long long ntiles = (long long)B * p.tiles_w * p.tiles_h;
    CNA_CHECK(ntiles < (1LL << 31), "too many tiles for one launch");
    p.ntiles = (int)ntiles;
    size_t smem = plan_smem(p, CK, k_pad, (size_t)max_smem);
    CNA_CHECK(smem != 0, "shared-memory plan does not fit");

    CUtensorMap tm_x, tm_wv, tm_wh, tm_y;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)c_pad * 2, (cuuint64_t)W * c_pad * 2, (cuuint64_t)H * W * c_pad * 2};
        cuuint32_t box[4] = {(cuuint32_t)CK, kTileCols, kXRows, 1};
        if (encode(&tm_x, x, 4, dims, str, box, swizzle_for(CK * 2))) return 1;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)c_pad, (cuuint64_t)3 * k_pad};
        cuuint64_t str[1] = {(cuuint64_t)c_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)CK, (cuuint32_t)(3 * k_pad)};
        if (encode(&tm_wv, wv_packed, 2, dims, str, box, swizzle_for(CK * 2))) return 1;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)3 * N};
        cuuint64_t str[1] = {(cuuint64_t)k_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)k_pad, kNT};
        if (encode(&tm_wh, wh_packed, 2, dims, str, box, swizzle_for(k_pad * 2))) return 1;
    }
    {
        const int Wy = pool ? W / 2 : W, Hy = pool ? H / 2 : H;
        // dimension order (C, H, W, B): the staged tile is then (col, row)-major, which is what the epilogue's TMEM lane
        // order (m = col*8 + row) writes without bank conflicts
        cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)Hy, (cuuint64_t)Wy, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)Wy * N * 2, (cuuint64_t)N * 2, (cuuint64_t)Hy * Wy * N * 2};
        cuuint32_t box[4] = {kNT, (cuuint32_t)(pool ? kTileRows / 2 : kTileRows), (cuuint32_t)(pool ? kValidCols / 2 : kValidCols), 1};
        if (encode(&tm_y, y, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    int grid = std::min(p.ntiles, num_sms);
    cudaStream_t st = as_stream(stream);
#define CNA_LR2_LAUNCH(CKV, KPV)                                                                                         \
    do {                                                                                                                 \
        static bool attr_set = false;                                                                                    \
        if (!attr_set) {                                                                                                 \
            CNA_CUDA(cudaFuncSetAttribute(lowrank2_fused_kernel<CKV, KPV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          max_smem));                                                                    \
            attr_set = true;                                                                                             \
        }                                                                                                                \
        lowrank2_fused_kernel<CKV, KPV><<<grid, kThreads, smem, st>>>(tm_x, tm_wv, tm_wh, tm_y, bias, p);                 \
    } while (0)
#define CNA_LR2_DISPATCH_KP(CKV)                                       \
    do {                                                               \
        if (k_pad == 16) CNA_LR2_LAUNCH(CKV, 16);                      \
        else if (k_pad == 32) CNA_LR2_LAUNCH(CKV, 32);                 \
        else CNA_LR2_LAUNCH(CKV, 64);                                  \
    } while (0)
    if (CK == 16) CNA_LR2_DISPATCH_KP(16);
    else if (CK == 32) CNA_LR2_DISPATCH_KP(32);
    else CNA_LR2_DISPATCH_KP(64);
#undef CNA_LR2_DISPATCH_KP
#undef CNA_LR2_LAUNCH
    CNA_LAUNCH_OK();
    return 0;
}

int cna_lowrank2_fused_forward(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y,
                               int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, int pool, void* stream) {
    return lr2_launch(x, wv_packed, wh_packed, bias, y, B, H, W, c_pad, k_pad, N, d, relu, pool, stream, nullptr, 0);
}

int cna_lowrank2_fused_forward_traced(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y,
                                      int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, void* stream,
                                      long long* trace, int trace_tiles) {
    return lr2_launch(x, wv_packed, wh_packed, bias, y, B, H, W, c_pad, k_pad, N, d, relu, 0, stream, trace, trace_tiles);
}

}  // extern "C"
