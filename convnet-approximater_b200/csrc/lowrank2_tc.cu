// Fused scheme-2 low-rank convolution for sm_100a:  y = h_conv(v_conv(x)) + bias  in ONE kernel.
//
//   T[b,h,w,k] = sum_dy sum_c x[b,h+dy-1,w,c] * Wv[dy][k][c]          GEMM1  (implicit GEMM, K-dim = 3*C)
//   y[b,h,w,n] = sum_dx sum_k T[b,h,w+dx-1,k] * Wh[dx][n][k] + b[n]   GEMM2  (implicit GEMM, K-dim = 3*K)
//
// Replaces LowRankExpConvV2.forward (reference src/approx/layers/low_rank_conv.py:83-84), where the
// K-channel intermediate T is a full HBM round trip between two cuDNN launches.  Here T never leaves
// the SM and never even touches shared memory: GEMM1 accumulates in TMEM, the T-epilogue warps round it
// to bf16 (exactly where the reference's bf16 modules materialise the v_conv output) and write it back
// to TMEM as the A operand of GEMM2 (tcgen05.mma with A in tensor memory).
//
// Tiling.  One tile = 8 output rows x 14 output columns of one image, handled as a 128-row UMMA M-tile of
// 8 rows x 16 columns, m = row*16 + col (one halo column each side, recomputed by the neighbouring tile).
//   * x tile (TMA, 4-D NHWC map, OOB zero fill = conv padding): 10 rows x 16 cols x CK channels in slot
//     order (row, col).  The three vertical taps of GEMM1 are the SAME smem tile read through UMMA
//     descriptors whose start address is shifted by 16 slots (one row): a multiple of the 8-row swizzle
//     atom, so no data movement and no re-load per tap.
//   * T: the horizontal tap dx of GEMM2 needs T shifted by dx-1 pixels = dx-1 TMEM lanes.  Lanes cannot be
//     shifted by the tensor core, so the T-epilogue writes THREE bf16 copies of T to TMEM — own lane, lane-1 and
//     lane+1, exchanged with one warp shuffle each.  A warp holds two whole tile rows, and the only lanes whose
//     neighbour lives in another warp are the halo columns, whose outputs are discarded anyway.
//   * GEMM2 then runs with A from TMEM (no shared-memory operand reads for A at all — on this part a
//     128x64x16 UMMA costs ~73 cycles with A in shared memory and ~41 with A in TMEM, benchmarks/ubench_tcgen05.cu)
//     and N = 128 output channels per instruction when the layer has them.
//   * y tile: TMEM -> registers (+bias, ReLU) -> swizzled smem staging [-> 2x2 max-pool] -> TMA store (clips at edges).
//     In the 64-output-channel instances the bias is not added by the epilogue: GEMM2 ends with one extra K = 16 UMMA whose A operand
//     is a ones column pair in 8 spare TMEM columns and whose B operand is a [N][16] bf16 tile with bf16(b), bf16(b - bf16(b)) in its
//     first two k (the bias to 16 mantissa bits, accumulated in fp32) - the epilogue there is issue-bound, the tensor pipe is not.
//
// Warp roles (640 threads, persistent CTAs, 1 per SM):
//   warp 0      x / Wv TMA producer          warp 15     Wh TMA producer (own ring, so neither ring blocks the other)
//   warp 1      GEMM1 issuer                 warp 10     GEMM2 issuer   (two issuing warps: their waits overlap)
//   warps 2-5 / 16-19   T-epilogue groups 0 / 1 (even / odd tiles)
//   warps 6-9 / 11-14   Y-epilogue groups 0 / 1 (two 64-channel items in flight)
// TMEM (512 columns): T accumulators 2 x 64 | A copies NABUF x 3 x KP/2 | Y accumulators 2 x NT2, with
// NABUF = 2 for NT2 = 64 and 1 for NT2 = 128 (TMEM is full there; GEMM1 of the next tiles fills the tensor pipe
// while the single A buffer is rewritten).
//
// Round-2 experiments that were measured and NOT kept (same box, same run: profiles/r02_fused_conv_experiments.txt):
//   * a second GEMM1 issuer on the otherwise idle warp (even / odd tiles on the two T accumulators; needs a progress hand-shake
//     between the issuers, because an mbarrier can only be waited on one phase ahead) + one TMA producer warp polling both
//     rings: layers 3-4 gain 4 %, layer 2 loses 13 %, the streamed layers 4-7 % — trunk 2.90 -> 3.00 ms;
//   * handing the three bf16 copies of T to GEMM2 one by one (own full/empty barrier pair per copy): layer 2 0.68 -> 0.77 ms — issuer
//     and epilogue then wait for each other three times per tile instead of once.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <algorithm>

namespace cna {
using namespace ptx;

constexpr int kThreads = 640;
constexpr int kTileRows = 8;           // output rows per tile
constexpr int kTileCols = 16;          // computed columns per tile (14 valid + 2 halo)
constexpr int kValidCols = kTileCols - 2;
constexpr int kXRows = kTileRows + 2;  // x rows per tile (d = 3)
constexpr int kNT = 64;                // output channels per Y-epilogue item / staging tile
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 6;
constexpr int kMaxHSlots = 12;         // Wh ring slots (one slot = one tap of one N-tile)
constexpr int kMaxNTiles = 16;         // N <= 1024
constexpr int kTraceEv = 20;
constexpr uint32_t kYBuf = kTileRows * kValidCols * 128;   // 14336-byte staging tile (112 pixels x 64 channels)

struct LR2Params {
    int B, H, W, N;
    int nch;               // number of channel chunks = c_pad / CK
    int nn2;               // number of GEMM2 N-tiles = N / NT2
    int tiles_w, tiles_h;  // tile grid per image
    int ntiles;
    int stages;            // x (+Wv) ring depth
    int hslots;            // Wh ring depth (streamed mode)
    int res_v, res_h;      // Wv / Wh stay in smem for the whole kernel (1) or are streamed per tile (0): Wv with its x chunk, Wh through its ring
    int relu;
    int pool;              // fuse a 2x2/stride-2 max-pool into the epilogue (y is then [B, H/2, W/2, N])
    // smem byte offsets from the 1024-aligned base
    uint32_t off_x, x_stage;      // x ring
    uint32_t off_v, v_chunk;      // Wv chunks (resident: nch of them; streamed: one per x stage)
    uint32_t off_h, h_slot;       // Wh tap slots (resident: nn2*3; streamed: hslots)
    uint32_t off_y;               // y staging: kYBuf x (pool ? 1 : nystg) per epilogue group
    uint32_t off_p;               // pooled staging: 4 KB x nystg per epilogue group, only when pool
    int nystg;                    // TMA-store source buffers per epilogue group = stores a group keeps in flight (1..4)
    int bias_mma;                 // 1: the bias enters through one extra K = 16 UMMA per N-tile (a ones column in TMEM against a [N][16] bias
    uint32_t off_bm;              //    tile in smem at off_bm) instead of 64 FADD + 16 LDS per thread and item in the Y-epilogue
    uint32_t off_bias, off_bar;
    // optional: the T-epilogue also writes bf16 T = v_conv(x) to global memory (element strides: image, tile-row axis, tile-col
    // axis); t_only = 1 skips GEMM2 and the y path altogether (a plain 3-tap C -> KP convolution along the tile-row axis)
    __nv_bfloat16* t_out;
    long long t_sb, t_sr, t_sc;
    int t_only;
    int debug;             // diagnostic bit mask (CNA_LR2_DEBUG): 1 skip y TMA store, 2 skip y staging stores, 16 load x only once, 32 skip A copies, 64 skip Y TMEM load, 128 skip GEMM1 MMAs
    long long* trace;      // diagnostic: per-tile clock64() stamps of CTA 0 (NULL = off), kTraceEv slots per tile
    int trace_tiles, trace_skip;
};

__host__ __device__ constexpr uint32_t swz_type(int row_bytes) { return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);     // .x = a in the low half
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// Same store without the compiler-level memory clobber: used where the only later readers are the async proxy / other
// threads behind a fence + barrier (both volatile asm, so program order is kept) — lets the compiler hoist the bias
// loads of the following chunks above the store instead of serialising load -> math -> store eight times.
__device__ __forceinline__ void st_shared_v4_relaxed(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d));
}
// (a + b) on two fp32 lanes at once (SASS FADD2), then one rounding to a bf16 pair — same arithmetic as two scalar adds
__device__ __forceinline__ uint32_t add2_pack_bf16(uint32_t a0, uint32_t a1, float b0, float b1) {
#ifdef CNA_LR2_FADD2
    unsigned long long ra = (unsigned long long)a0 | ((unsigned long long)a1 << 32), rd;
    unsigned long long rb = (unsigned long long)__float_as_uint(b0) | ((unsigned long long)__float_as_uint(b1) << 32);
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return pack_bf16(__uint_as_float((uint32_t)rd), __uint_as_float((uint32_t)(rd >> 32)));
#else
    return pack_bf16(__uint_as_float(a0) + b0, __uint_as_float(a1) + b1);
#endif
}
// ReLU on a packed bf16 pair: rounding is monotone, so max(round(v), 0) == round(max(v, 0))
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t v) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(0u));
    return d;
}
__device__ __forceinline__ float4 ld_shared_f4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// volatile (keeps its place among the other volatile asm statements, i.e. stays where the software pipeline put it), no memory clobber
__device__ __forceinline__ float4 ld_shared_f4_relaxed(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}

// barrier indices (each 8 bytes) inside the barrier block
enum : int {
    BAR_XFULL = 0,                         // [kMaxStages]
    BAR_XEMPTY = BAR_XFULL + kMaxStages,   // [kMaxStages]
    BAR_WFULL = BAR_XEMPTY + kMaxStages,   // resident weights landed
    BAR_HFULL = BAR_WFULL + 1,             // [kMaxHSlots] streamed Wh ring
    BAR_HEMPTY = BAR_HFULL + kMaxHSlots,   // [kMaxHSlots]
    BAR_TACC_FULL = BAR_HEMPTY + kMaxHSlots,   // [2] GEMM1 accumulator complete
    BAR_TACC_EMPTY = BAR_TACC_FULL + 2,    // [2] ... drained by the T-epilogue
    BAR_TA_FULL = BAR_TACC_EMPTY + 2,      // [2] bf16 copies of T written to TMEM
    BAR_TA_EMPTY = BAR_TA_FULL + 2,        // [2] GEMM2 of the tile has finished reading them
    BAR_YACC_FULL = BAR_TA_EMPTY + 2,      // [2]
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

// (image, tile row, tile column) of a CTA's current tile plus the decomposition of the grid stride: advancing costs a few additions
struct TileStep { int b, th, tw, db, dth, dtw; };
__device__ __forceinline__ TileStep tile_step_init(int first, int step, const LR2Params& p) {
    const int per_img = p.tiles_w * p.tiles_h;
    TileStep t;
    t.b = first / per_img; int rem = first - t.b * per_img;
    t.th = rem / p.tiles_w; t.tw = rem - t.th * p.tiles_w;
    t.db = step / per_img; rem = step - t.db * per_img;
    t.dth = rem / p.tiles_w; t.dtw = rem - t.dth * p.tiles_w;
    return t;
}
__device__ __forceinline__ void tile_step_advance(TileStep& t, const LR2Params& p) {
    t.tw += t.dtw; t.th += t.dth; t.b += t.db;
    if (t.tw >= p.tiles_w) { t.tw -= p.tiles_w; ++t.th; }
    if (t.th >= p.tiles_h) { t.th -= p.tiles_h; ++t.b; }
}

// TRACE = true compiles the per-tile clock stamps in (benchmarks/trace_lowrank2.py); the production instances carry none: the
// predicated-off stamp code was ~10 % of the instructions the epilogue warps issue per tile, and layer 1 / the K <= 32 forms are
// bound by exactly that issue rate (profiles/r02_fused_conv_l1_ncu.txt).
template <int CK, int KP, int NT2, bool TRACE = false>
__global__ void __launch_bounds__(kThreads, 1)
lowrank2_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wv,
                      const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_y,
                      const float* __restrict__ bias, const LR2Params p) {   // tm_y describes the pooled tensor when p.pool
    constexpr int XROW = CK * 2;            // bytes per x / Wv smem row (== swizzle span)
    constexpr int HROW = KP * 2;            // bytes per Wh smem row
    constexpr int ACOLS = KP / 2;           // TMEM columns of one bf16 copy of T
    constexpr int NABUF = NT2 == 64 ? 2 : 1;   // buffers of bf16 T copies (TMEM is full at NT2 = 128)
    // The bias UMMA (p.bias_mma: the bias enters as one extra K = 16 MMA, the Y-epilogue adds nothing) is THE bias path of the
    // 64-output-channel instances and absent from the 128-wide ones: measured per VGG layer on one box
    // (profiles/r02_fused_conv_experiments.txt) it pays on layers 1-2 (-8 %, -3 %; K = 16: layer 1 -17 %), is neutral on layer 4 and
    // costs 3.5 % on layer 3, and carrying both epilogue bodies in one instance costs ~2 % (spills in the 96-register CTA).
    constexpr bool kBM = NT2 == 64;
    constexpr uint32_t kTmemT = 0, kTmemA = 128, kTmemY = kTmemA + NABUF * 96;   // NT2=64: 0/128/320, NT2=128: 0/128/224
    static_assert(kTmemY + 2 * NT2 + 8 <= 512, "TMEM budget (incl. the 8 columns of the bias UMMA's ones operand)");
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
        for (int i = 0; i < kMaxHSlots; ++i) { mbar_init(bar(BAR_HFULL + i), 1); mbar_init(bar(BAR_HEMPTY + i), 1); }
        mbar_init(bar(BAR_WFULL), 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(BAR_TA_FULL + i), 128);  mbar_init(bar(BAR_TA_EMPTY + i), 1);
            mbar_init(bar(BAR_TACC_FULL + i), 1);  mbar_init(bar(BAR_TACC_EMPTY + i), 128);
            mbar_init(bar(BAR_YACC_FULL + i), 1);  mbar_init(bar(BAR_YACC_EMPTY + i), NT2 == 64 ? 128 : 256);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<kTmemCols>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    for (int i = threadIdx.x; i < p.N; i += kThreads) sbias[i] = bias ? bias[i] : 0.f;
    if (kBM && p.bias_mma) {
        // B operand of the bias UMMA: per N-tile a K-major [NT2 rows][16 k] bf16 tile in the 32-byte-swizzled layout TMA would write
        // (16-byte chunk ^= bit 2 of the row): k = 0 holds bf16(b), k = 1 bf16(b - bf16(b)), the other 14 k are zero.
        uint32_t* bm = reinterpret_cast<uint32_t*>(base_ptr + p.off_bm);
        for (int i = threadIdx.x; i < p.N * 8; i += kThreads) bm[i] = 0u;
        __syncthreads();
        for (int n = threadIdx.x; n < p.N; n += kThreads) {
            const float b = bias ? bias[n] : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(b), lo = __float2bfloat16_rn(b - __bfloat162float(hi));
            const int r = n % NT2;
            bm[(n / NT2) * (NT2 * 8) + r * 8 + (((r >> 2) & 1) ? 4 : 0)] =
                (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_T = tmem + kTmemT;   // + buf*64
    const uint32_t tmem_A = tmem + kTmemA;   // + abuf*96 + dx*ACOLS
    const uint32_t tmem_Y = tmem + kTmemY;   // + yb*NT2
    const uint32_t tmem_ones = tmem + kTmemY + 2 * NT2;      // A operand of the bias UMMA: 8 columns, k = 0, 1 are 1.0, the rest 0
    if (kBM && p.bias_mma) {
        if (warp >= 2 && warp < 6) {
            const uint32_t ones[8] = {0x3F803F80u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            tmem_st8(tmem_ones + ((uint32_t)((warp & 3) * 32) << 16), ones);
            tmem_st_wait();
        }
        tc_fence_before_sync();
        __syncthreads();
        tc_fence_after_sync();
    }

    // diagnostic: row `trace_tiles` of the trace holds {clock64, globaltimer} at CTA 0's start and end (effective SM clock)
    auto stamp_wall = [&](int slot) {
        if constexpr (!TRACE) return;
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
        if constexpr (!TRACE) return;
        if (p.trace != nullptr && blockIdx.x == 0 && it >= p.trace_skip && it < p.trace_skip + p.trace_tiles)
            p.trace[(it - p.trace_skip) * kTraceEv + ev] = clock64();
    };

    if (warp == 0) {
        // =========================== x / Wv TMA producer ===========================
        if (my_tiles > 0) {
            constexpr uint32_t x_bytes = kXRows * kTileCols * XROW;
            constexpr uint32_t v_bytes = 3 * KP * XROW;
            constexpr uint32_t h_bytes = NT2 * HROW;
            if ((p.res_v || p.res_h) && elect_one()) {
                const bool rh = p.res_h && !p.t_only;
                mbar_arrive_expect_tx(bar(BAR_WFULL), (p.res_v ? v_bytes * p.nch : 0u) + (rh ? h_bytes * 3 * p.nn2 : 0u));
                for (int ch = 0; ch < (p.res_v ? p.nch : 0); ++ch)
                    tma_load_2d(base + p.off_v + ch * p.v_chunk, &tm_wv, bar(BAR_WFULL), ch * CK, 0);
                for (int n2 = 0; n2 < (rh ? p.nn2 : 0); ++n2)
                    for (int dx = 0; dx < 3; ++dx)
                        tma_load_2d(base + p.off_h + (n2 * 3 + dx) * p.h_slot, &tm_wh, bar(BAR_WFULL), 0, dx * p.N + n2 * NT2);
            }
            __syncwarp();
            int xs = 0; uint32_t xph = 0;   // x ring stage / phase
            for (int it = 0; it < my_tiles; ++it) {
                TileCoord t = decode_tile(first + it * step, p);
                for (int ch = 0; ch < p.nch; ++ch) {
                    if ((p.debug & 16) && it * p.nch + ch >= p.stages) continue;
                    mbar_wait(bar(BAR_XEMPTY + xs), xph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(BAR_XFULL + xs), x_bytes + (p.res_v ? 0u : v_bytes));
                        tma_load_4d(base + p.off_x + xs * p.x_stage, &tm_x, bar(BAR_XFULL + xs), ch * CK, t.w0 - 1, t.h0 - 1, t.b);
                        if (!p.res_v)
                            tma_load_2d(base + p.off_v + xs * p.v_chunk, &tm_wv, bar(BAR_XFULL + xs), ch * CK, 0);
                    }
                    __syncwarp();
                    if (++xs == p.stages) { xs = 0; xph ^= 1u; }
                }
                if (lane == 0) stamp(it, 0);
            }
        }
    } else if (warp == 15 && !p.res_h && !p.t_only) {
        // =========================== Wh TMA producer (streamed weights only) ===========================
        if (my_tiles > 0) {
            constexpr uint32_t h_bytes = NT2 * HROW;
            int hs = 0; uint32_t hph = 0;
            for (int it = 0; it < my_tiles; ++it)
                for (int n2 = 0; n2 < p.nn2; ++n2)
                    for (int dx = 0; dx < 3; ++dx) {
                        mbar_wait(bar(BAR_HEMPTY + hs), hph ^ 1u);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(bar(BAR_HFULL + hs), h_bytes);
                            tma_load_2d(base + p.off_h + hs * p.h_slot, &tm_wh, bar(BAR_HFULL + hs), 0, dx * p.N + n2 * NT2);
                        }
                        __syncwarp();
                        if (++hs == p.hslots) { hs = 0; hph ^= 1u; }
                    }
        }
    } else if (warp == 1) {
        // =========================== GEMM1 issuer ===========================
        if (my_tiles > 0) {
            constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);
            constexpr uint64_t xdesc = desc_hi(16, 8 * XROW, swz_type(XROW));     // x tile and Wv chunk (swizzled, K-major)
            int xs = 0; uint32_t xph = 0;
            if (p.res_v) { mbar_wait(bar(BAR_WFULL), 0); tc_fence_after_sync(); }
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;          // barrier pair of this tile (its T-epilogue group)
                if (lane == 0) stamp(it, 1);
                {   // accumulator drained by tile it-2, first x chunk landed: probe both so their latencies overlap
                    const uint32_t b1 = bar(BAR_TACC_EMPTY + buf), p1 = (uint32_t)((it >> 1) & 1) ^ 1u;
                    const uint32_t ok1 = mbar_try_wait(b1, p1), ok2 = mbar_try_wait(bar(BAR_XFULL + xs), xph);
                    if (!ok1) mbar_wait(b1, p1);
                    (void)ok2;
                }
                tc_fence_after_sync();
                const uint32_t d = tmem_T + buf * 64;
                for (int ch = 0; ch < p.nch; ++ch) {
                    if (!((p.debug & 16) && it * p.nch + ch >= p.stages)) mbar_wait(bar(BAR_XFULL + xs), xph);
                    tc_fence_after_sync();
                    if (lane == 0 && ch == 0) stamp(it, 2);
                    const uint64_t a0 = xdesc + ((base + p.off_x + xs * p.x_stage) >> 4);
                    const uint64_t b0 = xdesc + ((base + p.off_v + (p.res_v ? ch : xs) * p.v_chunk) >> 4);
                    if (elect_one()) {
                        if (!(p.debug & 128))
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
    } else if (p.t_only && (warp == 10 || warp == 15)) {
        // 3-tap convolution only: no GEMM2
    } else if (warp == 10 || (warp == 15 && p.res_h && p.nn2 == 1)) {
        // =========================== GEMM2 issuer (A = bf16 T in TMEM, B = Wh tap slot in smem) ===========================
        // With resident weights and a single N-tile per tile (the 64-channel layers) the even and the odd tiles form two
        // independent pipelines (own T accumulator, A buffer, Y accumulator, epilogue groups), so they also get an issuing
        // warp each — warp 10 and the otherwise idle Wh producer warp 15: the per-tile barrier latencies of one issuer
        // (~750 cycles) then overlap the other's MMAs instead of idling the tensor pipe.
        const bool split = p.res_h && p.nn2 == 1;
        const int j0 = split ? (warp == 15 ? 1 : 0) : 0, jstep = split ? 2 : 1;
        if (my_tiles > 0) {
            constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT2 >> 3) << 17) | (8u << 24);
            constexpr uint64_t hdesc = desc_hi(16, 8 * HROW, swz_type(HROW));
            const uint64_t bmdesc = desc_hi(16, 8 * 32, swz_type(32)) + ((base + p.off_bm) >> 4);      // bias tile: 32-byte rows, 32B swizzle
            int hs = 0; uint32_t hph = 0;    // Wh ring
            if (p.res_h) { mbar_wait(bar(BAR_WFULL), 0); tc_fence_after_sync(); }
            for (int j = j0; j < my_tiles; j += jstep) {
                if (lane == 0) stamp(j, 4);
                const int ab = j & 1;
                int q = j * p.nn2;           // N-tile counter (Y accumulator index)
                {   // both barriers are usually complete already: probe them together so their latencies overlap
                    const uint32_t b1 = bar(BAR_TA_FULL + ab), p1 = (uint32_t)((j >> 1) & 1);
                    const uint32_t b2 = bar(BAR_YACC_EMPTY + (q & 1)), p2 = (uint32_t)((q >> 1) & 1) ^ 1u;
                    const uint32_t ok1 = mbar_try_wait(b1, p1), ok2 = mbar_try_wait(b2, p2);
                    if (!ok1) mbar_wait(b1, p1);
                    if (!ok2) mbar_wait(b2, p2);
                }
                tc_fence_after_sync();
                if (lane == 0) stamp(j, 5);
                for (int n2 = 0; n2 < p.nn2; ++n2, ++q) {
                    const int yb = q & 1;
                    if (n2 > 0) {
                        mbar_wait(bar(BAR_YACC_EMPTY + yb), (uint32_t)((q >> 1) & 1) ^ 1u);
                        tc_fence_after_sync();
                    }
                    if (lane == 0 && n2 == 0) stamp(j, 16);
                    const uint32_t d = tmem_Y + yb * NT2;
                    if (p.res_h) {
                        const uint64_t b0 = hdesc + ((base + p.off_h + n2 * 3 * p.h_slot) >> 4);
                        const uint32_t slot16 = p.h_slot >> 4;
                        if (elect_one()) {
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
#pragma unroll
                                for (int kk = 0; kk < KP / 16; ++kk)
                                    umma_bf16_ts(d, tmem_A + (NABUF == 2 ? ab * 96 : 0) + dx * ACOLS + kk * 8, b0 + dx * slot16 + ((kk * 32) >> 4),
                                                 idesc2, (uint32_t)((dx | kk) != 0));
                            if (kBM && p.bias_mma) umma_bf16_ts(d, tmem_ones, bmdesc + (uint64_t)((n2 * NT2 * 32) >> 4), idesc2, 1u);
                            umma_commit(bar(BAR_YACC_FULL + yb));
                            if (n2 == p.nn2 - 1) umma_commit(bar(BAR_TA_EMPTY + ab));
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            mbar_wait(bar(BAR_HFULL + hs), hph);
                            tc_fence_after_sync();
                            const uint64_t b0 = hdesc + ((base + p.off_h + hs * p.h_slot) >> 4);
                            if (elect_one()) {
#pragma unroll
                                for (int kk = 0; kk < KP / 16; ++kk)
                                    umma_bf16_ts(d, tmem_A + (NABUF == 2 ? ab * 96 : 0) + dx * ACOLS + kk * 8, b0 + ((kk * 32) >> 4), idesc2,
                                                 (uint32_t)((dx | kk) != 0));
                                umma_commit(bar(BAR_HEMPTY + hs));
                                if (dx == 2) {
                                    if (kBM && p.bias_mma) umma_bf16_ts(d, tmem_ones, bmdesc + (uint64_t)((n2 * NT2 * 32) >> 4), idesc2, 1u);
                                    umma_commit(bar(BAR_YACC_FULL + yb));
                                    if (n2 == p.nn2 - 1) umma_commit(bar(BAR_TA_EMPTY + ab));
                                }
                            }
                            __syncwarp();
                            if (++hs == p.hslots) { hs = 0; hph ^= 1u; }
                        }
                    }
                    if (lane == 0 && n2 == p.nn2 - 1) stamp(j, 17);
                }
            }
        }
    } else if (warp < 6 || warp >= 16) {
        // =========================== T-epilogue warps: group 0 = warps 2..5 (even tiles), group 1 = warps 16..19 (odd tiles) ===========================
        // GEMM1 accumulator (fp32, TMEM) -> bf16 -> three lane-shifted copies back into TMEM (the A operand of GEMM2).
        const int tg = warp >= 16 ? 1 : 0;
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const bool stamper = (warp == 2 || warp == 16) && lane == 0;
        const uint32_t tacc = tmem_T + tg * 64 + lane_addr;
        const uint32_t ta = tmem_A + (NABUF == 2 ? tg * 96 : 0) + lane_addr;
        for (int it = tg; it < my_tiles; it += 2) {
            const uint32_t ph = (uint32_t)((it >> 1) & 1);
            mbar_wait(bar(BAR_TACC_FULL + tg), ph);
            tc_fence_after_sync();
            if (stamper) stamp(it, 6);
            uint32_t v[KP];
#pragma unroll
            for (int c16 = 0; c16 < KP / 16; ++c16) tmem_ld16(tacc + c16 * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c16 * 16]));
            tmem_ld_wait();
            if (stamper) stamp(it, 10);
            tc_fence_before_sync();
            mbar_arrive(bar(BAR_TACC_EMPTY + tg));            // accumulator drained: the next GEMM1 into it may start
            uint32_t pk[ACOLS];                               // column c holds k = 2c (low half) and 2c+1 (high half)
#pragma unroll
            for (int c = 0; c < ACOLS; ++c) pk[c] = pack_bf16(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
            if (p.t_out != nullptr) {
                const TileCoord t = decode_tile(first + it * step, p);
                const int m = lg * 32 + lane, r = m >> 4, xc = m & 15;
                const int h = t.h0 + r, w = t.w0 + xc - 1;
                if (xc >= 1 && xc <= kValidCols && h < p.H && w < p.W) {
                    uint4* dst = reinterpret_cast<uint4*>(p.t_out + (long long)t.b * p.t_sb + (long long)h * p.t_sr + (long long)w * p.t_sc);
#pragma unroll
                    for (int c4 = 0; c4 < ACOLS / 4; ++c4) dst[c4] = make_uint4(pk[c4 * 4], pk[c4 * 4 + 1], pk[c4 * 4 + 2], pk[c4 * 4 + 3]);
                }
            }
            if (!p.t_only) {
            // the A buffer must have been read: by GEMM2 of tile it-2 (two buffers) or of tile it-1 (one, the other group's)
            if (NABUF == 2) mbar_wait(bar(BAR_TA_EMPTY + tg), ph ^ 1u);
            else if (it > 0) mbar_wait(bar(BAR_TA_EMPTY + (tg ^ 1)), (uint32_t)(((it - 1) >> 1) & 1));
            tc_fence_after_sync();
            if (stamper) stamp(it, 11);
            if (!(p.debug & 32))
#pragma unroll
            for (int c8 = 0; c8 < ACOLS / 8; ++c8) {
                uint32_t up[8], dn[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    up[i] = __shfl_up_sync(0xffffffffu, pk[c8 * 8 + i], 1);      // T of the pixel to the left  -> tap dx = 0
                    dn[i] = __shfl_down_sync(0xffffffffu, pk[c8 * 8 + i], 1);    // T of the pixel to the right -> tap dx = 2
                }
                tmem_st8(ta + 0 * ACOLS + c8 * 8, up);
                tmem_st8(ta + 1 * ACOLS + c8 * 8, *reinterpret_cast<uint32_t(*)[8]>(&pk[c8 * 8]));
                tmem_st8(ta + 2 * ACOLS + c8 * 8, dn);
            }
            if (stamper) stamp(it, 12);
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(bar(BAR_TA_FULL + tg));
            }
            if (stamper) stamp(it, 7);
        }
    } else if (warp == 15) {
        // resident weights and several N-tiles per tile: nothing to do for this warp
    } else {
        // =========================== Y-epilogue warps: group 0 = warps 6..9, group 1 = warps 11..14 ===========================
        // GEMM2 accumulator -> +bias, ReLU -> bf16 -> swizzled staging [-> 2x2 max-pool] -> TMA store, 64 channels per item.
        // NT2 == 64:  group g owns accumulator g and every N-tile q with (q & 1) == g.
        // NT2 == 128: both groups work on every N-tile, group g on its channels [64g, 64g+64).
        // Either way two 64-channel items are in flight: one group's ld -> math -> st -> fence -> store chain
        // (~1300 cycles) would otherwise bound the whole pipeline.
        const int g = warp >= 11 ? 1 : 0;
        const int lg = warp & 3;
        const int m = lg * 32 + lane;            // accumulator row: m = r*16 + xc
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const int et = (warp - (g ? 11 : 6)) * 32 + lane;   // 0..127 inside the group
        const uint32_t bar_a = 1 + 3 * g, bar_b = 2 + 3 * g, bar_c = 3 + 3 * g;
        const int r = m >> 4, xc = m & 15;
        const bool valid = xc >= 1 && xc <= kValidCols;
        const uint32_t srow = (uint32_t)(r * kValidCols + (xc - 1));      // staging row (TMA box order: row, col)
        const uint32_t row_off = srow * 128u, sw = srow & 7u;             // 128B-swizzle: chunk ^= row & 7
        // The store of an item drains for ~1500-2100 cycles once every SM is writing (profiles/r02_fused_conv_trace.txt), so a group
        // cycles through p.nystg source buffers and only waits for the store issued nystg items ago.
        const int nys = p.nystg;
        const uint32_t sb0 = base + p.off_y + g * (p.pool ? 1 : nys) * kYBuf;
        const uint32_t pb0 = base + p.off_p + g * nys * 4096;
        int bi = 0;
        const int nq = p.t_only ? 0 : my_tiles * p.nn2;
        const int q0 = NT2 == 64 ? g : 0, qstep = NT2 == 64 ? 2 : 1;
        int j = 0, n2 = q0;                      // q = j * nn2 + n2
        // tile coordinates of tile j, advanced by additions (a decode per item is two integer divisions = ~170 instructions in
        // every epilogue warp, and only the storing thread needs the result)
        TileStep tc = tile_step_init(first, step, p);
        while (n2 >= p.nn2 && j < my_tiles) { n2 -= p.nn2; ++j; tile_step_advance(tc, p); }
        for (int q = q0; q < nq; q += qstep) {
            const int yb = q & 1;
            const int chan0 = NT2 == 64 ? n2 * 64 : n2 * 128 + g * 64;
            mbar_wait(bar(BAR_YACC_FULL + yb), (uint32_t)((q >> 1) & 1));
            tc_fence_after_sync();
            if (et == 0 && n2 == 0) stamp(j, 8);
            const uint32_t tacc = tmem_Y + yb * NT2 + (NT2 == 64 ? 0 : g * 64) + lane_addr;
            uint32_t v[kNT];
            if (p.debug & 64) {
#pragma unroll
                for (int c = 0; c < kNT; ++c) v[c] = 0;
            } else {
#pragma unroll
                for (int c16 = 0; c16 < kNT / 16; ++c16) tmem_ld16(tacc + c16 * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c16 * 16]));
            }
            tmem_ld_wait();
            if (et == 0 && n2 == 0) stamp(j, 13);
            tc_fence_before_sync();
            mbar_arrive(bar(BAR_YACC_EMPTY + yb));       // accumulator drained: GEMM2 of N-tile q+2 may start
            // this group's staging buffer was handed to a TMA store one item ago: make sure it has been read
            // (waited for here, after the TMEM load, so that the store's smem read overlaps the load)
            if (et == 0) {
                if (nys == 1) tma_store_wait_read<0>();
                else if (nys == 2) tma_store_wait_read<1>();
                else if (nys == 3) tma_store_wait_read<2>();
                else tma_store_wait_read<3>();
            }
            const uint32_t sb = sb0 + (p.pool ? 0 : bi) * kYBuf, pb = pb0 + bi * 4096;
            if (++bi == nys) bi = 0;
            named_bar_sync(bar_a, 128);
            if (et == 0 && n2 == 0) stamp(j, 14);
            if constexpr (kBM) { if (valid && !(p.debug & 2)) {
                // the bias is already in the accumulator: round, ReLU, stage
#pragma unroll
                for (int ch = 0; ch < kNT / 8; ++ch) {
                    uint32_t o0 = pack_bf16(__uint_as_float(v[ch * 8 + 0]), __uint_as_float(v[ch * 8 + 1]));
                    uint32_t o1 = pack_bf16(__uint_as_float(v[ch * 8 + 2]), __uint_as_float(v[ch * 8 + 3]));
                    uint32_t o2 = pack_bf16(__uint_as_float(v[ch * 8 + 4]), __uint_as_float(v[ch * 8 + 5]));
                    uint32_t o3 = pack_bf16(__uint_as_float(v[ch * 8 + 6]), __uint_as_float(v[ch * 8 + 7]));
                    if (p.relu) { o0 = relu_bf16x2(o0); o1 = relu_bf16x2(o1); o2 = relu_bf16x2(o2); o3 = relu_bf16x2(o3); }
                    st_shared_v4_relaxed(sb + row_off + (((uint32_t)ch ^ sw) << 4), o0, o1, o2, o3);
                }
            } } else if (valid && !(p.debug & 2)) {
                // bias chunks are fetched two chunks ahead of their use: a shared-memory load issued right before its
                // first use costs the whole load latency eight times per item (short-scoreboard stalls in the profile)
                const uint32_t bz = smem_u32(sbias + chan0);
                float4 bq[4];
                bq[0] = ld_shared_f4_relaxed(bz); bq[1] = ld_shared_f4_relaxed(bz + 16);
                bq[2] = ld_shared_f4_relaxed(bz + 32); bq[3] = ld_shared_f4_relaxed(bz + 48);
#pragma unroll
                for (int ch = 0; ch < kNT / 8; ++ch) {
                    const float4 b0 = bq[(2 * ch) & 3], b1 = bq[(2 * ch + 1) & 3];
                    if (ch + 2 < kNT / 8) {
                        bq[(2 * ch) & 3] = ld_shared_f4_relaxed(bz + (ch + 2) * 32);
                        bq[(2 * ch + 1) & 3] = ld_shared_f4_relaxed(bz + (ch + 2) * 32 + 16);
                    }
                    uint32_t o0 = add2_pack_bf16(v[ch * 8 + 0], v[ch * 8 + 1], b0.x, b0.y), o1 = add2_pack_bf16(v[ch * 8 + 2], v[ch * 8 + 3], b0.z, b0.w);
                    uint32_t o2 = add2_pack_bf16(v[ch * 8 + 4], v[ch * 8 + 5], b1.x, b1.y), o3 = add2_pack_bf16(v[ch * 8 + 6], v[ch * 8 + 7], b1.z, b1.w);
                    if (p.relu) { o0 = relu_bf16x2(o0); o1 = relu_bf16x2(o1); o2 = relu_bf16x2(o2); o3 = relu_bf16x2(o3); }
                    st_shared_v4_relaxed(sb + row_off + (((uint32_t)ch ^ sw) << 4), o0, o1, o2, o3);
                }
            }
            if (!p.pool) {
                if (et == 0 && n2 == 0) stamp(j, 18);
                if (!(p.debug & 8)) fence_proxy_async_smem();
                if (et == 0 && n2 == 0) stamp(j, 19);
                named_bar_sync(bar_b, 128);
                if (et == 0 && n2 == 0) stamp(j, 15);
                if (et == 0 && !(p.debug & 1)) {
                    tma_store_4d(&tm_y, sb, chan0, tc.tw * kValidCols, tc.th * kTileRows, tc.b);
                    tma_store_commit();
                    if (n2 == p.nn2 - 1) stamp(j, 9);
                }
            } else {
                // 2x2 max-pool of the staged 8 x 14 tile -> 4 x 7 pooled tile (bf16 max commutes with the rounding)
                named_bar_sync(bar_b, 128);
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int id = et + rep * 128;            // 28 pooled pixels x 8 chunks of 16 bytes
                    if (id < 28 * 8) {
                        const int pp = id >> 3, ch = id & 7, pr = pp / 7, pc = pp - pr * 7;
                        uint4 m4 = make_uint4(0, 0, 0, 0);
#pragma unroll
                        for (int a = 0; a < 2; ++a)
#pragma unroll
                            for (int bq = 0; bq < 2; ++bq) {
                                const uint32_t row = (uint32_t)((2 * pr + a) * kValidCols + 2 * pc + bq);
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
                    tma_store_4d(&tm_y, pb, chan0, (tc.tw * kValidCols) >> 1, (tc.th * kTileRows) >> 1, tc.b);
                    tma_store_commit();
                    if (n2 == p.nn2 - 1) stamp(j, 9);
                }
            }
            n2 += qstep;
            while (n2 >= p.nn2) { n2 -= p.nn2; ++j; tile_step_advance(tc, p); }
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
static int pick_nt2(int N) { return N % 128 == 0 ? 128 : 64; }
static uint32_t up1k(uint32_t v) { return (v + 1023u) & ~1023u; }

// Fills the smem plan for `ystg` store buffers per epilogue group and the residency mode (rv, rh) = (Wv resident, Wh resident);
// returns total dynamic smem bytes (incl. 1 KB alignment slack) or 0 if it cannot fit.  A streamed Wv chunk travels with its x chunk
// (same ring stage), streamed Wh taps go through their own ring of `hslots` slots.
static size_t plan_smem_mode(LR2Params& p, int CK, int KP, int NT2, size_t limit, int ystg, int rv, int rh, int min_stages) {
    const uint32_t pool_bytes = p.pool ? 2 * ystg * 4096 : 0;
    const uint32_t ybytes = up1k((p.pool ? 2 : 2 * ystg) * kYBuf);
    const uint32_t xrow = CK * 2, hrow = KP * 2;
    p.nystg = ystg; p.res_v = rv; p.res_h = rh;      // (p.bias_mma is an input)
    p.x_stage = up1k(kXRows * kTileCols * xrow);
    p.v_chunk = up1k(3 * KP * xrow);
    p.h_slot = up1k(NT2 * hrow);
    const uint32_t bm_bytes = p.bias_mma ? up1k(p.N * 32) : 0;       // [N][16] bf16 bias tile of the bias UMMA
    const size_t fixed = (size_t)ybytes + pool_bytes + up1k(p.N * 4) + bm_bytes + 1024 /*barriers*/ + 1024 /*align slack*/ +
                         (rv ? (size_t)p.nch * p.v_chunk : 0) + (rh ? (size_t)3 * p.nn2 * p.h_slot : 0);
    const size_t per_stage = (size_t)p.x_stage + (rv ? 0 : p.v_chunk);
    for (int stages = kMaxStages; stages >= min_stages; --stages) {
        const size_t used = fixed + (size_t)stages * per_stage;
        if (used + (rh ? 0 : 3 * (size_t)p.h_slot) > limit) continue;
        const int hslots = rh ? 0 : (int)std::min<size_t>(kMaxHSlots, (limit - used) / p.h_slot);
        p.stages = stages; p.hslots = hslots;
        uint32_t o = 0;
        p.off_x = o; o += stages * p.x_stage;
        p.off_v = o; o += (rv ? p.nch : stages) * p.v_chunk;
        p.off_h = o; o += (rh ? 3 * p.nn2 : hslots) * p.h_slot;
        p.off_y = o; o += ybytes;
        p.off_p = o; o += pool_bytes;
        p.off_bias = o; o += up1k(p.N * 4);
        p.off_bm = o; o += bm_bytes;
        p.off_bar = o;
        return (size_t)o + 2048;
    }
    return 0;
}

// Residency: everything resident when it fits next to an x ring of >= 3 stages (the 64- and 128-channel layers).  Otherwise one of
// Wv / Wh may stay resident while the other streams - measured per VGG layer on one box (profiles/r02_fused_conv_residency.txt): it pays
// where the x ring still holds a whole tile (<= 2 channel chunks: 128 -> 256 @56, -5 %) or where Wh is the bulk of the traffic
// (>= 4 N-tiles: 256 -> 512 @28, -8 %), and costs 15 % on 256 -> 256 @56, whose rings shrink to 2 stages / 3 slots.
// CNA_LR2_MODE=<rv><rh> (e.g. 01) forces a mode wherever it fits (diagnostic sweeps).
static size_t plan_smem_ystg(LR2Params& p, int CK, int KP, int NT2, size_t limit, int ystg) {
    static const int forced = [] { const char* e = getenv("CNA_LR2_MODE"); return (e && e[0] && e[1]) ? ((e[0] - '0') * 2 + (e[1] - '0')) : -1; }();
    if (forced >= 0) {
        size_t s = plan_smem_mode(p, CK, KP, NT2, limit, ystg, forced >> 1, forced & 1, 2);
        if (s) return s;
    }
    if (size_t s = plan_smem_mode(p, CK, KP, NT2, limit, ystg, 1, 1, 3)) return s;
    if (p.nch <= 2 || p.nn2 >= 4) {
        const size_t v_all = (size_t)p.nch * up1k(3 * KP * CK * 2), h_all = (size_t)3 * p.nn2 * up1k(NT2 * KP * 2);
        const int first_v = v_all >= h_all ? 1 : 0;          // try to keep the larger operand resident first
        if (size_t s = plan_smem_mode(p, CK, KP, NT2, limit, ystg, first_v, 1 - first_v, 2)) return s;
        if (size_t s = plan_smem_mode(p, CK, KP, NT2, limit, ystg, 1 - first_v, first_v, 2)) return s;
    }
    return plan_smem_mode(p, CK, KP, NT2, limit, ystg, 0, 0, 2);
}

// Store-buffer depth: the deepest (<= kAutoYStg) that costs the plan neither its resident weights nor an x stage nor Wh slots below
// a tile's worth; CNA_LR2_YSTG forces a depth (diagnostic sweeps), falling back towards 1 if it does not fit at all.
constexpr int kAutoYStg = 3;
static size_t plan_smem_nobm(LR2Params& p, int CK, int KP, int NT2, size_t limit) {
    static const int forced = [] { const char* e = getenv("CNA_LR2_YSTG"); return e ? std::max(1, std::min(4, atoi(e))) : 0; }();
    if (forced) {
        for (int y = forced; y >= 1; --y) { size_t s = plan_smem_ystg(p, CK, KP, NT2, limit, y); if (s) return s; }
        return 0;
    }
    LR2Params base = p;
    size_t s1 = plan_smem_ystg(base, CK, KP, NT2, limit, 1);
    if (!s1) return 0;
    for (int y = kAutoYStg; y >= 2; --y) {
        LR2Params q = p;
        size_t s = plan_smem_ystg(q, CK, KP, NT2, limit, y);
        if (s && q.res_v == base.res_v && q.res_h == base.res_h && q.stages >= base.stages &&
            (q.res_h || q.hslots >= std::min(base.hslots, 3 * p.nn2))) { p = q; return s; }
    }
    p = base;
    return s1;
}

// The bias UMMA (p.bias_mma) needs a [N][16] bf16 tile in shared memory (2 KB per 64 output channels): always on for the
// 64-output-channel instances (the kernel has no other bias path there), never for the 128-wide ones or the 3-tap-only mode.
static size_t plan_smem(LR2Params& p, int CK, int KP, int NT2, size_t limit) {
    p.bias_mma = (NT2 == 64 && !p.t_only) ? 1 : 0;
    return plan_smem_nobm(p, CK, KP, NT2, limit);
}

static bool shape_ok(int H, int W, int c_pad, int k_pad, int N, int d) {
    return d == 3 && H >= 1 && W >= 1 && c_pad >= 16 && c_pad % 16 == 0 && (k_pad == 16 || k_pad == 32 || k_pad == 64) &&
           N >= 64 && N % 64 == 0 && N / 64 <= kMaxNTiles;
}

}  // namespace cna

using namespace cna;

template <int CKV, int KPV, int NT2V>
static int lr2_launch_one(int grid, size_t smem, int max_smem, cudaStream_t st, const CUtensorMap& tm_x, const CUtensorMap& tm_wv,
                          const CUtensorMap& tm_wh, const CUtensorMap& tm_y, const float* bias, const LR2Params& p) {
    if (p.trace != nullptr) {
        // traced instances exist for the configurations the trace tool is used on (each one is a full extra compile)
        constexpr bool kTraced = (KPV == 64 && (CKV == 16 || CKV == 64)) || (KPV == 16 && CKV == 64 && NT2V == 64);
        if constexpr (kTraced) {
            static std::atomic<unsigned long long> attr_done_t{0};
            CNA_CUDA(ensure_dyn_smem(lowrank2_fused_kernel<CKV, KPV, NT2V, true>, max_smem, attr_done_t));
            lowrank2_fused_kernel<CKV, KPV, NT2V, true><<<grid, kThreads, smem, st>>>(tm_x, tm_wv, tm_wh, tm_y, bias, p);
            return 0;
        } else {
            CNA_CHECK(false, "no traced instance of the fused kernel for this (channel chunk, rank, N-tile) combination");
        }
    }
    static std::atomic<unsigned long long> attr_done{0};
    CNA_CUDA(ensure_dyn_smem(lowrank2_fused_kernel<CKV, KPV, NT2V>, max_smem, attr_done));
    lowrank2_fused_kernel<CKV, KPV, NT2V><<<grid, kThreads, smem, st>>>(tm_x, tm_wv, tm_wh, tm_y, bias, p);
    return 0;
}

template <int CKV, int KPV>
static int lr2_launch_nt(int NT2, int grid, size_t smem, int max_smem, cudaStream_t st, const CUtensorMap& tm_x, const CUtensorMap& tm_wv,
                         const CUtensorMap& tm_wh, const CUtensorMap& tm_y, const float* bias, const LR2Params& p) {
    return NT2 == 128 ? lr2_launch_one<CKV, KPV, 128>(grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p)
                      : lr2_launch_one<CKV, KPV, 64>(grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
}

template <int CKV>
static int lr2_launch_kp(int KP, int NT2, int grid, size_t smem, int max_smem, cudaStream_t st, const CUtensorMap& tm_x,
                         const CUtensorMap& tm_wv, const CUtensorMap& tm_wh, const CUtensorMap& tm_y, const float* bias, const LR2Params& p) {
    if (KP == 16) return lr2_launch_nt<CKV, 16>(NT2, grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
    if (KP == 32) return lr2_launch_nt<CKV, 32>(NT2, grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
    return lr2_launch_nt<CKV, 64>(NT2, grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
}

// axis: 0 = vertical taps run along the image H axis (the normal forward), 1 = along W (the image is presented to the kernel
// transposed through the tensor maps' strides; only available in t_only mode).  H, W are the IMAGE dims.
static int lr2_launch(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y, void* t_out, int t_only,
                      int axis, int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, int pool, void* stream,
                      long long* trace, int trace_tiles) {
    CNA_CHECK(x && wv_packed && (t_only || (wh_packed && y)) && (!t_only || t_out), "null pointer");
    CNA_CHECK(B > 0, "bad batch");
    if (t_only) { N = 64; pool = 0; }
    CNA_CHECK(axis == 0 || t_only, "axis 1 is only available for the 3-tap convolution");
    CNA_CHECK(shape_ok(H, W, c_pad, k_pad, N, d), "configuration not supported by the fused kernel "
              "(needs d=3, c_pad%16==0, k_pad in {16,32,64}, N%64==0)");
    DevInfo di;
    if (dev_info(&di)) return 1;
    const int num_sms = di.num_sms, max_smem = di.max_smem;
    const int CK = pick_ck(c_pad), NT2 = pick_nt2(N);
    const int KH = axis == 0 ? H : W, KW = axis == 0 ? W : H;      // dims along the kernel's tile-row / tile-col axes
    LR2Params p{};
    CNA_CHECK(!pool || (H % 2 == 0 && W % 2 == 0), "fused 2x2 max-pool needs even H and W");
    p.B = B; p.H = KH; p.W = KW; p.N = N; p.relu = relu; p.pool = pool ? 1 : 0; p.trace = trace; p.trace_tiles = trace_tiles;
    p.t_out = (__nv_bfloat16*)t_out; p.t_only = t_only ? 1 : 0;
    p.t_sb = (long long)H * W * k_pad;
    p.t_sr = axis == 0 ? (long long)W * k_pad : (long long)k_pad;
    p.t_sc = axis == 0 ? (long long)k_pad : (long long)W * k_pad;
    { const char* dbg = getenv("CNA_LR2_DEBUG"); p.debug = dbg ? atoi(dbg) : 0; }
    { const char* sk = getenv("CNA_LR2_TRACE_SKIP"); p.trace_skip = sk ? atoi(sk) : 0; }
    p.nch = c_pad / CK; p.nn2 = N / NT2;
    p.tiles_w = ceil_div(KW, kValidCols); p.tiles_h = ceil_div(KH, kTileRows);
    long long ntiles = (long long)B * p.tiles_w * p.tiles_h;
    CNA_CHECK(ntiles < (1LL << 31), "too many tiles for one launch");
    p.ntiles = (int)ntiles;
    size_t smem = plan_smem(p, CK, k_pad, NT2, (size_t)max_smem);
    CNA_CHECK(smem != 0, "shared-memory plan does not fit");
    if (getenv("CNA_LR2_VERBOSE") != nullptr)
        fprintf(stderr, "lowrank2 plan: C=%d K=%d N=%d H=%d W=%d pool=%d | CK=%d NT2=%d Wv %s Wh %s x-stages=%d hslots=%d ystg=%d bias_mma=%d smem=%zu\n", c_pad, k_pad,
                N, H, W, pool, CK, NT2, p.res_v ? "resident" : "streamed", p.res_h ? "resident" : "streamed", p.stages, p.hslots, p.nystg, p.bias_mma, smem);

    CUtensorMap tm_x, tm_wv, tm_wh, tm_y;
    {
        // dims (channels, tile-col axis, tile-row axis, image)
        const cuuint64_t s_h = (cuuint64_t)W * c_pad * 2, s_w = (cuuint64_t)c_pad * 2;
        cuuint64_t dims[4] = {(cuuint64_t)c_pad, (cuuint64_t)KW, (cuuint64_t)KH, (cuuint64_t)B};
        cuuint64_t str[3] = {axis == 0 ? s_w : s_h, axis == 0 ? s_h : s_w, (cuuint64_t)H * W * c_pad * 2};
        cuuint32_t box[4] = {(cuuint32_t)CK, kTileCols, kXRows, 1};
        if (encode(&tm_x, x, 4, dims, str, box, swizzle_for(CK * 2))) return 1;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)c_pad, (cuuint64_t)3 * k_pad};
        cuuint64_t str[1] = {(cuuint64_t)c_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)CK, (cuuint32_t)(3 * k_pad)};
        if (encode(&tm_wv, wv_packed, 2, dims, str, box, swizzle_for(CK * 2))) return 1;
    }
    if (t_only) {
        tm_wh = tm_wv; tm_y = tm_x;        // never dereferenced in this mode
    } else {
        {
            cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)3 * N};
            cuuint64_t str[1] = {(cuuint64_t)k_pad * 2};
            cuuint32_t box[2] = {(cuuint32_t)k_pad, (cuuint32_t)NT2};
            if (encode(&tm_wh, wh_packed, 2, dims, str, box, swizzle_for(k_pad * 2))) return 1;
        }
        const int Wy = pool ? W / 2 : W, Hy = pool ? H / 2 : H;
        cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)Wy, (cuuint64_t)Hy, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)N * 2, (cuuint64_t)Wy * N * 2, (cuuint64_t)Hy * Wy * N * 2};
        cuuint32_t box[4] = {kNT, (cuuint32_t)(pool ? kValidCols / 2 : kValidCols), (cuuint32_t)(pool ? kTileRows / 2 : kTileRows), 1};
        if (encode(&tm_y, y, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    const int grid = std::min(p.ntiles, num_sms);
    cudaStream_t st = as_stream(stream);
    int rc;
    if (CK == 16) rc = lr2_launch_kp<16>(k_pad, NT2, grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
    else if (CK == 32) rc = lr2_launch_kp<32>(k_pad, NT2, grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
    else rc = lr2_launch_kp<64>(k_pad, NT2, grid, smem, max_smem, st, tm_x, tm_wv, tm_wh, tm_y, bias, p);
    if (rc) return rc;
    CNA_LAUNCH_OK();
    return 0;
}

extern "C" {

int cna_lowrank2_fused_supported(int H, int W, int c_pad, int k_pad, int N, int d) {
    if (!shape_ok(H, W, c_pad, k_pad, N, d)) return 0;
    LR2Params p{};
    const int CK = pick_ck(c_pad), NT2 = pick_nt2(N);
    p.N = N; p.nch = c_pad / CK; p.nn2 = N / NT2; p.pool = 1;
    return plan_smem(p, CK, k_pad, NT2, 227 * 1024) != 0 ? 1 : 0;
}

int cna_lowrank2_fused_plan(int c_pad, int k_pad, int N, int pool, int smem_limit, int* out) {
    CNA_CHECK(out != nullptr, "null pointer");
    CNA_CHECK(shape_ok(8, 16, c_pad, k_pad, N, 3), "configuration not supported by the fused kernel");
    LR2Params p{};
    const int CK = pick_ck(c_pad), NT2 = pick_nt2(N);
    p.N = N; p.nch = c_pad / CK; p.nn2 = N / NT2; p.pool = pool ? 1 : 0;
    const size_t smem = plan_smem(p, CK, k_pad, NT2, (size_t)smem_limit);
    CNA_CHECK(smem != 0, "shared-memory plan does not fit");
    out[0] = (int)smem; out[1] = p.res_v; out[2] = p.res_h; out[3] = p.stages; out[4] = p.hslots; out[5] = p.nystg; out[6] = CK; out[7] = NT2;
    return 0;
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

int cna_lowrank2_fused_forward(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y,
                               int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, int pool, void* stream) {
    return lr2_launch(x, wv_packed, wh_packed, bias, y, nullptr, 0, 0, B, H, W, c_pad, k_pad, N, d, relu, pool, stream, nullptr, 0);
}

int cna_lowrank2_fused_forward_t(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y, void* t_out,
                                 int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, void* stream) {
    CNA_CHECK(t_out != nullptr, "null pointer");
    return lr2_launch(x, wv_packed, wh_packed, bias, y, t_out, 0, 0, B, H, W, c_pad, k_pad, N, d, relu, 0, stream, nullptr, 0);
}

int cna_tapconv3_tc_forward(const void* x, const void* w_packed, void* out, int B, int H, int W, int c_pad, int k_pad, int axis,
                            void* stream) {
    CNA_CHECK(axis == 0 || axis == 1, "axis must be 0 (taps along H) or 1 (along W)");
    return lr2_launch(x, w_packed, nullptr, nullptr, nullptr, out, 1, axis, B, H, W, c_pad, k_pad, 64, 3, 0, 0, stream, nullptr, 0);
}

int cna_lowrank2_fused_forward_traced(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y,
                                      int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, void* stream,
                                      long long* trace, int trace_tiles) {
    return lr2_launch(x, wv_packed, wh_packed, bias, y, nullptr, 0, 0, B, H, W, c_pad, k_pad, N, d, relu, 0, stream, trace, trace_tiles);
}

}  // extern "C"
