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
#include <map>
#include <mutex>
#include <tuple>

namespace cna {
using namespace ptx;

constexpr int kThreads = 192;
constexpr int kTileRows = 8;           // output rows per tile
constexpr int kTileCols = 16;          // computed columns per tile (14 valid + 2 halo)
constexpr int kValidCols = kTileCols - 2;
constexpr int kXRows = kTileRows + 2;  // x rows per tile (d = 3)
constexpr int kTCols = kTileCols + 2;  // T columns kept in smem
constexpr int kNT = 64;                // GEMM2 N-tile (output channels per accumulator)
constexpr int kTmemCols = 256;
constexpr int kMaxStages = 4;
constexpr int kMaxNTiles = 16;         // N <= 1024

struct LR2Params {
    int B, H, W, N;
    int CK, KP;            // channel chunk (16/32/64), padded rank (16/32/64)
    int nch;               // number of channel chunks = c_pad / CK
    int nnt;               // N / 64
    int tiles_w, tiles_h;  // tile grid per image
    long long ntiles;
    int stages;            // x ring depth
    int resident;          // 1: all weights stay in smem for the whole kernel; 0: streamed per tile
    int relu;
    // smem byte offsets from the 1024-aligned base
    uint32_t off_x, x_stage;      // x ring
    uint32_t off_v, v_chunk;      // Wv chunks (resident: nch of them; streamed: one per x stage)
    uint32_t off_h, h_tile;       // Wh N-tiles (resident: nnt; streamed: 2)
    uint32_t off_t, t_buf;        // T buffers (2)
    uint32_t off_y, y_buf;        // y staging (2)
    uint32_t off_bias, off_bar;
};

__device__ __forceinline__ uint32_t swz_type(int row_bytes) { return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u); }
// canonical Swizzle<B,4,3> on a byte offset relative to a 1024-aligned base
__device__ __forceinline__ uint32_t swz_off(uint32_t off, int row_bytes) {
    uint32_t mask = row_bytes == 128 ? 7u : (row_bytes == 64 ? 3u : 1u);
    return off ^ (((off >> 7) & mask) << 4);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

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

__global__ void __launch_bounds__(kThreads, 1)
lowrank2_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wv,
                      const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_y,
                      const float* __restrict__ bias, const LR2Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + p.off_bar;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    float* sbias = reinterpret_cast<float*>(base_ptr + p.off_bias);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + p.off_bar + 8 * BAR_COUNT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int xrow_b = p.CK * 2;   // bytes per x / Wv smem row
    const int trow_b = p.KP * 2;   // bytes per T / Wh smem row

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
    // T columns 0 and 17 are never produced (they only feed the two discarded halo outputs): keep them finite.
    for (int i = threadIdx.x; i < 2 * 2 * kTileRows * trow_b / 16; i += kThreads) {
        int per_buf = 2 * kTileRows * trow_b / 16;
        int bsel = i / per_buf, j = i % per_buf;
        int slot = j / (trow_b / 16);                              // 0..15 -> slots 0..7 and 136..143
        int ch = j % (trow_b / 16);
        int s = slot < kTileRows ? slot : (kTCols - 1) * kTileRows + (slot - kTileRows);
        st_shared_v4(base + p.off_t + bsel * p.t_buf + s * trow_b + ch * 16, 0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_T = tmem;            // + buf*64
    const uint32_t tmem_Y = tmem + 128;      // + buf*64

    const long long first = blockIdx.x, step = gridDim.x;
    const long long my_tiles = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;
    const int tiles_per_img = p.tiles_w * p.tiles_h;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t x_bytes = (uint32_t)(kXRows * kTileCols * xrow_b);
            const uint32_t v_bytes = (uint32_t)(3 * p.KP * xrow_b);
            const uint32_t h_bytes = (uint32_t)(3 * kNT * trow_b);
            if (p.resident) {
                mbar_arrive_expect_tx(bar(BAR_WFULL), v_bytes * p.nch + h_bytes * p.nnt);
                for (int ch = 0; ch < p.nch; ++ch)
                    tma_load_2d(base + p.off_v + ch * p.v_chunk, &tm_wv, bar(BAR_WFULL), ch * p.CK, 0);
                for (int nt = 0; nt < p.nnt; ++nt)
                    for (int dx = 0; dx < 3; ++dx)
                        tma_load_2d(base + p.off_h + nt * p.h_tile + dx * (kNT * trow_b), &tm_wh, bar(BAR_WFULL), 0,
                                    dx * p.N + nt * kNT);
            }
            long long kx = 0, qh = 0;  // global x-chunk / Wh-tile counters
            auto load_x = [&](long long it) {
                long long tile = first + it * step;
                int b = (int)(tile / tiles_per_img);
                int rem = (int)(tile % tiles_per_img);
                int h0 = (rem / p.tiles_w) * kTileRows, w0 = (rem % p.tiles_w) * kValidCols;
                for (int ch = 0; ch < p.nch; ++ch, ++kx) {
                    int s = (int)(kx % p.stages);
                    mbar_wait(bar(BAR_XEMPTY + s), (uint32_t)((kx / p.stages) & 1) ^ 1u);
                    mbar_arrive_expect_tx(bar(BAR_XFULL + s), x_bytes + (p.resident ? 0u : v_bytes));
                    tma_load_4d(base + p.off_x + s * p.x_stage, &tm_x, bar(BAR_XFULL + s), ch * p.CK, w0 - 1, h0 - 1, b);
                    if (!p.resident)
                        tma_load_2d(base + p.off_v + s * p.v_chunk, &tm_wv, bar(BAR_XFULL + s), ch * p.CK, 0);
                }
            };
            auto load_h = [&]() {
                if (p.resident) return;
                for (int nt = 0; nt < p.nnt; ++nt, ++qh) {
                    int s = (int)(qh & 1);
                    mbar_wait(bar(BAR_HEMPTY + s), (uint32_t)((qh >> 1) & 1) ^ 1u);
                    mbar_arrive_expect_tx(bar(BAR_HFULL + s), h_bytes);
                    for (int dx = 0; dx < 3; ++dx)
                        tma_load_2d(base + p.off_h + s * p.h_tile + dx * (kNT * trow_b), &tm_wh, bar(BAR_HFULL + s), 0,
                                    dx * p.N + nt * kNT);
                }
            };
            // same order as the MMA warp consumes: x(0) | x(1) h(0) | x(2) h(1) | ...
            load_x(0);
            for (long long it = 0; it < my_tiles; ++it) {
                if (it + 1 < my_tiles) load_x(it + 1);
                load_h();
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0 && my_tiles > 0) {
            const uint32_t idesc1 = umma_idesc_bf16(128, (uint32_t)p.KP);
            const uint32_t idesc2 = umma_idesc_bf16(128, (uint32_t)kNT);
            const uint32_t xs = swz_type(xrow_b), ts = swz_type(trow_b);
            const int kk1 = p.CK / 16, kk2 = p.KP / 16;
            long long kx = 0, q = 0;
            if (p.resident) { mbar_wait(bar(BAR_WFULL), 0); tc_fence_after_sync(); }
            auto gemm1 = [&](long long it) {
                int buf = (int)(it & 1);
                mbar_wait(bar(BAR_TACC_EMPTY + buf), (uint32_t)((it >> 1) & 1) ^ 1u);
                tc_fence_after_sync();
                for (int ch = 0; ch < p.nch; ++ch, ++kx) {
                    int s = (int)(kx % p.stages);
                    mbar_wait(bar(BAR_XFULL + s), (uint32_t)((kx / p.stages) & 1));
                    tc_fence_after_sync();
                    uint32_t a0 = base + p.off_x + s * p.x_stage;
                    uint32_t b0 = base + p.off_v + (p.resident ? ch : s) * p.v_chunk;
                    for (int dy = 0; dy < 3; ++dy)
                        for (int kk = 0; kk < kk1; ++kk) {
                            uint64_t ad = umma_desc(a0 + dy * kTileCols * xrow_b + kk * 32, 8 * xrow_b, xs);
                            uint64_t bd = umma_desc(b0 + dy * p.KP * xrow_b + kk * 32, 8 * xrow_b, xs);
                            umma_bf16(tmem_T + buf * 64, ad, bd, idesc1, (ch | dy | kk) != 0);
                        }
                    umma_commit(bar(BAR_XEMPTY + s));
                }
                umma_commit(bar(BAR_TACC_FULL + buf));
            };
            auto gemm2 = [&](long long j) {
                int buf = (int)(j & 1);
                mbar_wait(bar(BAR_TSM_FULL + buf), (uint32_t)((j >> 1) & 1));
                tc_fence_after_sync();
                uint32_t a0 = base + p.off_t + buf * p.t_buf;
                for (int nt = 0; nt < p.nnt; ++nt, ++q) {
                    int yb = (int)(q & 1);
                    mbar_wait(bar(BAR_YACC_EMPTY + yb), (uint32_t)((q >> 1) & 1) ^ 1u);
                    uint32_t b0;
                    if (p.resident) {
                        b0 = base + p.off_h + nt * p.h_tile;
                    } else {
                        mbar_wait(bar(BAR_HFULL + yb), (uint32_t)((q >> 1) & 1));
                        b0 = base + p.off_h + yb * p.h_tile;
                    }
                    tc_fence_after_sync();
                    for (int dx = 0; dx < 3; ++dx)
                        for (int kk = 0; kk < kk2; ++kk) {
                            uint64_t ad = umma_desc(a0 + dx * kTileRows * trow_b + kk * 32, 8 * trow_b, ts);
                            uint64_t bd = umma_desc(b0 + dx * (kNT * trow_b) + kk * 32, 8 * trow_b, ts);
                            umma_bf16(tmem_Y + yb * 64, ad, bd, idesc2, (dx | kk) != 0);
                        }
                    if (!p.resident) umma_commit(bar(BAR_HEMPTY + yb));
                    umma_commit(bar(BAR_YACC_FULL + yb));
                }
                umma_commit(bar(BAR_TSM_EMPTY + buf));
            };
            gemm1(0);
            for (long long it = 0; it < my_tiles; ++it) {
                if (it + 1 < my_tiles) gemm1(it + 1);
                gemm2(it);
            }
        }
    } else {
        // =========================== epilogue warps ===========================
        const int lg = warp & 3;                 // TMEM lane group this warp may access
        const int m = lg * 32 + lane;            // accumulator row (UMMA M index)
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const int et = (warp - 2) * 32 + lane;   // 0..127 index among epilogue threads
        long long q = 0;
        auto t_epilogue = [&](long long it) {
            int buf = (int)(it & 1);
            mbar_wait(bar(BAR_TACC_FULL + buf), (uint32_t)((it >> 1) & 1));
            mbar_wait(bar(BAR_TSM_EMPTY + buf), (uint32_t)((it >> 1) & 1) ^ 1u);
            tc_fence_after_sync();
            // GEMM1 row order is (row, col): m = r*16 + xc.  T slot order is (col, row): (xc+1)*8 + r.
            const int r = m >> 4, xc = m & 15;
            const uint32_t row_off = (uint32_t)(((xc + 1) * kTileRows + r) * trow_b);
            const uint32_t tb = base + p.off_t + buf * p.t_buf;
            for (int c16 = 0; c16 < p.KP / 16; ++c16) {
                uint32_t v[16];
                tmem_ld16(tmem_T + buf * 64 + lane_addr + c16 * 16, v);
                tmem_ld_wait();
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                st_shared_v4(tb + swz_off(row_off + (c16 * 2) * 16, trow_b), w[0], w[1], w[2], w[3]);
                st_shared_v4(tb + swz_off(row_off + (c16 * 2 + 1) * 16, trow_b), w[4], w[5], w[6], w[7]);
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            mbar_arrive(bar(BAR_TSM_FULL + buf));
            mbar_arrive(bar(BAR_TACC_EMPTY + buf));
        };
        auto y_epilogue = [&](long long j) {
            long long tile = first + j * step;
            int b = (int)(tile / tiles_per_img);
            int rem = (int)(tile % tiles_per_img);
            int h0 = (rem / p.tiles_w) * kTileRows, w0 = (rem % p.tiles_w) * kValidCols;
            // GEMM2 row order is (col, row): m = yc*8 + r ; staging order is the TMA box order (row, col, channel)
            const int yc = m >> 3, r = m & 7;
            const bool valid = yc >= 1 && yc <= kValidCols;
            const uint32_t row_off = (uint32_t)((r * kValidCols + (yc - 1)) * 128);
            for (int nt = 0; nt < p.nnt; ++nt, ++q) {
                int yb = (int)(q & 1);
                const uint32_t sb = base + p.off_y + yb * p.y_buf;
                mbar_wait(bar(BAR_YACC_FULL + yb), (uint32_t)((q >> 1) & 1));
                tc_fence_after_sync();
                // staging buffer yb was handed to a TMA store two N-tiles ago: wait until it has been read
                if (et == 0) tma_store_wait_read<1>();
                named_bar_sync(1, 128);
#pragma unroll
                for (int c16 = 0; c16 < kNT / 16; ++c16) {
                    uint32_t v[16];
                    tmem_ld16(tmem_Y + yb * 64 + lane_addr + c16 * 16, v);
                    tmem_ld_wait();
                    if (valid) {
                        uint32_t w[8];
                        const float* bz = sbias + nt * kNT + c16 * 16;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float a = __uint_as_float(v[2 * i]) + bz[2 * i];
                            float c = __uint_as_float(v[2 * i + 1]) + bz[2 * i + 1];
                            if (p.relu) { a = fmaxf(a, 0.f); c = fmaxf(c, 0.f); }
                            w[i] = pack_bf16(a, c);
                        }
                        st_shared_v4(sb + swz_off(row_off + (c16 * 2) * 16, 128), w[0], w[1], w[2], w[3]);
                        st_shared_v4(sb + swz_off(row_off + (c16 * 2 + 1) * 16, 128), w[4], w[5], w[6], w[7]);
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(bar(BAR_YACC_EMPTY + yb));
                fence_proxy_async_smem();
                named_bar_sync(2, 128);
                if (et == 0) {
                    tma_store_4d(&tm_y, sb, nt * kNT, w0, h0, b);
                    tma_store_commit();
                }
            }
        };
        for (long long it = 0; it < my_tiles; ++it) {
            t_epilogue(it);
            if (it > 0) y_epilogue(it - 1);
        }
        if (my_tiles > 0) y_epilogue(my_tiles - 1);
        if (et == 0) tma_store_wait_all<0>();
    }

    // ---------------- teardown ----------------
    tc_fence_before_sync();
    __syncthreads();
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
static size_t plan_smem(LR2Params& p, int c_pad, size_t limit) {
    const uint32_t xrow = p.CK * 2, trow = p.KP * 2;
    p.x_stage = up1k(kXRows * kTileCols * xrow);
    p.v_chunk = up1k(3 * p.KP * xrow);
    p.h_tile = up1k(3 * kNT * trow);
    p.t_buf = up1k(kTCols * kTileRows * trow);
    p.y_buf = 16384;
    const uint32_t fixed = 2 * p.t_buf + 2 * p.y_buf + up1k(p.N * 4) + 1024 /*barriers*/ + 1024 /*align slack*/;
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
    p.N = N; p.CK = pick_ck(c_pad); p.KP = k_pad; p.nch = c_pad / p.CK; p.nnt = N / kNT;
    return plan_smem(p, c_pad, 227 * 1024) != 0 ? 1 : 0;
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
                               int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, void* stream) {
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
        CNA_CUDA(cudaFuncSetAttribute(lowrank2_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    }
    LR2Params p{};
    p.B = B; p.H = H; p.W = W; p.N = N; p.relu = relu;
    p.CK = pick_ck(c_pad); p.KP = k_pad; p.nch = c_pad / p.CK; p.nnt = N / kNT;
    p.tiles_w = ceil_div(W, kValidCols); p.tiles_h = ceil_div(H, kTileRows);
    p.ntiles = (long long)B * p.tiles_w * p.tiles_h;
    size_t smem = plan_smem(p, c_pad, (size_t)max_smem);
    CNA_CHECK(smem != 0, "shared-memory plan does not fit");

    CUtensorMap tm_x, tm_wv, tm_wh, tm_y;
    {
        cuuint64_t dims[4] = {(cuuint64_t)c_pad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)c_pad * 2, (cuuint64_t)W * c_pad * 2, (cuuint64_t)H * W * c_pad * 2};
        cuuint32_t box[4] = {(cuuint32_t)p.CK, kTileCols, kXRows, 1};
        if (encode(&tm_x, x, 4, dims, str, box, swizzle_for(p.CK * 2))) return 1;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)c_pad, (cuuint64_t)3 * k_pad};
        cuuint64_t str[1] = {(cuuint64_t)c_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)p.CK, (cuuint32_t)(3 * k_pad)};
        if (encode(&tm_wv, wv_packed, 2, dims, str, box, swizzle_for(p.CK * 2))) return 1;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)3 * N};
        cuuint64_t str[1] = {(cuuint64_t)k_pad * 2};
        cuuint32_t box[2] = {(cuuint32_t)k_pad, kNT};
        if (encode(&tm_wh, wh_packed, 2, dims, str, box, swizzle_for(k_pad * 2))) return 1;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)N * 2, (cuuint64_t)W * N * 2, (cuuint64_t)H * W * N * 2};
        cuuint32_t box[4] = {kNT, kValidCols, kTileRows, 1};
        if (encode(&tm_y, y, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    int grid = (int)std::min<long long>(p.ntiles, num_sms);
    lowrank2_fused_kernel<<<grid, kThreads, smem, as_stream(stream)>>>(tm_x, tm_wv, tm_wh, tm_y, bias, p);
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
