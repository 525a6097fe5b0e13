// Fused scheme-1 low-rank convolution (Jaderberg et al. scheme 1) for bf16 NHWC tensors:
//     y[p, n] = sum_{c,m} D[n, c*M+m] * s[p, c*M+m] + db[n]
//     s[p, c*M+m] = bf16( sum_{i,j} x[p + (i,j), c] * w[c*M+m][i][j] + sb[c*M+m] )                     (d x d basis filters)
//                 = bf16( sum_j h[c*M+m][j] * bf16( sum_i v[c*M+m][i] * x[p + (i,j), c] ) )             (after decomp())
// Replaces LowRankExpConvV1.forward = d_conv(s_conv(x)) (reference src/approx/layers/low_rank_conv.py:44-52, SeparableConv
// :17-25): the reference materialises the (B, C*M, H, W) intermediate — M times the input — between a grouped conv and a
// 1x1 conv; here it only ever exists as 128 x 64 bf16 slabs in shared memory, produced by CUDA cores straight into the
// UMMA A-operand layout and consumed by tcgen05 (the layer is tensor-bound once fused, AI ~ C*M*N/(C+N)).
// The intermediate is rounded to bf16 exactly where the reference's bf16 modules materialise it.
//
//   tile     R = min(128 / W, H) full-width rows of the (B*H)-row image stack = up to 128 output pixels (M-tile rows
//            m = r*W + w); a tile may straddle two images: each image segment gets its own zero-filled TMA halo box
//   K-chunk  64 wide channels = 64/M input channels; per chunk the TMA warp stages the x halo box(es) (OOB zero fill = the
//            conv padding) and that chunk's basis filters (+ bias) with one bulk copy
//   warps    0: TMA producer (x + basis filters ring)   1: MMA issuer (TMEM owner)   2: TMA producer (D slab ring)
//            3-18: A producers, two groups of 8 warps on alternating chunks — thread = (8 wide channels, 4 pixels): d*d (or
//                 separable d+d) taps in packed FFMA2 with immediate-offset shared loads, bf16 pack, one 16-byte store into the
//                 128B-swizzled K-major A slab (the XOR pattern makes the stores conflict-free)
//            19-22: epilogue (TMEM -> + bias -> bf16 -> global)
//   D slabs  bf16 [N][C*M] streamed as (NT x 64) 128B-swizzled slabs through a ring of up to 8;
//            N <= 256: one N-tile and two TMEM accumulators; 256 < N <= 512: two N-tiles sharing every A slab
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace cna {
using namespace ptx;

constexpr int kL1Groups = 2;               // producer groups: group g builds the A slabs of chunks q with q % 2 == g
constexpr int kL1GroupThreads = 256;       // 8 warps per group: thread = (8 wide channels, 4 pixels)
constexpr int kL1ProdWarps = kL1Groups * kL1GroupThreads / 32;
constexpr int kL1Threads = 32 * (3 + kL1ProdWarps + 4);
constexpr int kL1SXMax = 8;               // x / basis-filter staging ring (at most; depth chosen by the smem plan)
constexpr int kL1SA = 2;                  // A slabs (one per producer group)
constexpr int kL1SWMax = 8;               // D slab ring (at most)
constexpr uint32_t kL1ABytes = 128 * 128;  // A slab: 128 pixels x 64 wide channels bf16, K-major rows of 128 bytes, 128B swizzle

struct L1Params {
    int B, H, W, C, N;
    int R, G, ntiles, nchunks;            // rows per tile, B*H, tiles, K-chunks of 64 wide channels
    int NT, nnt, nbuf, sw, sx;            // N-tile width, N-tiles, TMEM accumulator buffers, D ring depth, x ring depth
    int WP, RP, pad;                      // halo box: W + 2 pad columns, R + 2 pad rows
    int hrun;                             // 1: row-run slot mapping
    int WS;                               // slot pitch of a tile row: W (scattered mapping) or W rounded up to 4 (row runs)
    int debug;                            // CNA_L1_DEBUG ablations: 1 no taps, 2 no UMMAs, 8 no x loads, 16 no D loads
    long long* prof;                      // CNA_L1_PROF: per-role wait/work cycle counters of CTA 0 (debugging aid), else NULL
    uint32_t seg_box_bytes, seg_bytes, wblk_bytes, xs_stage;
    uint32_t off_xs, off_a, a_buf, off_w, w_sub, w_stage, off_bias, off_bar;
};

__host__ __device__ constexpr uint64_t l1_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7) << 61);
}

__device__ __forceinline__ float2 l1_ffma2(float2 a, float2 b, float2 c) {      // SASS FFMA2
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
// compile-time loop: the tap indices become immediate offsets of the shared-memory loads
template <int I, int N, typename F>
__device__ __forceinline__ void l1_static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        l1_static_for<I + 1, N>(f);
    }
}
template <int IMM>
__device__ __forceinline__ float l1_lds_bf16(uint32_t addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1+%2];" : "=h"(v) : "r"(addr), "n"(IMM));
    return __uint_as_float((uint32_t)v << 16);
}
// one row of basis-filter taps for this thread's 8 wide channels: floats [0,4) at +0, [4,8) at +128 (conflict-free layout)
template <int IMM>
__device__ __forceinline__ void l1_lds_f8(uint32_t addr, float2 (&w)[4]) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(w[0].x), "=f"(w[0].y), "=f"(w[1].x), "=f"(w[1].y) : "r"(addr), "n"(IMM));
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(w[2].x), "=f"(w[2].y), "=f"(w[3].x), "=f"(w[3].y) : "r"(addr), "n"(IMM + 128));
}
__device__ __forceinline__ void l1_sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void l1_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ uint32_t l1_pack(float2 v) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v.x, v.y);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 l1_round_bf16(float2 v) {
    const uint32_t u = l1_pack(v);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}
__device__ __forceinline__ void l1_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
        "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

enum : int {
    L1_XS_FULL = 0, L1_XS_EMPTY = L1_XS_FULL + kL1SXMax, L1_A_FULL = L1_XS_EMPTY + kL1SXMax, L1_A_EMPTY = L1_A_FULL + kL1SA,
    L1_W_FULL = L1_A_EMPTY + kL1SA, L1_W_EMPTY = L1_W_FULL + kL1SWMax, L1_ACC_FULL = L1_W_EMPTY + kL1SWMax,
    L1_ACC_EMPTY = L1_ACC_FULL + 2, L1_COUNT = L1_ACC_EMPTY + 2
};

// profiling aid: cycles one thread of each role spends in each phase, CTA 0 only
struct L1Prof {
    long long* out; long long t; bool on;
    __device__ __forceinline__ L1Prof(long long* o, bool sel) : out(o), t(0), on(o != nullptr && sel && blockIdx.x == 0) { if (on) t = clock64(); }
    __device__ __forceinline__ void lap(int slot) { if (on) { const long long n = clock64(); atomicAdd((unsigned long long*)(out + slot), (unsigned long long)(n - t)); t = n; } }
};

// D: basis filter size (d x d), M: bases per input channel (wide channel c*M+m), SEP: separable (d,1).(1,d) form
//   HRUN: a thread's 4 pixels are 4 consecutive columns of one row (slot m = r*WS + w, WS = W rounded up to 4): the D+3 inputs of
//   a row are loaded once and reused by all D taps and 4 outputs — 1.6x (5x5) fewer instructions than the scattered mapping
template <int D, int M, bool SEP, bool HRUN>
__global__ void __launch_bounds__(kL1Threads, 1)
lowrank1_fused_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                      const float* __restrict__ s_packed, const float* __restrict__ d_bias, __nv_bfloat16* __restrict__ y,
                      const L1Params p) {
    constexpr int CPC = 64 / M;                  // input channels per K-chunk
    constexpr int CBOX = CPC < 8 ? 8 : CPC;      // channels per pixel of the staged x box: TMA's inner extent is >= 16 bytes, so the
                                                 // 16-basis instance stages an aligned octet per chunk and uses one half of it
    constexpr int NCU = (M >= 8) ? 1 : 8 / M;    // input channels behind one 8-wide unit
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + p.off_bar;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    float* sbias = reinterpret_cast<float*>(base_ptr + p.off_bias);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + p.off_bar + 8 * L1_COUNT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_x); prefetch_tmap(&tm_w);
        for (int i = 0; i < kL1SXMax; ++i) { mbar_init(bar(L1_XS_FULL + i), 1); mbar_init(bar(L1_XS_EMPTY + i), kL1GroupThreads); }
        for (int i = 0; i < kL1SA; ++i) { mbar_init(bar(L1_A_FULL + i), kL1GroupThreads); mbar_init(bar(L1_A_EMPTY + i), 1); }
        for (int i = 0; i < kL1SWMax; ++i) { mbar_init(bar(L1_W_FULL + i), 1); mbar_init(bar(L1_W_EMPTY + i), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(bar(L1_ACC_FULL + i), 1); mbar_init(bar(L1_ACC_EMPTY + i), 128); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    for (int i = threadIdx.x; i < p.N; i += kL1Threads) sbias[i] = d_bias ? d_bias[i] : 0.f;
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int first = blockIdx.x, step = gridDim.x;
    const int my_tiles = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;

    if (warp == 0) {
        // ---------------- TMA producer 1: per chunk the x halo box(es) + that chunk's basis filters ----------------
        int xs = 0; uint32_t xph = 0;
        L1Prof pf(p.prof, lane == 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int g0 = (first + it * step) * p.R;
            const int b0 = g0 / p.H, h0 = g0 - b0 * p.H;
            const int R1 = min(p.R, p.H - h0);
            const bool two = (R1 < p.R) && (b0 + 1 < p.B);
            for (int ch = 0; ch < p.nchunks; ++ch) {
                pf.lap(0);
                mbar_wait(bar(L1_XS_EMPTY + xs), xph ^ 1u);
                pf.lap(1);
                if (elect_one()) {
                    const uint32_t dst = base + p.off_xs + xs * p.xs_stage;
                    const bool skip_x = (p.debug & 8) != 0;
                    mbar_arrive_expect_tx(bar(L1_XS_FULL + xs), (skip_x ? 0u : (two ? 2u : 1u) * p.seg_box_bytes) + p.wblk_bytes);
                    // first channel of the box: the chunk's own, or (16 bases: 4 channels per chunk, 8 per box) the enclosing 16-byte
                    // aligned channel octet — the producers then read its lower or upper half
                    const int c0 = CPC < 8 ? (ch * CPC) / 8 * 8 : ch * CPC;
                    if (!skip_x) tma_load_4d(dst, &tm_x, bar(L1_XS_FULL + xs), c0, -p.pad, h0 - p.pad, b0);
                    if (two && !skip_x) tma_load_4d(dst + p.seg_bytes, &tm_x, bar(L1_XS_FULL + xs), c0, -p.pad, -p.pad - R1, b0 + 1);
                    l1_bulk_load(dst + 2 * p.seg_bytes, s_packed + (size_t)ch * (p.wblk_bytes / 4), p.wblk_bytes, bar(L1_XS_FULL + xs));
                }
                __syncwarp();
                if (++xs == p.sx) { xs = 0; xph ^= 1u; }
            }
        }
    } else if (warp == 2) {
        // ---------------- TMA producer 2: the D slab of every (chunk, N-tile), its own ring so neither stream stalls the other ----------------
        int ws = 0; uint32_t wph = 0;
        L1Prof pf(p.prof, lane == 0);
        for (int it = 0; it < my_tiles; ++it) {
            for (int ch = 0; ch < p.nchunks; ++ch) {
                for (int nt = 0; nt < p.nnt; ++nt) {
                    pf.lap(20);
                    mbar_wait(bar(L1_W_EMPTY + ws), wph ^ 1u);
                    pf.lap(2);
                    if (elect_one()) {
                        const uint32_t dst = base + p.off_w + ws * p.w_stage;
                        if (p.debug & 16) {
                            mbar_arrive(bar(L1_W_FULL + ws));
                        } else {
                            mbar_arrive_expect_tx(bar(L1_W_FULL + ws), (uint32_t)p.NT * 128u);
                            tma_load_2d(dst, &tm_w, bar(L1_W_FULL + ws), ch * 64, nt * p.NT);
                        }
                    }
                    __syncwarp();
                    if (++ws == p.sw) { ws = 0; wph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | (8u << 24);
        constexpr uint64_t adsc = l1_desc_hi(16, 8 * 128, 2);           // 128-byte rows, 128B swizzle, K-major (written by the producers)
        constexpr uint64_t bdsc = l1_desc_hi(16, 8 * 128, 2);           // 128-byte rows, 128B swizzle, K-major
        int ab = 0, ws = 0; uint32_t aph = 0, wph = 0;
        L1Prof pf(p.prof, lane == 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it % p.nbuf;
            const uint32_t use = (uint32_t)(it / p.nbuf);
            pf.lap(17);
            mbar_wait(bar(L1_ACC_EMPTY + buf), (use & 1u) ^ 1u);
            pf.lap(3);
            tc_fence_after_sync();
            for (int ch = 0; ch < p.nchunks; ++ch) {
                pf.lap(17);
                mbar_wait(bar(L1_A_FULL + ab), aph);
                pf.lap(4);
                tc_fence_after_sync();
                const uint32_t a_addr = base + p.off_a + ab * p.a_buf;
                for (int nt = 0; nt < p.nnt; ++nt) {
                    pf.lap(17);
                    mbar_wait(bar(L1_W_FULL + ws), wph);
                    pf.lap(5);
                    tc_fence_after_sync();
                    const uint32_t w_addr = base + p.off_w + ws * p.w_stage;
                    const uint32_t d = tmem + buf * 256 + nt * p.NT;
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint64_t a = adsc + ((a_addr + ks * 32) >> 4);
                            const uint64_t b = bdsc + ((w_addr + ks * 32) >> 4);
                            if (!(p.debug & 2)) umma_bf16(d, a, b, idesc, (uint32_t)((ch | ks) != 0));
                        }
                        umma_commit(bar(L1_W_EMPTY + ws));
                        if (nt == p.nnt - 1) {
                            umma_commit(bar(L1_A_EMPTY + ab));
                            if (ch == p.nchunks - 1) umma_commit(bar(L1_ACC_FULL + buf));
                        }
                    }
                    __syncwarp();
                    if (++ws == p.sw) { ws = 0; wph ^= 1u; }
                }
                if (++ab == kL1SA) { ab = 0; aph ^= 1u; }
            }
        }
    } else if (warp < 3 + kL1ProdWarps) {
        // ---------------- A producers: two groups of 8 warps on alternating chunks ----------------
        const int grp = (warp - 3) >> 3;
        const int pt = threadIdx.x - 96 - grp * kL1GroupThreads;
        const int u = pt & 7, pl = pt >> 3;
        int rk[4], wk[4];
        uint32_t xoff[4], aoff[4];
        if constexpr (HRUN) {
            // 32 runs per group = (row, 4-column run); a warp's four runs sit in four different rows (row pitch % 128 B != 0, 64 B:
            // the x loads of a warp then fall into distinct banks)
            const int CR = p.WS >> 2;
            const int r_low = pl & 3, rest = pl >> 2;
            const int r_high = rest / CR, cj = rest - r_high * CR;
            const int rr = r_high * 4 + r_low;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                rk[k] = rr; wk[k] = 4 * cj + k;
                const int m = rr * p.WS + wk[k];
                xoff[k] = rr < p.R ? (uint32_t)((rr * p.WP + 4 * cj) * CBOX + u) * 2u : (uint32_t)u * 2u;      // the run's first column
                aoff[k] = (uint32_t)m * 128u + (uint32_t)((u ^ (m & 7)) * 16);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int m = pl + 32 * k;
                rk[k] = m / p.W;
                wk[k] = m - rk[k] * p.W;
                // slots past the tile (rk >= R) compute on slot 0's data and are never stored
                xoff[k] = rk[k] < p.R ? (uint32_t)((rk[k] * p.WP + wk[k]) * CBOX + (u * 8) / M) * 2u : (uint32_t)((u * 8) / M) * 2u;
                aoff[k] = (uint32_t)m * 128u + (uint32_t)((u ^ (m & 7)) * 16);      // 128B swizzle: 16-byte unit u of row m lands at u ^ (m % 8)
            }
        }
        const uint32_t row_pitch = (uint32_t)p.WP * CBOX * 2u;
        const uint32_t a_addr = base + p.off_a + grp * p.a_buf;
        uint32_t q = 0;                                   // chunk counter over this CTA's tiles
        L1Prof pf(p.prof, pt == 0);
        const int pslot = 6 + 6 * grp;                   // 6.. group 0, 12.. group 1 : XS_FULL wait, taps, A_EMPTY wait, store
        for (int it = 0; it < my_tiles; ++it) {
            const int g0 = (first + it * step) * p.R;
            const int b0 = g0 / p.H, h0 = g0 - b0 * p.H;
            const int R1 = min(p.R, p.H - h0);
            bool valid[4];
            uint32_t xo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                valid[k] = (rk[k] < p.R) && (wk[k] < p.W) && (g0 + rk[k] < p.G);
                xo[k] = xoff[k] + ((rk[k] >= R1 && rk[k] < p.R) ? p.seg_bytes : 0u);
            }
            for (int ch = 0; ch < p.nchunks; ++ch, ++q) {
                if ((int)(q & 1u) != grp) continue;
                const uint32_t xs = q % (uint32_t)p.sx, xph = (q / (uint32_t)p.sx) & 1u, aph = (q >> 1) & 1u;
                const uint32_t stage = base + p.off_xs + xs * p.xs_stage;
                const uint32_t wb = stage + 2 * p.seg_bytes + (uint32_t)u * 16u;     // this unit's floats of every tap row
                pf.lap(pslot + 4);
                mbar_wait(bar(L1_XS_FULL + xs), xph);
                pf.lap(pslot);
                uint32_t pk[4][4];
                uint32_t xr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) xr[k] = stage + xo[k];
                if constexpr (CPC < 8) {                          // this chunk's channels within the staged octet
                    const uint32_t half = (uint32_t)((ch * CPC) % 8) * 2u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) xr[k] += half;
                }
                if constexpr (HRUN && !SEP) {
                    float2 acc[4][4];
                    float2 bs[4];
                    l1_lds_f8<D * D * 256>(wb, bs);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[k][e] = bs[e];
                    if (!(p.debug & 1)) {
                        uint32_t xrow = xr[0];
                        l1_static_for<0, D>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            float2 X[D + 3];                                  // the row's inputs, each feeding up to D taps x 4 outputs
                            l1_static_for<0, D + 3>([&](auto cc) {
                                constexpr int c = decltype(cc)::value;
                                const float v = l1_lds_bf16<c * (CBOX * 2)>(xrow);
                                X[c] = make_float2(v, v);
                            });
                            l1_static_for<0, D>([&](auto jc) {
                                constexpr int j = decltype(jc)::value;
                                float2 w[4];
                                l1_lds_f8<(i * D + j) * 256>(wb, w);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
#pragma unroll
                                    for (int e = 0; e < 4; ++e) acc[k][e] = l1_ffma2(X[k + j], w[e], acc[k][e]);
                            });
                            xrow += row_pitch;
                        });
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) pk[k][e] = l1_pack(acc[k][e]);
                } else if constexpr (HRUN && SEP) {
                    float2 acc[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[k][e] = make_float2(0.f, 0.f);
                    if (!(p.debug & 1)) {
                        uint32_t xrow[D];
#pragma unroll
                        for (int i = 0; i < D; ++i) xrow[i] = xr[0] + i * row_pitch;
                        // column by column: the bf16-rounded vertical pass of column c feeds the outputs k = c - j
                        l1_static_for<0, D + 3>([&](auto cc) {
                            constexpr int c = decltype(cc)::value;
                            float2 tc[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) tc[e] = make_float2(0.f, 0.f);
                            l1_static_for<0, D>([&](auto ic) {
                                constexpr int i = decltype(ic)::value;
                                const float v = l1_lds_bf16<c * (CBOX * 2)>(xrow[i]);
                                float2 w[4];
                                l1_lds_f8<i * 256>(wb, w);                        // v taps: rows 0 .. D-1
#pragma unroll
                                for (int e = 0; e < 4; ++e) tc[e] = l1_ffma2(make_float2(v, v), w[e], tc[e]);
                            });
#pragma unroll
                            for (int e = 0; e < 4; ++e) tc[e] = l1_round_bf16(tc[e]);
                            l1_static_for<0, 4>([&](auto kc) {
                                constexpr int k = decltype(kc)::value;
                                constexpr int j = c - k;
                                if constexpr (j >= 0 && j < D) {
                                    float2 wh[4];
                                    l1_lds_f8<(D + j) * 256>(wb, wh);             // h taps: rows D .. 2D-1
#pragma unroll
                                    for (int e = 0; e < 4; ++e) acc[k][e] = l1_ffma2(tc[e], wh[e], acc[k][e]);
                                }
                            });
                        });
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) pk[k][e] = l1_pack(acc[k][e]);
                } else                 if constexpr (!SEP) {
                    float2 acc[4][4];
                    float2 bs[4];
                    l1_lds_f8<D * D * 256>(wb, bs);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[k][e] = bs[e];
                    if (!(p.debug & 1)) {
                        l1_static_for<0, D>([&](auto ic) {
                            constexpr int i = decltype(ic)::value;
                            l1_static_for<0, D>([&](auto jc) {
                                constexpr int j = decltype(jc)::value;
                                float2 w[4];
                                l1_lds_f8<(i * D + j) * 256>(wb, w);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    float xv[NCU];
                                    l1_static_for<0, NCU>([&](auto cc) {
                                        constexpr int c = decltype(cc)::value;
                                        xv[c] = l1_lds_bf16<j * (CBOX * 2) + 2 * c>(xr[k]);
                                    });
#pragma unroll
                                    for (int e = 0; e < 4; ++e) {
                                        const float x0 = xv[(2 * e) / M % NCU], x1 = xv[(2 * e + 1) / M % NCU];
                                        acc[k][e] = l1_ffma2(make_float2(x0, x1), w[e], acc[k][e]);
                                    }
                                }
                            });
#pragma unroll
                            for (int k = 0; k < 4; ++k) xr[k] += row_pitch;
                        });
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int e = 0; e < 4; ++e) pk[k][e] = l1_pack(acc[k][e]);
                } else {
                    // two pixels at a time: keeps the column accumulators t and the outputs within the register budget
                    l1_static_for<0, 2>([&](auto kc) {
                        constexpr int kp = decltype(kc)::value;
                        float2 acc[2][4];
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[kk][e] = make_float2(0.f, 0.f);
                        if (!(p.debug & 1)) {
                            l1_static_for<0, D>([&](auto jc) {
                                constexpr int j = decltype(jc)::value;
                                float2 t[2][4];
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                                    for (int e = 0; e < 4; ++e) t[kk][e] = make_float2(0.f, 0.f);
                                uint32_t xc[2] = {xr[2 * kp], xr[2 * kp + 1]};
                                l1_static_for<0, D>([&](auto ic) {
                                    constexpr int i = decltype(ic)::value;
                                    float2 w[4];
                                    l1_lds_f8<i * 256>(wb, w);                        // v taps: rows 0 .. D-1
#pragma unroll
                                    for (int kk = 0; kk < 2; ++kk) {
                                        float xv[NCU];
                                        l1_static_for<0, NCU>([&](auto cc) {
                                            constexpr int c = decltype(cc)::value;
                                            xv[c] = l1_lds_bf16<j * (CBOX * 2) + 2 * c>(xc[kk]);
                                        });
#pragma unroll
                                        for (int e = 0; e < 4; ++e) {
                                            const float x0 = xv[(2 * e) / M % NCU], x1 = xv[(2 * e + 1) / M % NCU];
                                            t[kk][e] = l1_ffma2(make_float2(x0, x1), w[e], t[kk][e]);
                                        }
                                        xc[kk] += row_pitch;
                                    }
                                });
                                float2 wh[4];
                                l1_lds_f8<(D + j) * 256>(wb, wh);                     // h taps: rows D .. 2D-1
#pragma unroll
                                for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                                    for (int e = 0; e < 4; ++e) acc[kk][e] = l1_ffma2(l1_round_bf16(t[kk][e]), wh[e], acc[kk][e]);
                            });
                        }
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                            for (int e = 0; e < 4; ++e) pk[2 * kp + kk][e] = l1_pack(acc[kk][e]);
                    });
                }
                // x stage consumed (the loads above have returned: their values fed the FMAs)
                mbar_arrive(bar(L1_XS_EMPTY + xs));
                pf.lap(pslot + 1);
                mbar_wait(bar(L1_A_EMPTY + grp), aph ^ 1u);
                pf.lap(pslot + 2);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (valid[k]) l1_sts_v4(a_addr + aoff[k], pk[k][0], pk[k][1], pk[k][2], pk[k][3]);
                fence_proxy_async_smem();
                mbar_arrive(bar(L1_A_FULL + grp));
                pf.lap(pslot + 3);
            }
        }
    } else {
        // ---------------- epilogue: TMEM -> + bias -> bf16 -> global ----------------
        const int lg = warp & 3;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const int m = lg * 32 + lane;
        const int r = m / p.WS, w = m - r * p.WS;
        L1Prof pf(p.prof, m == 0);
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it % p.nbuf;
            const uint32_t use = (uint32_t)(it / p.nbuf);
            const int g0 = (first + it * step) * p.R;
            const bool ok = (r < p.R) && (w < p.W) && (g0 + r < p.G);
            __nv_bfloat16* yrow = y + ((long long)(g0 + r) * p.W + w) * p.N;
            pf.lap(19);
            mbar_wait(bar(L1_ACC_FULL + buf), use & 1u);
            pf.lap(18);
            tc_fence_after_sync();
            for (int n0 = 0; n0 < p.N; n0 += 32) {
                uint32_t v[32];
                if (n0 + 32 <= p.N) {
                    l1_tmem_ld32(tmem + buf * 256 + lane_addr + n0, v);
                } else {   // N % 32 == 16
                    tmem_ld16(tmem + buf * 256 + lane_addr + n0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
                    for (int i = 16; i < 32; ++i) v[i] = 0;
                }
                tmem_ld_wait();
                if (n0 + 32 >= p.N) {
                    tc_fence_before_sync();
                    mbar_arrive(bar(L1_ACC_EMPTY + buf));      // accumulator drained
                }
                if (!ok) continue;
                const int nvalid = min(32, p.N - n0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {                  // 16 columns = 32 bytes per store (rows are 32-byte aligned: N % 16 == 0)
                    if (h * 16 >= nvalid) break;
                    uint4 o[2];
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int c = h * 16 + t * 8;
                        o[t].x = l1_pack(make_float2(__uint_as_float(v[c + 0]) + sbias[n0 + c + 0], __uint_as_float(v[c + 1]) + sbias[n0 + c + 1]));
                        o[t].y = l1_pack(make_float2(__uint_as_float(v[c + 2]) + sbias[n0 + c + 2], __uint_as_float(v[c + 3]) + sbias[n0 + c + 3]));
                        o[t].z = l1_pack(make_float2(__uint_as_float(v[c + 4]) + sbias[n0 + c + 4], __uint_as_float(v[c + 5]) + sbias[n0 + c + 5]));
                        o[t].w = l1_pack(make_float2(__uint_as_float(v[c + 6]) + sbias[n0 + c + 6], __uint_as_float(v[c + 7]) + sbias[n0 + c + 7]));
                    }
                    st_global_v8(yrow + n0 + h * 16, o[0], o[1]);
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc<512>(tmem);
    }
}

typedef CUresult (*EncodeTiledFnL1)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnL1 l1_get_encode() {
    static EncodeTiledFnL1 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnL1>(f);
    });
    return fn;
}
static int l1_encode(CUtensorMap* m, const void* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                     const cuuint32_t* box, CUtensorMapSwizzle swz) {
    EncodeTiledFnL1 fn = l1_get_encode();
    CNA_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CNA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

static uint32_t l1_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

// Fills the tile / shared-memory plan; returns the dynamic smem bytes (incl. alignment slack) or 0 when the layer is not covered.
static size_t l1_plan(L1Params& p, int B, int H, int W, int C, int N, int M, int d, int separable, size_t limit) {
    if (!(d == 3 || d == 5) || !(M == 2 || M == 4 || M == 8 || M == 16)) return 0;
    if (B <= 0 || H <= 0 || W <= 0 || W > 128) return 0;
    const int cpc = 64 / M;
    const int cbox = std::max(cpc, 8);          // channels per pixel of the x box (kernel: CBOX)
    if (C <= 0 || C % cpc != 0 || C % 8 != 0) return 0;
    if (N < 16 || N % 16 != 0 || N > 512 || (N > 256 && N % 32 != 0)) return 0;
    if ((long long)B * H > (1LL << 30)) return 0;
    p.B = B; p.H = H; p.W = W; p.C = C; p.N = N;
    p.pad = d / 2;
    p.R = std::min(128 / W, H);
    p.WS = W;
    p.hrun = 0;
    {   // row-run mapping: rows of WS = W rounded up to 4 slots, R a multiple of 4 (a warp's four runs sit in four rows)
        const int ws = (W + 3) / 4 * 4;
        const int rh = ws <= 32 ? (128 / ws) / 4 * 4 : 0;
        static const int force = [] { const char* e = getenv("CNA_L1_MAP"); return e ? atoi(e) : -1; }();
        const bool fits = M == 8 && rh >= 4 && rh <= H;
        // fewer instructions per pixel, but padded rows waste MMA rows: take it when the producers are the bound (separable form,
        // 5x5 filters) or when the padding costs < 5 % of the tile
        const bool wins = separable || d >= 5 || rh * W * 100 >= p.R * W * 95;
        if (fits && (force < 0 ? wins : force == 1)) { p.hrun = 1; p.R = rh; p.WS = ws; }
    }
    p.G = B * H;
    p.ntiles = (p.G + p.R - 1) / p.R;
    p.nchunks = C / cpc;
    p.nnt = N > 256 ? 2 : 1;
    p.NT = N / p.nnt;
    p.nbuf = N <= 256 ? 2 : 1;
    p.WP = W + 2 * p.pad; p.RP = p.R + 2 * p.pad;
    if (p.WP > 256 || p.RP > 256) return 0;
    p.seg_box_bytes = (uint32_t)(cbox * 2 * p.WP * p.RP);
    p.seg_bytes = l1_up(p.seg_box_bytes, 128);
    p.wblk_bytes = (uint32_t)((separable ? 2 * d : d * d + 1) * 64 * 4);
    p.xs_stage = l1_up(2 * p.seg_bytes + p.wblk_bytes, 128);
    p.a_buf = l1_up(kL1ABytes, 1024);
    p.w_sub = (uint32_t)p.NT * 64;
    p.w_stage = (uint32_t)p.NT * 128;
    const uint32_t bias_bytes = l1_up((uint32_t)N * 4, 128);
    const uint32_t fixed = kL1SA * p.a_buf + bias_bytes + 1024 /*barriers*/ + 1024 /*alignment*/ + 1024 /*x ring round-up*/;
    // ring depths: as many chunks of look-ahead as shared memory allows, balanced between the two streams (their latency, not
    // their bandwidth, is what a shallow ring exposes)
    int sw = std::min(kL1SWMax, 4 * p.nnt), sx = 6;
    while ((size_t)fixed + (size_t)sw * p.w_stage + (size_t)sx * p.xs_stage > limit) {
        if (sx > 2 && sx * p.nnt >= sw) --sx;
        else if (sw > 2) --sw;
        else if (sx > 2) --sx;
        else return 0;
    }
    p.sw = sw; p.sx = sx;
    p.off_a = 0;
    p.off_w = kL1SA * p.a_buf;
    p.off_xs = p.off_w + sw * p.w_stage;
    p.off_bias = p.off_xs + l1_up(sx * p.xs_stage, 1024);
    p.off_bar = p.off_bias + bias_bytes;
    return (size_t)p.off_bar + 1024 + 1024;
}

template <int D, int M, bool SEP, bool HRUN>
static int l1_launch(const CUtensorMap& tm_x, const CUtensorMap& tm_w, const float* s_packed, const float* d_bias, void* y,
                     const L1Params& p, size_t smem, int grid, cudaStream_t st) {
    static std::atomic<unsigned long long> attr_done{0};
    CNA_CUDA(ensure_dyn_smem(lowrank1_fused_kernel<D, M, SEP, HRUN>, 227 * 1024, attr_done));
    lowrank1_fused_kernel<D, M, SEP, HRUN><<<grid, kL1Threads, smem, st>>>(tm_x, tm_w, s_packed, d_bias, (__nv_bfloat16*)y, p);
    CNA_LAUNCH_OK();
    return 0;
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_lowrank1_fused_supported(int B, int H, int W, int C, int N, int M, int d, int separable) {
    L1Params p{};
    return l1_plan(p, B, H, W, C, N, M, d, separable, 227 * 1024) != 0 ? 1 : 0;
}

int cna_lowrank1_fused_forward(const void* x, const float* s_packed, const void* dw_packed, const float* d_bias, void* y, int B, int H,
                               int W, int C, int N, int M, int d, int separable, void* stream) {
    CNA_CHECK(x && s_packed && dw_packed && y, "bad arguments");
    L1Params p{};
    const size_t smem = l1_plan(p, B, H, W, C, N, M, d, separable, 227 * 1024);
    CNA_CHECK(smem != 0, "configuration not covered by the fused scheme-1 kernel (bf16, stride 1, 'same' padding, d in {3,5}, "
                         "M in {2,4,8,16}, C*M % 64 == 0, C % 8 == 0, N % 16 == 0, N <= 512, W <= 128)");
    DevInfo di;
    if (dev_info(&di)) return 1;
    const int num_sms = di.num_sms;
    CUtensorMap tm_x, tm_w;
    {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {(cuuint32_t)std::max(64 / M, 8), (cuuint32_t)p.WP, (cuuint32_t)p.RP, 1};   // CBOX channels
        if (l1_encode(&tm_x, x, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
    }
    {
        const cuuint64_t CM = (cuuint64_t)C * M;
        cuuint64_t dims[2] = {CM, (cuuint64_t)N};
        cuuint64_t str[1] = {CM * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.NT};
        if (l1_encode(&tm_w, dw_packed, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    const int grid = std::min(p.ntiles, num_sms);
    cudaStream_t st = as_stream(stream);
    static const int debug = [] { const char* e = getenv("CNA_L1_DEBUG"); return e ? atoi(e) : 0; }();
    p.debug = debug;
    static const bool prof_on = getenv("CNA_L1_PROF") != nullptr;
    static long long* prof_dev = nullptr;
    if (prof_on) {
        if (!prof_dev) CNA_CUDA(cudaMalloc(&prof_dev, 32 * sizeof(long long)));
        CNA_CUDA(cudaMemsetAsync(prof_dev, 0, 32 * sizeof(long long), st));
        p.prof = prof_dev;
    }
#define CNA_L1_CASE(DD, MM)                                                                                            \
    if (d == DD && M == MM && !p.hrun)                                                                                 \
        rc = separable ? l1_launch<DD, MM, true, false>(tm_x, tm_w, s_packed, d_bias, y, p, smem, grid, st)            \
                       : l1_launch<DD, MM, false, false>(tm_x, tm_w, s_packed, d_bias, y, p, smem, grid, st);
    int rc = -1;
    CNA_L1_CASE(3, 8)
    CNA_L1_CASE(5, 8)
    CNA_L1_CASE(3, 4)
    CNA_L1_CASE(5, 4)
    CNA_L1_CASE(3, 16)
    CNA_L1_CASE(5, 16)
    CNA_L1_CASE(3, 2)
    CNA_L1_CASE(5, 2)
#undef CNA_L1_CASE
    if (p.hrun && M == 8 && d == 3)
        rc = separable ? l1_launch<3, 8, true, true>(tm_x, tm_w, s_packed, d_bias, y, p, smem, grid, st)
                       : l1_launch<3, 8, false, true>(tm_x, tm_w, s_packed, d_bias, y, p, smem, grid, st);
    if (p.hrun && M == 8 && d == 5)
        rc = separable ? l1_launch<5, 8, true, true>(tm_x, tm_w, s_packed, d_bias, y, p, smem, grid, st)
                       : l1_launch<5, 8, false, true>(tm_x, tm_w, s_packed, d_bias, y, p, smem, grid, st);
    if (rc < 0) CNA_FAIL("no kernel instance for this (d, M)");
    if (rc == 0 && prof_on) {
        long long h[32];
        CNA_CUDA(cudaStreamSynchronize(st));
        CNA_CUDA(cudaMemcpy(h, prof_dev, sizeof(h), cudaMemcpyDeviceToHost));
        const int steps = ((p.ntiles + grid - 1) / grid) * p.nchunks;      // chunk-steps of CTA 0
        fprintf(stderr, "[L1PROF d=%d N=%d sep=%d hrun=%d tiles/cta=%d chunks=%d nnt=%d sx=%d sw=%d] per chunk-step cycles: xprod wait %lld other %lld | wprod wait %lld other %lld | "
                        "mma acc_empty %lld a_full %lld w_full %lld issue %lld | g0 xs_full %lld taps %lld a_empty %lld store %lld loop %lld | "
                        "g1 xs_full %lld taps %lld a_empty %lld store %lld loop %lld | epi wait %lld drain %lld (per tile)\n",
                d, N, separable, p.hrun, (p.ntiles + grid - 1) / grid, p.nchunks, p.nnt, p.sx, p.sw, h[1] / steps, h[0] / steps, h[2] / steps, h[20] / steps,
                h[3] / steps, h[4] / steps, h[5] / steps, h[17] / steps, 2 * h[6] / steps, 2 * h[7] / steps, 2 * h[8] / steps, 2 * h[9] / steps,
                2 * h[10] / steps, 2 * h[12] / steps, 2 * h[13] / steps, 2 * h[14] / steps, 2 * h[15] / steps, 2 * h[16] / steps,
                h[18] / ((p.ntiles + grid - 1) / grid), h[19] / ((p.ntiles + grid - 1) / grid));
    }
    return rc;
}

}  // extern "C"
