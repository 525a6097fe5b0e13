// Weight gradient of a 3-tap 'same' convolution on tcgen05, for the layer-wise reconstruction step (reference: autograd's
// conv backward-filter reached from total_loss.backward(), src/approx/hooks/finetune.py:329-330):
//
//     D_s[a][k] = sum over pixels p of  A[p + s*e_axis][a] * Bm[p][k],      s in {-1, 0, +1}
//
// A  = the WIDE tensor [B,H,W,Ca] (x for v_conv's gradient, dL/dy for h_conv's), Bm = the 64-channel one ([B,H,W,64]:
// dL/dT resp. T).  The reduction runs over PIXELS, so both operands are "MN-major" for the tensor core: a pixel's channels
// are contiguous in memory.  That is exactly what a TMA box (64 channels x pixels) with the 128-byte swizzle lands in shared
// memory — the canonical MN-major SW128 UMMA layout with K = pixel slot — so the NHWC tensors feed tcgen05 with no transpose.
//
//   tile      = 8 rows x 16 columns of one image = 128 pixel slots (slot = row*16 + col); A is loaded with one halo row above
//               and below (10 rows, OOB zero fill = the conv's zero padding), Bm without.
//   K-step    = 16 slots = one tile row: UMMA #kk of tap s reads A rows (1 + s + kk) and Bm row kk — tap shifts are start-address
//               shifts of whole 2 KB rows, like the forward kernel's vertical taps.  Taps along W use tensor maps that present
//               the image transposed (tile rows run along W).
//   CTA       = one group of up to 256 A-channels (2 M-blocks of 128) x a slice of the pixel tiles; 3 taps x 2 M-blocks x 64
//               fp32 columns = 384 TMEM columns accumulate over all its tiles; at the end each thread red.adds its lane
//               (= channel) to the fp32 gradient in the reference's OIHW layout (split-K over CTAs by atomics).
//   warp 0 TMA producer (2-stage ring), warp 1 MMA issuer, warps 2-5 final epilogue.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <cmath>
#include <mutex>

namespace cna {
using namespace ptx;

constexpr int kWgThreads = 192;
constexpr int kWgStages = 2;
constexpr uint32_t kWgASub = 10 * 16 * 128;      // one 64-channel sub-block of an A tile: 160 slots x 128 B
constexpr uint32_t kWgBTile = 8 * 16 * 128;      // Bm tile: 128 slots x 128 B

struct WgParams {
    int B, KH, KW;          // image count and dims along the tile-row / tile-col axes
    int tiles_h, tiles_w, ntiles;
    int nsub;               // 64-channel sub-blocks of A handled by this launch's channel groups (per group: 2 or 4)
    int groups;             // channel groups (gridDim.y)
    int Ca;                 // real A channels (rows beyond it are not written)
    int kb;                 // real Bm channels (<= 64)
    long long s_a, s_k;     // output strides (in floats) of the A-channel and Bm-channel index
    int flip;               // tap index written: flip ? 1 - s : 1 + s
    uint32_t a_stage;       // bytes of one A stage (nsub sub-blocks)
    int vec4;               // output rows are contiguous per lane and 16-byte aligned: use vector reductions in the epilogue
};

__host__ __device__ constexpr uint64_t wg_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);   // SWIZZLE_128B
}

enum : int { WG_FULL = 0, WG_EMPTY = WG_FULL + kWgStages, WG_DONE = WG_EMPTY + kWgStages, WG_COUNT = WG_DONE + 1 };

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, float* __restrict__ out,
                 const WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t off_b = kWgStages * p.a_stage;
    const uint32_t off_bar = off_b + kWgStages * kWgBTile;
    const uint32_t bars = base + off_bar;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + off_bar + 8 * WG_COUNT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = blockIdx.y;                       // channel group: sub-blocks [grp*nsub, grp*nsub + nsub)
    const int nblk = p.nsub / 2;                      // 128-channel M-blocks in this group (1 or 2)

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a); prefetch_tmap(&tm_b);
        for (int i = 0; i < kWgStages; ++i) { mbar_init(bar(WG_FULL + i), 1); mbar_init(bar(WG_EMPTY + i), 1); }
        mbar_init(bar(WG_DONE), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int first = blockIdx.x, step = gridDim.x;
    const int my_tiles = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;

    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = first + it * step, per = p.tiles_h * p.tiles_w;
            const int b = tile / per, rem = tile - b * per, th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
            mbar_wait(bar(WG_EMPTY + s), ph ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(bar(WG_FULL + s), (uint32_t)p.nsub * kWgASub + kWgBTile);
                for (int sb = 0; sb < p.nsub; ++sb)
                    tma_load_4d(base + s * p.a_stage + sb * kWgASub, &tm_a, bar(WG_FULL + s), (grp * p.nsub + sb) * 64, tw * 16, th * 8 - 1, b);
                tma_load_4d(base + off_b + s * kWgBTile, &tm_b, bar(WG_FULL + s), 0, tw * 16, th * 8, b);
            }
            __syncwarp();
            if (++s == kWgStages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // idesc: fp32 accumulate, bf16 A/B, A and B both MN-major (bits 15, 16), N = 64, M = 128
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
        constexpr uint64_t adesc = wg_desc_hi(kWgASub, 1024);     // LBO = next 64-channel sub-block, SBO = next 8 slots
        constexpr uint64_t bdesc = wg_desc_hi(kWgBTile, 1024);
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < my_tiles; ++it) {
            mbar_wait(bar(WG_FULL + s), ph);
            tc_fence_after_sync();
            const uint32_t a_base = base + s * p.a_stage, b_base = base + off_b + s * kWgBTile;
            if (elect_one()) {
                for (int blk = 0; blk < nblk; ++blk)
#pragma unroll
                    for (int t = 0; t < 3; ++t) {                 // t = s + 1: A rows (t + kk) against Bm row kk
                        const uint32_t d = tmem + (uint32_t)(blk * 3 + t) * 64;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk)
                            umma_bf16(d, adesc + ((a_base + blk * 2 * kWgASub + (t + kk) * 2048) >> 4), bdesc + ((b_base + kk * 2048) >> 4),
                                      idesc, (uint32_t)((it | kk) != 0));
                    }
                umma_commit(bar(WG_EMPTY + s));
                if (it == my_tiles - 1) umma_commit(bar(WG_DONE));
            }
            __syncwarp();
            if (++s == kWgStages) { s = 0; ph ^= 1u; }
        }
    } else if (my_tiles > 0) {
        const int lg = warp & 3;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        mbar_wait(bar(WG_DONE), 0);
        tc_fence_after_sync();
        for (int blk = 0; blk < nblk; ++blk) {
            const int a = (grp * p.nsub + blk * 2) * 64 + lg * 32 + lane;      // A channel of this lane
            if (p.vec4) {
                // h_conv layout dw[n][k][tap]: a lane's 64 x 3 results are ONE contiguous 768-byte row -> 16-byte vector reductions
                // (red.global.add.v4.f32), a quarter of the scalar atomics.  Output element k*3 + tap comes from accumulator t = 2 - tap.
                for (int c16 = 0; c16 < 4; ++c16) {
                    uint32_t v0[16], v1[16], v2[16];
                    tmem_ld16(tmem + (uint32_t)(blk * 3 + 0) * 64 + lane_addr + c16 * 16, v0);
                    tmem_ld16(tmem + (uint32_t)(blk * 3 + 1) * 64 + lane_addr + c16 * 16, v1);
                    tmem_ld16(tmem + (uint32_t)(blk * 3 + 2) * 64 + lane_addr + c16 * 16, v2);
                    tmem_ld_wait();
                    if (a < p.Ca) {
                        float* row = out + (long long)a * 192 + c16 * 48;
                        float e[48];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            e[i * 3 + 0] = __uint_as_float(v2[i]);      // tap 0 <- accumulator 2
                            e[i * 3 + 1] = __uint_as_float(v1[i]);
                            e[i * 3 + 2] = __uint_as_float(v0[i]);
                        }
#pragma unroll
                        for (int j = 0; j < 12; ++j)
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + 4 * j), "f"(e[4 * j]), "f"(e[4 * j + 1]),
                                         "f"(e[4 * j + 2]), "f"(e[4 * j + 3])
                                         : "memory");
                    }
                }
                continue;
            }
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const int tap = p.flip ? 2 - t : t;                             // s = t - 1; flip: 1 - s
                for (int c16 = 0; c16 < 4; ++c16) {
                    uint32_t v[16];
                    tmem_ld16(tmem + (uint32_t)(blk * 3 + t) * 64 + lane_addr + c16 * 16, v);
                    tmem_ld_wait();
                    if (a < p.Ca) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int k = c16 * 16 + i;
                            if (k < p.kb) atomicAdd(out + (long long)a * p.s_a + (long long)k * p.s_k + tap, __uint_as_float(v[i]));
                        }
                    }
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

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW wg_get_encode() {
    static EncodeTiledFnW fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFnW>(f);
    });
    return fn;
}

// 4-D map (channels, tile-col axis, tile-row axis, image) of an NHWC bf16 tensor with `ch` channels; axis 1 swaps H and W
static int wg_encode(CUtensorMap* m, const void* ptr, int ch, int B, int H, int W, int axis, int box_rows) {
    EncodeTiledFnW fn = wg_get_encode();
    CNA_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t s_h = (cuuint64_t)W * ch * 2, s_w = (cuuint64_t)ch * 2;
    cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)(axis == 0 ? W : H), (cuuint64_t)(axis == 0 ? H : W), (cuuint64_t)B};
    cuuint64_t str[3] = {axis == 0 ? s_w : s_h, axis == 0 ? s_h : s_w, (cuuint64_t)H * W * ch * 2};
    cuuint32_t box[4] = {64, 16, (cuuint32_t)box_rows, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CNA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

}  // namespace cna

using namespace cna;

extern "C" {

static int wgrad3_impl(const void* a, const void* bm, float* dw, int B, int H, int W, int ca_stride, int Ca, int kb_stride, int kb, int axis,
                       int wide_is_input, bool zero_first, void* stream) {
    CNA_CHECK(a && bm && dw && B > 0 && H > 0 && W > 0, "bad arguments");
    CNA_CHECK(ca_stride >= 8 && ca_stride % 8 == 0 && Ca >= 1 && Ca <= ca_stride, "wide tensor: channel stride must be a multiple of 8");
    CNA_CHECK(kb_stride >= 8 && kb_stride % 8 == 0 && kb_stride <= 64 && kb >= 1 && kb <= kb_stride, "narrow tensor: at most 64 channels");
    CNA_CHECK(axis == 0 || axis == 1, "axis must be 0 or 1");
    DevInfo di;
    if (dev_info(&di)) return 1;
    const int num_sms = di.num_sms, max_smem = di.max_smem;
    static std::atomic<unsigned long long> attr_done{0};
    CNA_CUDA(ensure_dyn_smem(wgrad3_tc_kernel, max_smem, attr_done));
    WgParams p{};
    // 64-channel sub-blocks, rounded up to whole 128-channel M-blocks: channels past the tensor's extent are TMA zero fill
    const int nsub_total = 2 * ((ca_stride + 127) / 128);
    // channel groups of 4 sub-blocks (256 channels) when possible, else 2 (128)
    p.nsub = nsub_total % 4 == 0 ? 4 : 2;
    p.groups = nsub_total / p.nsub;
    p.B = B; p.KH = axis == 0 ? H : W; p.KW = axis == 0 ? W : H;
    p.tiles_h = ceil_div(p.KH, 8); p.tiles_w = ceil_div(p.KW, 16);
    long long nt = (long long)B * p.tiles_h * p.tiles_w;
    CNA_CHECK(nt < (1LL << 31), "too many tiles");
    p.ntiles = (int)nt;
    p.Ca = Ca; p.kb = kb;
    p.a_stage = (uint32_t)p.nsub * kWgASub;
    if (wide_is_input) {        // v_conv: dw[k][c][t] = D_{t-1}[c][k]  -> (K, C, 3, 1)
        p.s_a = 3; p.s_k = (long long)Ca * 3; p.flip = 0;
    } else {                    // h_conv: dw[n][k][t] = D_{1-t}[n][k]  -> (N, K, 1, 3)
        p.s_a = (long long)kb * 3; p.s_k = 3; p.flip = 1;
    }
    p.vec4 = (!wide_is_input && kb == 64 && (reinterpret_cast<uintptr_t>(dw) & 15) == 0) ? 1 : 0;
    const size_t smem = (size_t)kWgStages * (p.a_stage + kWgBTile) + 1024 + 1024;
    CNA_CHECK(smem <= (size_t)max_smem, "shared-memory plan does not fit");
    CUtensorMap tm_a, tm_b;
    if (wg_encode(&tm_a, a, ca_stride, B, H, W, axis, 10)) return 1;
    if (wg_encode(&tm_b, bm, kb_stride, B, H, W, axis, 8)) return 1;
    if (zero_first) CNA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)Ca * kb * 3, as_stream(stream)));
    // CTAs per channel group.  Every CTA ends with the split-K reduction of its accumulators — nsub/2 x 3 taps x 64 columns x 128
    // lanes fp32 atomics (49 152 per 256-channel group) whatever the number of tiles it summed — so with few tiles (small per-GPU
    // batches: the 8-GPU reconstruction step runs 32 images per rank) the epilogue, not the MMAs, sets the kernel's time
    // (88 us for a 512-channel layer at 28 x 28 x 32 images, 5x the batch-256 time / 8).  Balance the two terms:
    // T(g) ~ ntiles / g * t_tile + g * groups * atomics / rate, t_tile ~ 1 us, rate ~ 83 atomics/ns  ->  g* = sqrt(1.7 ntiles / groups).
    const int g_cap = std::max(1, num_sms / p.groups);
    int gx = (int)std::lround(std::sqrt(1.7 * (double)p.ntiles / p.groups));
    gx = std::max(1, std::min(gx, std::min(p.ntiles, g_cap)));
    dim3 grid((unsigned)gx, (unsigned)p.groups);
    wgrad3_tc_kernel<<<grid, kWgThreads, smem, as_stream(stream)>>>(tm_a, tm_b, dw, p);
    CNA_LAUNCH_OK();
    return 0;
}

/* dw (fp32, reference OIHW layout, overwritten) of a 3-tap 'same' conv, see include/cna_b200.h */
int cna_wgrad3_tc(const void* a, const void* bm, float* dw, int B, int H, int W, int ca_stride, int Ca, int kb_stride, int kb, int axis,
                  int wide_is_input, void* stream) {
    return wgrad3_impl(a, bm, dw, B, H, W, ca_stride, Ca, kb_stride, kb, axis, wide_is_input, true, stream);
}

/* same, ACCUMULATING into dw (the caller zeroed it: one memset for the flat gradient buffer of all layers) */
int cna_wgrad3_tc_acc(const void* a, const void* bm, float* dw, int B, int H, int W, int ca_stride, int Ca, int kb_stride, int kb, int axis,
                      int wide_is_input, void* stream) {
    return wgrad3_impl(a, bm, dw, B, H, W, ca_stride, Ca, kb_stride, kb, axis, wide_is_input, false, stream);
}

}  // extern "C"
