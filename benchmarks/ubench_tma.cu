// How fast can one SM (and all 148 together) pull the fused conv's x tiles through TMA, alone, next to TMA stores of
// the y tiles, and next to a busy tensor pipe?  Answers whether the x ring of lowrank2_fused_kernel is bound by TMA,
// by memory latency or by the consumer.   Build/run like ubench_tcgen05.cu.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../convnet-approximater_b200/csrc/ptx.cuh"
using namespace cna::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc;

static CUtensorMap make_map(void* ptr, int C, int W, int H, int B, int bc, int bw, int bh, CUtensorMapSwizzle swz) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bc, (cuuint32_t)bw, (cuuint32_t)bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

__host__ __device__ constexpr uint64_t desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7) << 61);
}

struct P {
    int tiles_w, tiles_h, B, ntiles;
    int stages, stage_bytes, nch, ck;  // x ring; nch chunks of ck channels per tile
    int bw, bh, vw, vh;                // box cols/rows, valid (stepping) cols/rows
    int do_store, store_bytes;         // also TMA-store a (64ch, vw, vh) tile per tile
    int do_mma;                        // keep the tensor pipe busy with SS N=64 UMMAs (12 per chunk)
    long long* out;
};

// warp 0: TMA producer; warp 1: consumer (waits full, optionally issues 12 UMMAs reading the stage, releases it);
// warp 2: TMA store issuer
__global__ void __launch_bounds__(96, 1) tma_rate(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y, P p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[32];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto full = [&](int i) { return smem_u32(&bars[i]); };
    auto empty = [&](int i) { return smem_u32(&bars[8 + i]); };
    const uint32_t donebar = smem_u32(&bars[16]);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
        mbar_init(donebar, 1);
        fence_mbar_init();
    }
    for (uint32_t i = threadIdx.x; i < (200u * 1024u) / 16u; i += blockDim.x)
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
    if (warp == 1) tmem_alloc<128>(smem_u32(&tmem_slot));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int first = blockIdx.x, step = gridDim.x;
    const int my = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;
    const long long t0 = clock64();
    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < my; ++it) {
            int tile = first + it * step, per = p.tiles_w * p.tiles_h;
            int b = tile / per, rem = tile % per, th = rem / p.tiles_w, tw = rem % p.tiles_w;
            for (int ch = 0; ch < p.nch; ++ch) {
                mbar_wait(empty(s), ph ^ 1u);
                if (elect_one()) {
                    mbar_arrive_expect_tx(full(s), (uint32_t)(p.bw * p.bh * p.ck * 2));
                    tma_load_4d(base + s * p.stage_bytes, &tm_x, full(s), ch * p.ck, tw * p.vw - 1, th * p.vh - 1, b);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        int s = 0; uint32_t ph = 0;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
        const uint64_t dsc = desc_hi(16, 8 * 128, 2);
        for (int it = 0; it < my; ++it)
            for (int ch = 0; ch < p.nch; ++ch) {
                mbar_wait(full(s), ph);
                tc_fence_after_sync();
                if (p.do_mma) {
                    if (elect_one()) {
                        const uint64_t a0 = dsc + ((base + s * p.stage_bytes) >> 4), b0 = dsc + ((base + 160 * 1024) >> 4);
#pragma unroll
                        for (int j = 0; j < 12 * 2; ++j)      // GEMM1 + GEMM2 worth of N=64 UMMAs per chunk
                            umma_bf16(tmem + (j >= 12 ? 64 : 0), a0 + (((j % 3) * 16 * 128 + ((j / 3) & 3) * 32) >> 4), b0 + ((((j / 3) & 3) * 32) >> 4), idesc, (uint32_t)(j % 12 != 0));
                        umma_commit(empty(s));
                    }
                    __syncwarp();
                } else if (lane == 0) {
                    mbar_arrive(empty(s));
                }
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        if (p.do_mma) {
            if (elect_one()) umma_commit(donebar);
            __syncwarp();
            mbar_wait(donebar, 0);
        }
    } else if (warp == 2 && p.do_store) {
        for (int it = 0; it < my; ++it) {
            int tile = first + it * step, per = p.tiles_w * p.tiles_h;
            int b = tile / per, rem = tile % per, th = rem / p.tiles_w, tw = rem % p.tiles_w;
            if (lane == 0) {
                tma_store_wait_read<1>();
                tma_store_4d(&tm_y, base + 176 * 1024 + (it & 1) * 16384, 0, tw * p.vw, th * p.vh, b);
                tma_store_commit();
            }
            __syncwarp();
            // pace the stores with the loads: one store per tile consumed
            // (crude: wait until the consumer has passed this tile's last chunk by polling the ring barrier phase is not
            // possible from here, so stores simply run at their own rate)
        }
        if (lane == 0) tma_store_wait_all<0>();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) p.out[blockIdx.x] = t1 - t0;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<128>(tmem);
}

int main() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    enc = (EncodeTiledFn)f;
    const int B = 64, H = 224, W = 224;
    void *x64, *x16, *y;
    CK(cudaMalloc(&x64, (size_t)B * H * W * 64 * 2)); CK(cudaMemset(x64, 0, (size_t)B * H * W * 64 * 2));
    CK(cudaMalloc(&x16, (size_t)B * H * W * 16 * 2)); CK(cudaMemset(x16, 0, (size_t)B * H * W * 16 * 2));
    CK(cudaMalloc(&y, (size_t)B * H * W * 64 * 2));
    long long* out;
    CK(cudaMallocManaged(&out, 256 * sizeof(long long)));
    CK(cudaFuncSetAttribute(tma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct Cfg { const char* name; int ck, bw, bh, vw, vh, stages, store, mma; };
    Cfg cfgs[] = {
        {"x 64ch box 16x10, 5 stages, loads only", 64, 16, 10, 14, 8, 5, 0, 0},
        {"x 64ch box 16x10, 8 stages, loads only", 64, 16, 10, 14, 8, 8, 0, 0},
        {"x 64ch box 16x10, 2 stages, loads only", 64, 16, 10, 14, 8, 2, 0, 0},
        {"x 64ch box 16x10, 5 stages, + y stores", 64, 16, 10, 14, 8, 5, 1, 0},
        {"x 64ch box 16x10, 5 stages, + 24 UMMA/tile", 64, 16, 10, 14, 8, 5, 0, 1},
        {"x 64ch box 16x10, 5 stages, + stores + UMMA", 64, 16, 10, 14, 8, 5, 1, 1},
        {"x 64ch box 16x18 (2 M-tiles), 4 stages, loads only", 64, 16, 18, 14, 16, 4, 0, 0},
        {"x 64ch box 32x10, 4 stages, loads only", 64, 32, 10, 28, 8, 4, 0, 0},
        {"x 16ch box 16x10, 8 stages, loads only", 16, 16, 10, 14, 8, 8, 0, 0},
        {"x 16ch box 16x10, 8 stages, + y stores", 16, 16, 10, 14, 8, 8, 1, 0},
    };
    for (auto& c : cfgs) {
        P p{};
        p.tiles_w = (W + c.vw - 1) / c.vw; p.tiles_h = (H + c.vh - 1) / c.vh; p.B = B; p.ntiles = B * p.tiles_w * p.tiles_h;
        p.stages = c.stages; p.stage_bytes = ((c.bw * c.bh * c.ck * 2 + 1023) / 1024) * 1024; p.nch = 1; p.ck = c.ck;
        p.bw = c.bw; p.bh = c.bh; p.vw = c.vw; p.vh = c.vh; p.do_store = c.store; p.do_mma = c.mma; p.out = out;
        CUtensorMap tx = make_map(c.ck == 64 ? x64 : x16, c.ck, W, H, B, c.ck, c.bw, c.bh, c.ck == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B);
        CUtensorMap ty = make_map(y, 64, W, H, B, 64, 14, 8, CU_TENSOR_MAP_SWIZZLE_128B);
        for (int grid : {1, 148}) {
            tma_rate<<<grid, 96, 220 * 1024>>>(tx, ty, p);   // warm-up
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            tma_rate<<<grid, 96, 220 * 1024>>>(tx, ty, p);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            long long tiles_cta = (p.ntiles + grid - 1) / grid;
            if (grid == 1) tiles_cta = p.ntiles;
            double cyc = (double)out[0];
            double ld_bytes = (double)c.bw * c.bh * c.ck * 2, st_bytes = c.store ? 14.0 * 8 * 128 : 0;
            printf("%-52s grid %3d: %8.0f cyc/tile  load %6.1f B/cyc/SM  store %5.1f B/cyc/SM  | chip %.2f TB/s fetched, %.3f ms\n", c.name, grid,
                   cyc / tiles_cta, ld_bytes * tiles_cta / cyc, st_bytes * tiles_cta / cyc,
                   (ld_bytes + st_bytes) * p.ntiles / (ms * 1e-3) / 1e12, ms);
            if (grid == 1) { /* a single CTA walks all tiles: too slow, shrink */ }
        }
    }
    return 0;
}
