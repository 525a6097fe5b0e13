// Streaming packed weights into shared memory: a tiled TMA load (2-D box of 128-byte rows, 128B swizzle - what the fused conv's weight
// rings use) against a 1-D bulk copy (cp.async.bulk, contiguous bytes) of the same chunk, per SM and with all 148 SMs pulling the same
// L2-resident 384 KB.  Question: is the ~3.3 cycles per box row of the tiled path (ubench_tma) a property of TMA or of tiled boxes?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_bulk ubench_bulk.cu -lcuda
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

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

struct P { int mode, stages, rows, nchunks, iters; const uint8_t* src; long long* out; };

__global__ void __launch_bounds__(64, 1) bulk_rate(const __grid_constant__ CUtensorMap tm, P p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bars[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    auto full = [&](int i) { return smem_u32(&bars[i]); };
    auto empty = [&](int i) { return smem_u32(&bars[8 + i]); };
    const uint32_t bytes = (uint32_t)p.rows * 128u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < p.iters; ++it) {
            const int chunk = (it + blockIdx.x) % p.nchunks;
            mbar_wait(empty(s), ph ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(full(s), bytes);
                if (p.mode == 0) tma_load_2d(base + s * bytes, &tm, full(s), 0, chunk * p.rows);
                else bulk_load_1d(base + s * bytes, p.src + (size_t)chunk * bytes, bytes, full(s));
            }
            __syncwarp();
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    } else {
        int s = 0; uint32_t ph = 0;
        for (int it = 0; it < p.iters; ++it) {
            mbar_wait(full(s), ph);
            if (lane == 0) mbar_arrive(empty(s));
            if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) p.out[blockIdx.x] = clock64() - t0;
}

int main() {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const int total_rows = 3072;                 // 384 KB of 128-byte rows
    uint8_t* src;
    CK(cudaMalloc(&src, (size_t)total_rows * 128)); CK(cudaMemset(src, 1, (size_t)total_rows * 128));
    long long* out;
    CK(cudaMallocManaged(&out, 256 * sizeof(long long)));
    CK(cudaFuncSetAttribute(bulk_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int rows : {192, 128, 64}) {            // Wv chunk (24 KB), Wh slot (16 KB), 8 KB
        CUtensorMap tm;
        cuuint64_t dims[2] = {64, (cuuint64_t)total_rows};
        cuuint64_t str[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)rows};
        cuuint32_t es[2] = {1, 1};
        if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
        for (int mode : {0, 1})
            for (int stages : {2, 4, 6})
                for (int grid : {1, 148}) {
                    P p{mode, stages, rows, total_rows / rows, 2000, src, out};
                    bulk_rate<<<grid, 64, 200 * 1024>>>(tm, p);
                    CK(cudaDeviceSynchronize());
                    bulk_rate<<<grid, 64, 200 * 1024>>>(tm, p);
                    CK(cudaDeviceSynchronize());
                    double cyc = (double)out[0] / p.iters;
                    printf("%-22s chunk %5d B  stages %d  grid %3d: %7.0f cycles/chunk  %6.1f B/cycle/SM  %5.2f cycles per 128 B\n",
                           mode == 0 ? "tiled TMA 128B rows" : "1-D cp.async.bulk", rows * 128, stages, grid, cyc, rows * 128 / cyc, cyc / rows);
                }
    }
    return 0;
}
