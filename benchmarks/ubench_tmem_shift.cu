// What does tcgen05.shift.down do to a TMEM matrix (which lanes / columns move), and what does it cost?
// Question behind it: the fused scheme-2 kernel's T-epilogue makes the lane-shifted copies of T with 64 warp shuffles per thread and
// tile; one stored copy + tensor-pipe shifts would take that off the epilogue warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o benchmarks/ubench_tmem_shift benchmarks/ubench_tmem_shift.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../convnet-approximater_b200/csrc/ptx.cuh"
using namespace cna::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tmem_shift_down(uint32_t taddr) {
    asm volatile("tcgen05.shift.cta_group::1.down [%0];" ::"r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128, 1) shift_probe(uint32_t* out, long long* cyc, int nshift, int reps) {
    __shared__ uint64_t bar_store;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, m = threadIdx.x;
    const uint32_t bar = smem_u32(&bar_store);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<64>(smem_u32(&tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot, la = (uint32_t)(warp * 32) << 16;
    uint32_t v[8];
    for (int c8 = 0; c8 < 4; ++c8) {
        for (int i = 0; i < 8; ++i) v[i] = (uint32_t)(m * 100 + c8 * 8 + i);
        tmem_st8(tmem + la + c8 * 8, v);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (warp == 0) {
        if (elect_one()) {
            tc_fence_after_sync();
            t0 = clock64();
            for (int r = 0; r < reps; ++r)
                for (int s = 0; s < nshift; ++s) tmem_shift_down(tmem + 0);          // columns 0..7 (one 32-byte element per row)
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) { t1 = clock64(); cyc[0] = t1 - t0; }
    tc_fence_after_sync();
    uint32_t r16[16];
    tmem_ld16(tmem + la, r16);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[m * 32 + i] = r16[i];
    tmem_ld16(tmem + la + 16, r16);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[m * 32 + 16 + i] = r16[i];
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem);
}

int main() {
    uint32_t* out; long long* cyc;
    CK(cudaMalloc(&out, 128 * 32 * 4)); CK(cudaMalloc(&cyc, 8));
    static uint32_t h[128 * 32];
    for (int nshift : {1, 2}) {
        shift_probe<<<1, 128>>>(out, cyc, nshift, 1);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
        printf("== %d shift(s) of columns 0..7; entry = lane*100 + column as stored ==\n", nshift);
        for (int m : {0, 1, 2, 3, 15, 16, 17, 30, 31, 32, 33, 34, 63, 64, 65, 95, 96, 97, 126, 127})
            printf("lane %3d: c0 %6u c1 %6u c7 %6u | c8 %6u c15 %6u c16 %6u\n", m, h[m * 32], h[m * 32 + 1], h[m * 32 + 7], h[m * 32 + 8],
                   h[m * 32 + 15], h[m * 32 + 16]);
    }
    for (int reps : {16, 64, 256}) {
        long long c;
        shift_probe<<<1, 128>>>(out, cyc, 1, reps);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        printf("%d shifts back to back + commit: %lld cycles = %.1f per shift\n", reps, c, (double)c / reps);
    }
    return 0;
}
