// Micro-benchmarks behind the fused scheme-2 kernel's design choices: how fast one SM retires
// tcgen05.mma for the operand shapes/layouts the kernel uses, and how fast epilogue warps can drain TMEM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o benchmarks/ubench_tcgen05 benchmarks/ubench_tcgen05.cu
//   ./benchmarks/ubench_tcgen05
//
// Every test runs ONE CTA (so one SM) and reports SM cycles (clock64) per unit of work.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../convnet-approximater_b200/csrc/ptx.cuh"
using namespace cna::ptx;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__host__ __device__ constexpr uint64_t desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7) << 61);
}

__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

struct P {
    int mode;      // A operand: 0 smem 128B-swizzle K-major (x tile), 1 smem no-swizzle pitch 144 (T tile), 2 no-swizzle pitch 128, 3 TMEM
    int N;         // UMMA N
    int groups;    // groups of 12 UMMAs (3 taps x 4 k-steps), one commit per group
    int sts_kb;    // KB of st.shared.v4 traffic per group issued by 4 other warps (0 = none)
    int ldtm;      // 1: 4 other warps run tcgen05.ld loops concurrently
    int rot;       // >1: rotate the accumulator per INSTRUCTION over `rot` independent 64-column accumulators (N=64 only)
    long long* out;
    int M = 128;   // UMMA M (128, or 64: half the A rows — the transposed-GEMM option of DESIGN.md §7)
};

// one CTA, 9 warps: warp 0 issues MMAs, warps 1-4 optional smem-store noise, warps 5-8 optional TMEM-load noise
__global__ void __launch_bounds__(288, 1) umma_rate(P p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_store[2];
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int stop_flag;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar = smem_u32(&bar_store[0]);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); stop_flag = 0; }
    // zero all operand memory (A region 48 KB at base, B region 48 KB at base+48K, noise region 32 KB after)
    for (uint32_t i = threadIdx.x; i < (128u * 1024u) / 16u; i += blockDim.x)
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
        const uint64_t bdesc = desc_hi(16, 8 * 128, 2) + ((base + 48 * 1024) >> 4);      // B: 128B swizzle, up to 3 x 256 rows? keep 1 tap = N rows
        uint64_t adesc; uint32_t tap_step, kk_step;
        if (p.mode == 0) { adesc = desc_hi(16, 8 * 128, 2) + (base >> 4); tap_step = 16 * 128; kk_step = 32; }
        else if (p.mode == 1) { adesc = desc_hi(18 * 144, 144, 0) + (base >> 4); tap_step = 144; kk_step = 2 * 18 * 144; }
        else { adesc = desc_hi(18 * 128, 128, 0) + (base >> 4); tap_step = 128; kk_step = 2 * 18 * 128; }
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int g = 0; g < p.groups; ++g) {
                const uint32_t d = tmem + (p.N <= 128 ? (g & 1) * 128 : 0);
#pragma unroll
                for (int tap = 0; tap < 3; ++tap)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t b = bdesc + ((tap * 0 + kk * 32) >> 4);
                        const uint32_t dd = p.rot > 1 ? tmem + ((tap * 4 + kk) % p.rot) * 64 : d;
                        const uint32_t accf = p.rot > 1 ? (uint32_t)(g != 0) : (uint32_t)((tap | kk) != 0);
                        if (p.mode == 3) umma_ts(dd, tmem + 384 + tap * 32 + kk * 8, b, idesc, accf);
                        else umma_bf16(dd, adesc + ((tap * tap_step + kk * kk_step) >> 4), b, idesc, accf);
                    }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
        if (lane == 0) { p.out[0] = t1 - t0; stop_flag = 1; }
    } else if (warp <= 4) {
        if (p.sts_kb < 0) {
            // shared-memory LOAD noise (what CUDA-core producer warps next to the MMA issuer generate)
            const uint32_t nb = base + 96 * 1024 + (warp - 1) * 8192 + lane * 16;
            long long n = 0; uint32_t sink = 0;
            while (!stop_flag) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    uint32_t a, b, c, d;
                    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(nb + i * 512));
                    sink ^= a ^ b ^ c ^ d;
                }
                ++n;
            }
            if (lane == 0) p.out[1 + warp] = n * 8 + (sink == 0x12345 ? 1 : 0);
        } else if (p.sts_kb > 0) {
            // each warp instruction stores 512 B; noise region at base + 96 KB (32 KB)
            const uint32_t nb = base + 96 * 1024 + (warp - 1) * 8192 + lane * 16;
            long long n = 0;
            while (!stop_flag) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(nb + i * 512), "r"((uint32_t)n) : "memory");
                ++n;
                // pace: sts_kb KB per group of 12 UMMAs is what the real kernel does; here we just saturate if sts_kb >= 1000
                if (p.sts_kb < 1000) __nanosleep(p.sts_kb);
            }
            if (lane == 0) p.out[1 + warp] = n * 8;   // KB stored by this warp
        }
    } else {
        if (p.ldtm) {
            const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
            long long n = 0;
            uint32_t acc = 0;
            while (!stop_flag) {
                uint32_t v[64];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld16(tmem + 256 + lane_addr + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c * 16]));
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 64; ++c) acc ^= v[c];
                ++n;
            }
            if (lane == 0) p.out[6 + (warp & 3)] = n + (acc == 0x12345678u);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// TMEM drain rate: W warps per lane quarter each reading `cols` columns per iteration
__global__ void __launch_bounds__(512, 1) tmem_rate(int warps, int cols, int iters, int do_store, long long* out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < warps) {
        const uint32_t cbase = (uint32_t)(warp >> 2) * cols;
        for (int it = 0; it < iters; ++it) {
            uint32_t v[16];
            for (int c = 0; c < cols; c += 16) {
                if (do_store) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = it + i;
                    tmem_st16(tmem + lane_addr + cbase + c, v);
                } else {
                    tmem_ld16(tmem + lane_addr + cbase + c, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) acc ^= v[i];
                }
            }
            if (do_store) tmem_st_wait();
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) out[1] = 1;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// same, but all loads of an iteration issued before one wait (how an epilogue would do it)
__global__ void __launch_bounds__(512, 1) tmem_rate_batched(int warps, int iters, long long* out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    if (warp < warps) {
        const uint32_t cbase = (uint32_t)(warp >> 2) * 64;
        for (int it = 0; it < iters; ++it) {
            uint32_t v[64];
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_ld16(tmem + lane_addr + cbase + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&v[c * 16]));
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 64; ++i) acc ^= v[i];
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) out[1] = 1;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}


// ---- cta_group::2: one CTA pair (cluster of 2), M = 256 (128 rows per CTA), B split N/2 rows per CTA ----
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc),
        "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc),
        "r"(acc), "r"(0u)
        : "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma2_rate(int mode, int N, int groups, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_store[2];
    __shared__ uint32_t tmem_slot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar = smem_u32(&bar_store[0]);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (uint32_t i = threadIdx.x; i < (128u * 1024u) / 16u; i += blockDim.x)
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0 && rank == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (16u << 24);   // M = 256
        const uint64_t bdesc = desc_hi(16, 8 * 128, 2) + ((base + 48 * 1024) >> 4);
        const uint64_t adesc = desc_hi(16, 8 * 128, 2) + (base >> 4);
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                const uint32_t d = tmem + (N <= 128 ? (g & 1) * 128 : 0);
#pragma unroll
                for (int tap = 0; tap < 3; ++tap)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint64_t b = bdesc + ((kk * 32) >> 4);
                        if (mode == 3) umma2_ts(d, tmem + 384 + tap * 32 + kk * 8, b, idesc, (uint32_t)((tap | kk) != 0));
                        else umma2_ss(d, adesc + ((tap * 16 * 128 + kk * 32) >> 4), b, idesc, (uint32_t)((tap | kk) != 0));
                    }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
        if (lane == 0) out[0] = t1 - t0;
    }
    tc_fence_before_sync();
    cluster_sync_all();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

// W issuing warps, each with its own accumulator chain (groups of 12 into one 64/128/256-column accumulator)
__global__ void __launch_bounds__(128, 1) umma_multi_issuer(int W, int N, int mode, int groups, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar_store[4];
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar_store[i]), 1); fence_mbar_init(); }
    for (uint32_t i = threadIdx.x; i < (128u * 1024u) / 16u; i += blockDim.x)
        asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < W) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
        const uint64_t bdesc = desc_hi(16, 8 * 128, 2) + ((base + 48 * 1024) >> 4);
        const uint64_t adesc = desc_hi(16, 8 * 128, 2) + ((base + warp * 2048) >> 4);
        const uint32_t d = tmem + warp * (N <= 128 ? N : 0);
        const uint32_t bar = smem_u32(&bar_store[warp]);
        if (elect_one()) {
            for (int g = 0; g < groups; ++g) {
#pragma unroll
                for (int tap = 0; tap < 3; ++tap)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (mode == 3) umma_ts(d, tmem + 384 + tap * 32 + kk * 8, bdesc + ((kk * 32) >> 4), idesc, (uint32_t)((tap | kk) != 0));
                        else umma_bf16(d, adesc + ((tap * 16 * 128 + kk * 32) >> 4), bdesc + ((kk * 32) >> 4), idesc, (uint32_t)((tap | kk) != 0));
                    }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    (void)lane;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// TMEM store shapes: 4 warps each writing 96 columns per iteration (the three bf16 copies of T) with x8 / x16 / x32 / x64 stores
template <int X>
__device__ __forceinline__ void tmem_stX(uint32_t taddr, const uint32_t* v) {
    if constexpr (X == 8) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                     "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    } else if constexpr (X == 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(v[0]),
                     "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                     "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
                     "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
                     "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
                     "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
    }
}
template <int X>
__global__ void __launch_bounds__(256, 1) tmem_st_shapes(int warps, int iters, long long* out) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc<512>(smem_u32(&tmem_slot));
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    if (warp < warps) {
        const uint32_t cbase = (uint32_t)(warp >> 2) * 96;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 96; c += X) tmem_stX<X>(tmem + lane_addr + cbase + c, v);
            tmem_st_wait();
            v[it & 31] += it;
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

int main() {
    long long* out;
    CK(cudaMallocManaged(&out, 64 * sizeof(long long)));
    CK(cudaFuncSetAttribute(umma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char* mname[4] = {"A smem sw128 (x tile)", "A smem noswz pitch144 (T tile)", "A smem noswz pitch128", "A in TMEM"};
    printf("== tcgen05.mma M=128 K=16 bf16, groups of 12 (3 taps x 4 k-steps), one SM ==\n");
    const int Ns[3] = {64, 128, 256};
    for (int mode = 0; mode < 4; ++mode)
        for (int ni = 0; ni < 3; ++ni)
            for (int noise = 0; noise < 3; ++noise) {
                P p{mode, Ns[ni], 400, noise == 1 ? 1000 : 0, noise == 2 ? 1 : 0, 1, out};
                for (int i = 0; i < 16; ++i) out[i] = 0;
                umma_rate<<<1, 288, 200 * 1024>>>(p);
                CK(cudaDeviceSynchronize());
                double per = (double)out[0] / (p.groups * 12.0);
                double floor_ = 128.0 * p.N / 256.0;
                printf("%-32s N=%3d %-12s : %7.1f cyc/UMMA (MMA floor %5.1f, %5.1f%% of tensor peak)", mname[mode], p.N,
                       noise == 0 ? "alone" : (noise == 1 ? "+STS flood" : "+LDTM x4"), per, floor_, 100.0 * floor_ / per);
                if (noise == 1) { double kb = out[2] + out[3] + out[4] + out[5]; printf("  STS %.1f B/cyc", kb * 1024.0 / out[0]); }
                if (noise == 2) { double n = out[6] + out[7] + out[8] + out[9]; printf("  LDTM %.1f B/cyc", n * 32 * 64 * 4 / out[0]); }
                printf("\n");
            }
    printf("== shared-memory LOAD flood from 4 other warps (ld.shared.v4) ==\n");
    for (int mode = 0; mode < 3; ++mode)
        for (int ni = 0; ni < 3; ++ni) {
            P p{mode, Ns[ni], 400, -1, 0, 1, out};
            for (int i = 0; i < 16; ++i) out[i] = 0;
            umma_rate<<<1, 288, 200 * 1024>>>(p);
            CK(cudaDeviceSynchronize());
            double kb = out[2] + out[3] + out[4] + out[5];
            printf("%-32s N=%3d +LDS flood   : %7.1f cyc/UMMA  LDS %.1f B/cyc\n", mname[mode], p.N, (double)out[0] / (p.groups * 12.0), kb * 1024.0 / out[0]);
        }
    printf("== M = 64 (weights as the A operand of a transposed GEMM: 64 ranks x 16, pixels along N) ==\n");
    for (int mode = 0; mode < 4; mode += 3)
        for (int ni = 0; ni < 3; ++ni) {
            P p{mode, Ns[ni], 400, 0, 0, 1, out, 64};
            for (int i = 0; i < 16; ++i) out[i] = 0;
            umma_rate<<<1, 288, 200 * 1024>>>(p);
            CK(cudaDeviceSynchronize());
            double per = (double)out[0] / (p.groups * 12.0);
            printf("%-32s M= 64 N=%3d : %7.1f cyc/UMMA (%5.1f%% of tensor peak on its 64 x N x 16 MACs)\n", mname[mode], p.N, per, 100.0 * (64.0 * p.N / 256.0) / per);
        }
    printf("== independent accumulators rotated per instruction (N=64) ==\n");
    for (int mode = 0; mode < 4; mode += 3)
        for (int rot = 1; rot <= 4; rot *= 2) {
            P p{mode, 64, 400, 0, 0, rot, out};
            umma_rate<<<1, 288, 200 * 1024>>>(p);
            CK(cudaDeviceSynchronize());
            printf("%-32s N= 64 rot %d : %7.1f cyc/UMMA\n", mname[mode], rot, (double)out[0] / (p.groups * 12.0));
        }
    printf("== several issuing warps, one accumulator chain each ==\n");
    CK(cudaFuncSetAttribute(umma_multi_issuer, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int mode = 0; mode < 4; mode += 3)
        for (int N = 64; N <= 128; N *= 2)
            for (int W = 1; W <= 4; ++W) {
                if (N == 128 && W > 2 && mode == 3) continue;
                umma_multi_issuer<<<1, 128, 200 * 1024>>>(W, N, mode, 400, out);
                CK(cudaDeviceSynchronize());
                printf("%-24s N=%3d  %d issuing warps: %7.1f cyc per UMMA overall (%5.1f%% of tensor peak)\n", mname[mode], N, W,
                       (double)out[0] / (400 * 12.0 * W), 100.0 * (128.0 * N / 256.0) * 400 * 12 * W / out[0]);
            }
    printf("== TMEM store shapes, 96 columns per warp per iteration ==\n");
    for (int warps = 4; warps <= 8; warps *= 2) {
        tmem_st_shapes<8><<<1, 256>>>(warps, 2000, out); CK(cudaDeviceSynchronize());
        printf("st x8  %d warps: %.1f cyc/iter (%.1f B/cyc/SM)\n", warps, out[0] / 2000.0, warps * 32 * 96 * 4 * 2000.0 / out[0]);
        tmem_st_shapes<16><<<1, 256>>>(warps, 2000, out); CK(cudaDeviceSynchronize());
        printf("st x16 %d warps: %.1f cyc/iter (%.1f B/cyc/SM)\n", warps, out[0] / 2000.0, warps * 32 * 96 * 4 * 2000.0 / out[0]);
        tmem_st_shapes<32><<<1, 256>>>(warps, 2000, out); CK(cudaDeviceSynchronize());
        printf("st x32 %d warps: %.1f cyc/iter (%.1f B/cyc/SM)\n", warps, out[0] / 2000.0, warps * 32 * 96 * 4 * 2000.0 / out[0]);
    }
    printf("== TMEM drain: tcgen05.ld 32x32b.x16 + wait per 16 columns ==\n");
    for (int warps = 4; warps <= 16; warps *= 2) {
        tmem_rate<<<1, 512>>>(warps, 64, 2000, 0, out);
        CK(cudaDeviceSynchronize());
        printf("ld  %2d warps x 64 cols: %.1f cyc/iter, %.1f B/cyc/SM\n", warps, out[0] / 2000.0, warps * 32 * 64 * 4 * 2000.0 / out[0]);
        tmem_rate_batched<<<1, 512>>>(warps, 2000, out);
        CK(cudaDeviceSynchronize());
        printf("ld  %2d warps x 64 cols batched (4 ld, 1 wait): %.1f cyc/iter, %.1f B/cyc/SM\n", warps, out[0] / 2000.0, warps * 32 * 64 * 4 * 2000.0 / out[0]);
        tmem_rate<<<1, 512>>>(warps, 64, 2000, 1, out);
        CK(cudaDeviceSynchronize());
        printf("st  %2d warps x 64 cols: %.1f cyc/iter, %.1f B/cyc/SM\n", warps, out[0] / 2000.0, warps * 32 * 64 * 4 * 2000.0 / out[0]);
    }
    printf("== tcgen05.mma cta_group::2, M=256 (CTA pair), groups of 12 ==\n");
    CK(cudaFuncSetAttribute(umma2_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int mode = 0; mode <= 3; mode += 3)
        for (int ni = 0; ni < 3; ++ni) {
            out[0] = 0;
            umma2_rate<<<2, 128, 200 * 1024>>>(mode, Ns[ni], 400, out);
            CK(cudaDeviceSynchronize());
            double per = (double)out[0] / (400 * 12.0);
            double floor_ = 256.0 * Ns[ni] / 512.0;
            printf("2-CTA %-22s N=%3d : %7.1f cyc/UMMA (MMA floor %5.1f, %5.1f%% of the pair's tensor peak)\n", mode == 0 ? "A smem sw128" : "A in TMEM",
                   Ns[ni], per, floor_, 100.0 * floor_ / per);
        }
    return 0;
}
