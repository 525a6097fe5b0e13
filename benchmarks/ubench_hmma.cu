// Micro-benchmark: legacy tensor path on sm_100a — mma.sync.m16n8k16 (bf16, fp32 accumulate) issue rate per SM sub-partition,
// alone, fed by ldmatrix.x4 from shared memory, and next to FFMA2.  Decides whether the MSCA strip convolutions (banded
// Toeplitz products, one channel at a time) pay on mma.sync instead of the FP32 pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_hmma ubench_hmma.cu && ./ubench_hmma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// mode 0: ILP independent MMAs per warp, operands in registers
// mode 1: each MMA preceded by its own ldmatrix.x4 (A from smem)
// mode 2: mode 1 + the C fragment packed to bf16 and stored to smem every MMA (the strip kernels' inner loop shape)
template <int ILP, int MODE>
__global__ void hmma_kernel(int iters, float* out, long long* cyc) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
    __syncthreads();
    float d[ILP][4];
    uint32_t a[ILP][4], b[2] = {0x3f803f80u, 0x3f803f80u};
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        for (int j = 0; j < 4; ++j) { d[i][j] = 0.f; a[i][j] = 0x3c003c00u + lane + i; }
    }
    // row pitch 80 bytes: 8 rows of 16 bytes hit 8 distinct 16-byte bank groups (conflict-free ldmatrix)
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(smem) + warp * 2048 + (lane & 15) * 80 + (lane >> 4) * 16;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE >= 1) ldsm4(a[i], sbase + ((it + i) & 7) * 16);
            mma16816(d[i], a[i], b);
            if (MODE == 2) {
                uint32_t p0, p1;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p0) : "f"(d[i][1]), "f"(d[i][0]));
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p1) : "f"(d[i][3]), "f"(d[i][2]));
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + 16384 + i * 8), "r"(p0));
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + 16384 + i * 8 + 4), "r"(p1));
                d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// m16n8k8 (bf16): is the half-K instruction half the pipe time?  MODE 0: k8 only; MODE 1: one k16 + two k8 alternating
__device__ __forceinline__ void mma1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
template <int MODE>
__global__ void k8_kernel(int iters, float* out, long long* cyc) {
    const int lane = threadIdx.x & 31;
    float d[4][4];
    uint32_t a[4], b[2] = {0x3f803f80u, 0x3f803f80u};
    for (int j = 0; j < 4; ++j) a[j] = 0x3c003c00u + lane;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (MODE == 0) { mma1688(d[i], a[0], a[1], b[0]); }
            else if (i & 1) { mma1688(d[i], a[2], a[3], b[0]); mma1688(d[i], a[0], a[1], b[1]); }
            else mma16816(d[i], a, b);
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 4; ++i) s += d[i][0] + d[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FFMA2 and HMMA interleaved: R packed FMAs per MMA, to see whether the two pipes overlap
template <int R>
__global__ void mix_kernel(int iters, float* out, long long* cyc) {
    const int lane = threadIdx.x & 31;
    float d[4][4];
    uint32_t a[4], b[2] = {0x3f803f80u, 0x3f803f80u};
    for (int j = 0; j < 4; ++j) a[j] = 0x3c003c00u + lane;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
    unsigned long long f[8];
    for (int i = 0; i < 8; ++i) f[i] = 0x3f8000003f800000ull + lane + i;
    const unsigned long long w = 0x3f7fff003f7fff00ull;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mma16816(d[i], a, b);
#pragma unroll
            for (int r = 0; r < R; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(f[(i * R + r) & 7]) : "l"(w));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 4; ++i) s += d[i][0] + d[i][3];
    for (int i = 0; i < 8; ++i) s += __uint_as_float((uint32_t)f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char* name, K kern, int threads, int iters, int per_iter, size_t smem) {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    if (smem) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<148, threads, smem>>>(iters, out, cyc);
    kern<<<148, threads, smem>>>(iters, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += (double)h[i]; c /= 148;
    const int warps = threads / 32;
    const double per_warp = c / ((double)iters * per_iter);                 // cycles between MMAs of one warp
    const double per_smsp = per_warp / ((warps + 3) / 4);                    // cycles per MMA per sub-partition
    printf("%-44s warps/SM %2d  cycles %.0f  cyc/MMA/warp %.2f  cyc/MMA/SMSP %.2f  (%s)\n", name, warps, c, per_warp, per_smsp, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    const int it = 2000;
    for (int threads : {128, 256, 512}) {
        run("mma.sync m16n8k16 bf16, regs, ILP4", hmma_kernel<4, 0>, threads, it, 4, 49152);
        run("mma.sync m16n8k16 bf16, regs, ILP8", hmma_kernel<8, 0>, threads, it, 8, 49152);
        run("ldmatrix.x4 + mma, ILP4", hmma_kernel<4, 1>, threads, it, 4, 49152);
        run("ldmatrix.x4 + mma + cvt/st, ILP4", hmma_kernel<4, 2>, threads, it, 4, 49152);
        run("m16n8k8 only (per k8 MMA)", k8_kernel<0>, threads, it, 4, 0);
        run("k16, k8+k8, k16, k8+k8 (per tile)", k8_kernel<1>, threads, it, 4, 0);
        run("mma + 2 FFMA2", mix_kernel<2>, threads, it, 4, 0);
        run("mma + 4 FFMA2", mix_kernel<4>, threads, it, 4, 0);
        run("mma + 8 FFMA2", mix_kernel<8>, threads, it, 4, 0);
    }
    return 0;
}
