// 1x1 convolution on tcgen05 with the reference's follow-up elementwise ops fused:
//     y[p, n] = act( sum_c a[p, c] * W[n, c] + b[n] )  [* gate[p, n]]
// Replaces the nn.Conv2d(C, N, 1) calls of MSCA.channel_mix together with the `x *` gate
// (reference src/approx/layers/msca.py:23,26) and of LowRankExpConvV1.d_conv (low_rank_conv.py:49,52) for bf16 NHWC
// tensors.  HBM-bound for C <= 256 (AI = C/3 .. C/2 FLOP/B): the kernel streams `a` once through a TMA ring, keeps W
// resident in shared memory, accumulates in TMEM and reads the gate / writes y straight from / to global memory in the
// epilogue, so the algorithmic traffic (a + gate + y) is the actual traffic.
//
//   M-tile = 128 pixels (rows of the [P, C] matrix; the last tile is zero-filled by TMA and clipped on store)
//   K-chunk = 32 channels (64-byte rows, 64B swizzle), ring of up to 8 chunks; UMMA 128 x N x 16, N <= 256
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (two accumulators of N columns in TMEM)
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <mutex>

namespace cna {
using namespace ptx;

constexpr int kC1Threads = 320;     // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9: two epilogue groups, each on its half of the N columns
constexpr int kC1Stages = 8;
constexpr int kC1KC = 32;          // channels per K-chunk

struct C1Params {
    long long P;       // pixels
    int C, N, nkc, ntiles, stages;
    long long ldy;     // row stride (elements) of y and of the gate / residual tensor (N for dense tensors)
    int nslab, slab_n; // output channels split into nslab slabs of slab_n (the last may be narrower): CTA b works on slab b % nslab, so
                       // a conv with N > 256 (fc1 of MSCAN's FFN) is ONE launch; p.N is the total then
    int act;           // 0 none, 1 ReLU, 2 GELU (erf form, nn.GELU default)
    int aux_mode;      // what `gate` is: 0 multiplicative gate (y = bf16(v) * gate), 1 residual (y = alpha[n] v + beta[n] gate + delta[n])
    uint32_t off_a, a_stage, off_w, w_chunk, off_bias, off_bar;
};

__host__ __device__ constexpr uint64_t c1_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
    return ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(layout_type & 7) << 61);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

enum : int { C1_FULL = 0, C1_EMPTY = C1_FULL + kC1Stages, C1_WFULL = C1_EMPTY + kC1Stages, C1_ACC_FULL = C1_WFULL + 1,
             C1_ACC_EMPTY = C1_ACC_FULL + 2, C1_COUNT = C1_ACC_EMPTY + 2 };

__global__ void __launch_bounds__(kC1Threads, 1)
conv1x1_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                  const float* __restrict__ bias, const __nv_bfloat16* __restrict__ gate, const float* __restrict__ alpha,
                  const float* __restrict__ beta, const float* __restrict__ delta, __nv_bfloat16* __restrict__ y, const C1Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + p.off_bar;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    float* sbias = reinterpret_cast<float*>(base_ptr + p.off_bias);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + p.off_bar + 8 * C1_COUNT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's slab of output channels: everything below sees only its N columns
    const int slab = (int)(blockIdx.x % (unsigned)p.nslab), n_off = slab * p.slab_n;
    const int N = min(p.slab_n, p.N - n_off);
    if (bias) bias += n_off;
    if (alpha) alpha += n_off;
    if (beta) beta += n_off;
    if (delta) delta += n_off;
    if (gate) gate += n_off;
    y += n_off;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a); prefetch_tmap(&tm_w);
        for (int i = 0; i < kC1Stages; ++i) { mbar_init(bar(C1_FULL + i), 1); mbar_init(bar(C1_EMPTY + i), 1); }
        mbar_init(bar(C1_WFULL), 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar(C1_ACC_FULL + i), 1); mbar_init(bar(C1_ACC_EMPTY + i), 256); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    for (int i = threadIdx.x; i < N; i += kC1Threads) {
        sbias[i] = bias ? bias[i] : 0.f;
        if (p.aux_mode == 1) {          // per-channel affine of the residual epilogue: [alpha | beta | delta] after the bias
            sbias[N + i] = alpha ? alpha[i] : 1.f;
            sbias[2 * N + i] = beta ? beta[i] : 1.f;
            sbias[3 * N + i] = delta ? delta[i] : 0.f;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const int first = blockIdx.x / p.nslab, step = gridDim.x / p.nslab;       // the launcher makes the grid a multiple of nslab
    const int my_tiles = first < p.ntiles ? (p.ntiles - first + step - 1) / step : 0;

    if (warp == 0) {
        if (my_tiles > 0) {
            if (elect_one()) {
                // the weight box is slab_n rows tall; rows past the tensor's N (narrow last slab) are zero-filled and still counted
                mbar_arrive_expect_tx(bar(C1_WFULL), (uint32_t)(p.slab_n * 64 * p.nkc));
                for (int kc = 0; kc < p.nkc; ++kc) tma_load_2d(base + p.off_w + kc * p.w_chunk, &tm_w, bar(C1_WFULL), kc * kC1KC, n_off);
            }
            __syncwarp();
            int s = 0; uint32_t ph = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int row0 = (first + it * step) * 128;
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(bar(C1_EMPTY + s), ph ^ 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar(C1_FULL + s), 128 * 64);
                        tma_load_2d(base + p.off_a + s * p.a_stage, &tm_a, bar(C1_FULL + s), kc * kC1KC, row0);
                    }
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (my_tiles > 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24);
            constexpr uint64_t dsc = c1_desc_hi(16, 8 * 64, 4);      // 64-byte rows, 64B swizzle, K-major
            int s = 0; uint32_t ph = 0;
            mbar_wait(bar(C1_WFULL), 0);
            tc_fence_after_sync();
            for (int it = 0; it < my_tiles; ++it) {
                const int buf = it & 1;
                mbar_wait(bar(C1_ACC_EMPTY + buf), (uint32_t)((it >> 1) & 1) ^ 1u);
                tc_fence_after_sync();
                const uint32_t d = tmem + buf * 256;
                for (int kc = 0; kc < p.nkc; ++kc) {
                    mbar_wait(bar(C1_FULL + s), ph);
                    tc_fence_after_sync();
                    const uint64_t a0 = dsc + ((base + p.off_a + s * p.a_stage) >> 4);
                    const uint64_t b0 = dsc + ((base + p.off_w + kc * p.w_chunk) >> 4);
                    if (elect_one()) {
                        umma_bf16(d, a0, b0, idesc, (uint32_t)(kc != 0));
                        umma_bf16(d, a0 + 2, b0 + 2, idesc, 1u);
                        umma_commit(bar(C1_EMPTY + s));
                        if (kc == p.nkc - 1) umma_commit(bar(C1_ACC_FULL + buf));
                    }
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        const int lg = warp & 3;
        const uint32_t lane_addr = (uint32_t)(lg * 32) << 16;
        const int m = lg * 32 + lane;
        const float lo = p.act == 1 ? 0.f : -INFINITY;
        // Two epilogue groups (warps 2-5, 6-9) split the N columns of every tile: with K = C <= 256 the GEMM is a few hundred cycles
        // and ONE group's ld -> math -> 32-byte stores chain per 32 columns bounded the kernel (40 us for 100 MB at C = 32).
        const int eg = warp >= 6 ? 1 : 0;
        const int half = ((N / 2 + 15) / 16) * 16;                    // columns of group 0, a multiple of 16
        const int c_lo = eg ? half : 0, c_hi = eg ? N : half;
        // gate rows are fetched one 32-column chunk AHEAD (next chunk of this tile, or the first chunk of the warp's next tile):
        // the global-load latency of a chunk would otherwise be exposed once per tile
        uint4 gq[4], gnext[4];
        auto gate_fetch = [&](int it, int n0, uint4 (&dst)[4]) {
            const long long r = (long long)(first + it * step) * 128 + m;
            if (gate && it < my_tiles && r < p.P) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (n0 + q * 8 < c_hi) {      // volatile: the load is issued HERE, not sunk down to its first use
                        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                                     : "=r"(dst[q].x), "=r"(dst[q].y), "=r"(dst[q].z), "=r"(dst[q].w)
                                     : "l"(reinterpret_cast<const uint4*>(gate + r * p.ldy + n0) + q));
                    }
            }
        };
        gate_fetch(0, c_lo, gnext);
        for (int it = 0; it < my_tiles; ++it) {
            const int buf = it & 1;
            const long long row = (long long)(first + it * step) * 128 + m;
            mbar_wait(bar(C1_ACC_FULL + buf), (uint32_t)((it >> 1) & 1));
            tc_fence_after_sync();
            const bool ok = row < p.P;
            const __nv_bfloat16* grow = gate ? gate + row * p.ldy : nullptr;
            __nv_bfloat16* yrow = y + row * p.ldy;
            if (c_lo >= c_hi) {          // N = 16: group 1 owns no columns, but the accumulator barrier counts its threads
                tc_fence_before_sync();
                mbar_arrive(bar(C1_ACC_EMPTY + buf));
            }
            for (int n0 = c_lo; n0 < c_hi; n0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int q = 0; q < 4; ++q) gq[q] = gnext[q];
                if (n0 + 32 < c_hi) gate_fetch(it, n0 + 32, gnext);
                else gate_fetch(it + 1, c_lo, gnext);
                if (n0 + 32 <= c_hi) {
                    tmem_ld32(tmem + buf * 256 + lane_addr + n0, v);
                } else {   // a 16-column tail
                    tmem_ld16(tmem + buf * 256 + lane_addr + n0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
                    for (int i = 16; i < 32; ++i) v[i] = 0;
                }
                tmem_ld_wait();
                if (n0 + 32 >= c_hi) {
                    tc_fence_before_sync();
                    mbar_arrive(bar(C1_ACC_EMPTY + buf));      // accumulator drained
                }
                if (!ok) continue;
                const int nvalid = min(32, c_hi - n0);
                uint4 pend = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q * 8 >= nvalid) break;
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = fmaxf(__uint_as_float(v[q * 8 + i]) + sbias[n0 + q * 8 + i], lo);
                    if (p.act == 2) {
#pragma unroll
                        for (int i = 0; i < 8; i += 2) { const float2 g2 = gelu_erf_fast2(f[i], f[i + 1]); f[i] = g2.x; f[i + 1] = g2.y; }
                    }
                    if (grow && p.aux_mode == 1) {
                        // residual epilogue: y = alpha[n] v + beta[n] res + delta[n] in fp32, one rounding
                        const uint32_t g4[4] = {gq[q].x, gq[q].y, gq[q].z, gq[q].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int n = n0 + q * 8 + 2 * i;
                            const float r0 = __uint_as_float(g4[i] << 16), r1 = __uint_as_float(g4[i] & 0xffff0000u);
                            f[2 * i] = sbias[N + n] * f[2 * i] + sbias[2 * N + n] * r0 + sbias[3 * N + n];
                            f[2 * i + 1] = sbias[N + n + 1] * f[2 * i + 1] + sbias[2 * N + n + 1] * r1 + sbias[3 * N + n + 1];
                        }
                    } else if (grow) {
                        // the reference multiplies two materialised bf16 tensors: round the conv output first
                        const uint32_t g4[4] = {gq[q].x, gq[q].y, gq[q].z, gq[q].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float g0 = __uint_as_float(g4[i] << 16), g1 = __uint_as_float(g4[i] & 0xffff0000u);
                            f[2 * i] = __bfloat162float(__float2bfloat16_rn(f[2 * i])) * g0;
                            f[2 * i + 1] = __bfloat162float(__float2bfloat16_rn(f[2 * i + 1])) * g1;
                        }
                    }
                    uint4 o;
                    __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
                    __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
                    o.x = *reinterpret_cast<uint32_t*>(&t0); o.y = *reinterpret_cast<uint32_t*>(&t1);
                    o.z = *reinterpret_cast<uint32_t*>(&t2); o.w = *reinterpret_cast<uint32_t*>(&t3);
                    // 16 columns = one 32-byte store (N % 16 == 0 keeps every row 32-byte aligned)
                    if (q & 1) st_global_v8(yrow + n0 + (q - 1) * 8, pend, o);
                    else pend = o;
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

// fp32 (N, C) conv weight -> bf16 [N][C] (K-major B operand: the reference's own layout, only the dtype changes)
__global__ void conv1x1_pack_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(w[i]);
}

typedef CUresult (*EncodeTiledFn1)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn1 c1_get_encode() {
    static EncodeTiledFn1 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn1>(f);
    });
    return fn;
}

static int c1_encode2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint32_t box_inner, uint32_t box_outer,
                       uint64_t row_stride_elems = 0) {
    EncodeTiledFn1 fn = c1_get_encode();
    CNA_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t str[1] = {(row_stride_elems ? row_stride_elems : inner) * 2};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CNA_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return 0;
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_conv1x1_tc_supported(int C, int N) {
    return (C >= 32 && C % 32 == 0 && C <= 1024 && N >= 16 && N % 16 == 0 && N <= 256 &&
            (size_t)N * C * 2 + 4 * 128 * 64 + 8192 <= 220 * 1024) ? 1 : 0;
}

int cna_conv1x1_pack(const float* w_nc, void* packed, int N, int C, void* stream) {
    CNA_CHECK(w_nc && packed && N > 0 && C > 0, "bad arguments");
    long long n = (long long)N * C;
    int grid = (int)std::min<long long>(ceil_div_ll(n, 256), 148LL * 8);
    conv1x1_pack_kernel<<<grid, 256, 0, as_stream(stream)>>>(w_nc, (__nv_bfloat16*)packed, n);
    CNA_LAUNCH_OK();
    return 0;
}

int cna_conv1x1_tc_forward_ex(const void* a, const void* w_packed, const float* bias, const void* aux, const float* alpha,
                              const float* beta, const float* delta, void* y, long long P, int C, int N, int act, int aux_mode,
                              void* stream);
int cna_conv1x1_tc_forward_strided(const void* a, long long lda, const void* w_packed, long long ldw, const float* bias, const void* aux,
                                   const float* alpha, const float* beta, const float* delta, void* y, long long ldy, long long P, int C,
                                   int N, int act, int aux_mode, void* stream);

int cna_conv1x1_tc_forward(const void* a, const void* w_packed, const float* bias, const void* gate, void* y, long long P, int C,
                           int N, int relu, void* stream) {
    return cna_conv1x1_tc_forward_ex(a, w_packed, bias, gate, nullptr, nullptr, nullptr, y, P, C, N, relu ? 1 : 0, 0, stream);
}

int cna_conv1x1_tc_forward_ex(const void* a, const void* w_packed, const float* bias, const void* aux, const float* alpha,
                              const float* beta, const float* delta, void* y, long long P, int C, int N, int act, int aux_mode,
                              void* stream) {
    return cna_conv1x1_tc_forward_strided(a, C, w_packed, C, bias, aux, alpha, beta, delta, y, N, P, C, N, act, aux_mode, stream);
}

int cna_conv1x1_tc_forward_strided(const void* a, long long lda, const void* w_packed, long long ldw, const float* bias, const void* aux,
                                   const float* alpha, const float* beta, const float* delta, void* y, long long ldy, long long P, int C,
                                   int N, int act, int aux_mode, void* stream) {
    const void* gate = aux;
    CNA_CHECK(lda >= C && ldw >= C && ldy >= N && lda % 8 == 0 && ldw % 8 == 0 && ldy % 16 == 0, "bad leading dimensions (lda, ldw multiples of 8, ldy of 16)");
    CNA_CHECK(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(w_packed)) & 15u) == 0 &&
              ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(aux)) & 31u) == 0, "operands must be 16-byte (a, w) / 32-byte (y, aux) aligned");
    CNA_CHECK(a && w_packed && y && P > 0, "bad arguments");
    CNA_CHECK(act >= 0 && act <= 2 && (aux_mode == 0 || aux_mode == 1), "bad epilogue selector");
    CNA_CHECK(aux_mode == 0 || aux != nullptr, "residual epilogue without a residual tensor");
    // N > 256: slabs of 256 output channels (two accumulators of 256 TMEM columns), one CTA group per slab in the same launch
    const int slab_n = N > 256 ? 256 : N, nslab = (N + slab_n - 1) / slab_n;
    CNA_CHECK(N % 16 == 0 && cna_conv1x1_tc_supported(C, slab_n),
              "configuration not supported (needs C%32==0, N%16==0, a 256-channel slab of W resident in smem)");
    CNA_CHECK(P < (1LL << 31) * 128LL, "too many pixels");
    DevInfo di;
    if (dev_info(&di)) return 1;
    const int num_sms = di.num_sms, max_smem = di.max_smem;
    static std::atomic<unsigned long long> attr_done{0};
    CNA_CUDA(ensure_dyn_smem(conv1x1_tc_kernel, max_smem, attr_done));
    C1Params p{};
    p.P = P; p.C = C; p.N = N; p.nkc = C / kC1KC; p.act = act; p.aux_mode = aux_mode; p.ldy = ldy; p.nslab = nslab; p.slab_n = slab_n;
    p.ntiles = (int)((P + 127) / 128);
    p.a_stage = 128 * 64;
    p.w_chunk = ((uint32_t)slab_n * 64 + 1023u) & ~1023u;
    const uint32_t wbytes = p.w_chunk * p.nkc;
    const uint32_t bias_bytes = ((uint32_t)slab_n * 16 + 1023u) & ~1023u;      // bias | alpha | beta | delta
    int stages = kC1Stages;
    while (stages > 2 && (size_t)wbytes + (size_t)stages * p.a_stage + bias_bytes + 2048 > (size_t)max_smem) --stages;
    CNA_CHECK((size_t)wbytes + (size_t)stages * p.a_stage + bias_bytes + 2048 <= (size_t)max_smem, "shared-memory plan does not fit");
    p.stages = stages;
    p.off_a = 0; p.off_w = stages * p.a_stage; p.off_bias = p.off_w + wbytes; p.off_bar = p.off_bias + bias_bytes;
    const size_t smem = (size_t)p.off_bar + 2048;
    CUtensorMap tm_a, tm_w;
    if (c1_encode2d(&tm_a, a, (uint64_t)C, (uint64_t)P, kC1KC, 128, (uint64_t)lda)) return 1;
    if (c1_encode2d(&tm_w, w_packed, (uint64_t)C, (uint64_t)N, kC1KC, (uint32_t)slab_n, (uint64_t)ldw)) return 1;
    const int grid = std::max(1, std::min(p.ntiles, num_sms / nslab)) * nslab;
    conv1x1_tc_kernel<<<grid, kC1Threads, smem, as_stream(stream)>>>(tm_a, tm_w, bias, (const __nv_bfloat16*)gate, alpha, beta, delta,
                                                                     (__nv_bfloat16*)y, p);
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
