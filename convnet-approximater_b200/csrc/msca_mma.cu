// Fused MSCA "sd" chain on the warp-level tensor path:   s = [conv0(x)] + sum_k cascade_k(conv0(x))   in ONE pass over HBM, bf16.
//
// Contract of the reference modules it replaces (src/approx/layers/msca.py:18-26, depth_seperable_conv.py:12-16, 45-47, 57-61;
// MscaRep's rank-r forms core/msca_rep.py:44-55): read x once, write s once, every tensor the reference's bf16 modules
// materialise (conv0 out, each 1xk out, each branch out, every partial sum) rounded to bf16 at the same point.
//
// Why tensor cores for a DEPTHWISE chain: the FFMA2 kernel (msca16.cu) is issue-bound at 103 MAC per element (ncu: 51 % issue,
// 41 % FMA pipe).  Per channel, a 1xk strip is a product with a banded Toeplitz matrix, and mma.sync.m16n8k16 (bf16, fp32
// accumulate) issues at 8.2 cycles per sub-partition on this part (benchmarks/ubench_hmma.cu) = 256 MAC/cycle against the 64 of
// FFMA2 — even with the band's zeros (a 16-wide window carries 7..14 useful taps) that is 4-6x fewer issue slots per output:
//     conv0 5x5   : per tap row dy, out[16 rows x 8 cols] += x[16 rows(+dy) x 16 cols] . Toep_dy[16 x 8]        5 MMAs / 128 outputs
//     1xk (hpass) : out[16 rows x 8 cols]  = a0[16 rows x 16|32 cols] . Toep_k                                   1 | 2 | 2 MMAs (k = 7 | 11 | 21)
//     kx1 (vpass) : out^T[16 cols x 8 rows] = T_k^T[16 cols x 16|32 rows] . Toep_k   (A through ldmatrix.trans)  1 | 2 | 2 MMAs
// = 75 MMAs per channel and 16x32-pixel step against ~950 FFMA2.  fp32 accumulators start at the bias; rounding points are the
// reference's.  Tap weights are rounded to bf16 — exact for the reference's bf16 modules (their parameters are bf16).
//
// Organisation: a CTA owns (image, 16 channels, 32 output columns) and marches down the image 16 rows per step; a WARP OWNS TWO
// CHANNELS: their Toeplitz B-fragments (60 registers) are built once and stay in registers for the whole march, and
// conv0 -> hpass -> vpass of a channel need no CTA barrier.  Shared memory is channel-planar bf16 (ldmatrix wants 16-byte runs of
// the contracted index), all planes laid out so that every ldmatrix / fragment store is bank-conflict free:
//     X   [16 ch][20 rows][64 cols]   144-byte rows (9 chunks: odd multiple of 16 bytes, every ldmatrix address = base + immediate)
//     A0  [16 ch][32-row ring][56]    112-byte rows (7 chunks: odd multiple of 16 bytes)
//     T_k [16 ch][ring_k][32 cols]    64-byte rows, chunk XOR ((slot >> 1) & 3); ring 32 rows (k = 7, 11) or 40 (k = 21)
//     O   [16 ch][32 cols][16 rows]   the step's output block, column-major pairs (the vpass fragment layout), XOR-swizzled
// The NHWC <-> planar transposes are done in registers (PRMT): x rows of the NEXT step are fetched while the current one computes;
// the output block is written out 32 bytes per pixel.
//
// Two kernels share the arithmetic (sdm_launch_t):
//   msca_sdw_kernel  warp-specialised (shipped): 8 compute warps (216 | 208 registers after setmaxnreg) + 4 | 8 IO warps (64 | 48)
//                    that do the transposes, four mbarriers, no CTA-wide barrier in the loop.  8 IO warps for a single cascade.
//   msca_sdm_kernel  8 warps that all do everything, two __syncthreads per step (kept behind CNA_MSCA_WS=0 as the A/B reference).
// Measured on B200 (13 MSCAN-T blocks at 512 x 512, batch 32, bf16, one CUDA graph; profiles/r02_msca_mma.txt): original block
// 1.75 ms (FFMA2 kernel, msca16.cu) -> 0.97 ms, MscaRep d1+fix 1.34 -> 0.90 ms.  ncu of the (32 ch, 128 x 128) launch: 96 us against 197;
// HMMA pipe 25 % active, issue 34 %: the kernel is bound by per-warp dependent latency (ldmatrix -> HMMA -> cvt -> st.shared chains,
// 7.8 cycles per instruction and warp with 2 compute warps per scheduler) and by the 255-register ceiling that keeps it at 8 compute
// warps, not by a pipe.  Rejected by measurement: 16 warps x 1 channel at 128 registers (spills; 1.28 ms), own ldmatrix loads for the
// odd-aligned A quads instead of register moves (fewer instructions, more shared-memory wavefronts: 1.03 -> 1.10 ms), m16n8k8 for
// the half windows (same 8.2-cycle issue as k16).
//
// Column bookkeeping (w0c = first output column of the CTA):  x position q <-> image column w0c - 13 + q;  a0 position q <-> column
// w0c - 11 + q (so conv0's window of a0 chunk j starts at x chunk j, and the 1x7 / 1x21 windows start on chunk boundaries).
#include "common.cuh"
#include "ptx.cuh"

namespace cna {
namespace {

constexpr int M_THREADS = 512;
constexpr int M_XROWS = 20, M_XPITCH = 144, M_XPLANE = M_XROWS * M_XPITCH;      // bytes (pitch = odd multiple of 16: conflict-free ldmatrix)
constexpr int M_AROWS = 32, M_APITCH = 112, M_APLANE = M_AROWS * M_APITCH;
constexpr int M_TPITCH = 64;
constexpr int M_OPLANE = 32 * 32;                                               // 32 columns x 16 rows x 2 bytes
constexpr int M_P = 10;                                                         // output lag (rows) = halo of the widest strip

template <int K> struct Br;      // ring rows | first a0 chunk of tile i's window (relative to i) | tap offset inside it | MMAs per tile
template <> struct Br<7>  { static constexpr int RING = 32, HC0 = 1, HOFF = 0, NM = 1; };
template <> struct Br<11> { static constexpr int RING = 32, HC0 = 0, HOFF = 6, NM = 2; };
template <> struct Br<21> { static constexpr int RING = 40, HC0 = 0, HOFF = 1, NM = 2; };
template <> struct Br<0>  { static constexpr int RING = 0; };

template <int K0, int K1, int K2> struct Lay {
    static constexpr int NB = (K0 > 0) + (K1 > 0) + (K2 > 0);
    static constexpr int X_OFF = 0;
    static constexpr int A_OFF = 16 * M_XPLANE;
    static constexpr int T0_OFF = A_OFF + 16 * M_APLANE;
    static constexpr int T1_OFF = T0_OFF + 16 * Br<K0>::RING * M_TPITCH;
    static constexpr int T2_OFF = T1_OFF + 16 * Br<K1>::RING * M_TPITCH;
    static constexpr int O_OFF = T2_OFF + 16 * Br<K2>::RING * M_TPITCH;
    static constexpr int R_OFF = O_OFF + 16 * M_OPLANE;              // border residues of the CTA's channels: [top | bottom][16 rows][16 ch] fp32
    static constexpr int BAR_OFF = R_OFF + 2 * 16 * 16 * 4;          // four mbarriers of the warp-specialised kernel
    static constexpr int TOTAL = BAR_OFF + 64;
};

struct SdmParams {
    int B, H, W, C;
    int identity, res_p, strips, steps;
    int lag;          // rows the output block trails the conv0 block: 10 (the widest strip's halo), or 0 when the whole image is one step
};

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t (&r)[2], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
// D += A . B, m16n8k16, bf16 operands, fp32 accumulators
// (the B fragment travels as one 64-bit value so that the allocator keeps its two registers adjacent, as HMMA wants them)
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, unsigned long long b) {
    asm("{\n\t.reg .b32 b0, b1;\n\tmov.b64 {b0, b1}, %8;\n\t"
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {b0,b1}, {%0,%1,%2,%3};\n\t}"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "l"(b));
}
// D = A . B + C with C kept (the bias quad of a conv is built once and never overwritten)
__device__ __forceinline__ void mma16816c(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, unsigned long long b,
                                          const float (&c)[4]) {
    asm("{\n\t.reg .b32 b0, b1;\n\tmov.b64 {b0, b1}, %8;\n\t"
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {b0,b1}, {%9,%10,%11,%12};\n\t}"
        : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "l"(b), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {      // {bf16(lo) in bits 0-15, bf16(hi) in bits 16-31}, round to nearest even
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {   // both halves: bf16(a + b)
    __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// slot of image row (16 s + off) in a ring of RING rows; base = (16 s) mod RING, off in [-RING, 2 RING)
template <int RING> __device__ __forceinline__ int ring_slot(int base, int off) {
    if constexpr (RING == 32) {
        return (base + off) & 31;
    } else {
        int v = base + off + RING;                     // [0, 4 RING)
        v -= v >= 2 * RING ? 2 * RING : 0;
        v -= v >= RING ? RING : 0;
        return v;
    }
}

// Toeplitz B fragment of one MMA: B[kk][n] = f[kk - n - off] (0 outside the K taps), kk = 16 q + {2t, 2t+1 | 2t+8, 2t+9}, n = g
template <int K>
__device__ __forceinline__ void toep_frag(unsigned long long& b, const float* f /* [tap][16 ch] */, int ch, int q, int off, int g, int t) {
    auto tp = [&](int idx) { return (idx >= 0 && idx < K) ? f[idx * 16 + ch] : 0.f; };
    const int i0 = 16 * q + 2 * t - g - off;
    b = (unsigned long long)pack2(tp(i0), tp(i0 + 1)) | ((unsigned long long)pack2(tp(i0 + 8), tp(i0 + 9)) << 32);
}

// Horizontal 1xK pass of one branch: T_k rows [16 s, 16 s + 16), 32 columns  <-  the a0 block in registers.  ae[q] / ao[q] are the
// A operands (16 rows x 16 columns) starting at a0 chunk 2q / 2q + 1 — every window is a register quad as ldmatrix delivered it.
// st0 / st1: shared-memory addresses of T chunk 0 of rows g / g + 8 with the row's swizzle already applied (chunk i = address ^ 16 i).
template <int K>
__device__ __forceinline__ void sdm_hpass(uint32_t st0, uint32_t st1, const uint32_t (&ae)[3][4], const uint32_t (&ao)[3][4],
                                          const unsigned long long (&bh)[2], float bias, bool rin0, bool rin1) {
    using C = Br<K>;
    const float cq[4] = {bias, bias, bias, bias};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float c[4];
#pragma unroll
        for (int q = 0; q < C::NM; ++q) {
            const int cb = i + C::HC0 + 2 * q;                   // first chunk of the 16-column window (compile time)
            const uint32_t(&a)[4] = (cb & 1) ? ao[cb >> 1] : ae[cb >> 1];
            if (q == 0) mma16816c(c, a[0], a[1], a[2], a[3], bh[q], cq);
            else mma16816(c, a[0], a[1], a[2], a[3], bh[q]);
        }
        sts32(st0 ^ (i << 4), rin0 ? pack2(c[0], c[1]) : 0u);             // rows outside the image: the vertical conv sees zero padding
        sts32(st1 ^ (i << 4), rin1 ? pack2(c[2], c[3]) : 0u);
    }
}

// Vertical Kx1 pass of one branch for 16 output columns and the step's 16 output rows [16 s - 10, 16 s + 6): adds the rounded
// branch output (bias included) to tot[n'][col half] (n' = 8-row half; words = rows 2t, 2t+1 of column g | g + 8).
// va[i]: ldmatrix addresses of the 8-row block pairs (2i, 2i+1) of the window (last one: the single block 2 NM), this channel and
// column half already applied.  Half n' = 1 starts on an odd block: its A quads are assembled with register moves — measured
// against loading the odd pairs with their own ldmatrix (no moves, 5 more x4 loads per 16 columns): the loads lose (13 blocks
// 1.03 -> 1.10 ms): shared-memory wavefronts are scarcer than issue slots here.
template <int K>
__device__ __forceinline__ void sdm_vbranch(const uint32_t (&va)[3], const unsigned long long (&bv)[2], float bias, bool first, uint32_t (&tot)[2][2]) {
    using C = Br<K>;
    constexpr int NM = C::NM;
    uint32_t qd[NM + 1][4];                                  // qd[i] = {block 2i: cols g | g+8, block 2i+1: cols g | g+8}
#pragma unroll
    for (int i = 0; i < NM; ++i) ldsm4t(qd[i], va[i]);
    {
        uint32_t r[2];
        ldsm2t(r, va[NM]);
        qd[NM][0] = r[0]; qd[NM][1] = r[1]; qd[NM][2] = 0u; qd[NM][3] = 0u;
    }
    const float cq[4] = {bias, bias, bias, bias};
#pragma unroll
    for (int np = 0; np < 2; ++np) {
        float c[4];
#pragma unroll
        for (int q = 0; q < NM; ++q) {
            // window of half n' = blocks (n' + 2q, n' + 2q + 1)
            const uint32_t a0 = np ? qd[q][2] : qd[q][0], a1 = np ? qd[q][3] : qd[q][1];
            const uint32_t a2 = np ? qd[q + 1][0] : qd[q][2], a3 = np ? qd[q + 1][1] : qd[q][3];
            if (q == 0) mma16816c(c, a0, a1, a2, a3, bv[q], cq);
            else mma16816(c, a0, a1, a2, a3, bv[q]);
        }
        const uint32_t v0 = pack2(c[0], c[1]), v1 = pack2(c[2], c[3]);
        tot[np][0] = first ? v0 : hadd2(tot[np][0], v0);
        tot[np][1] = first ? v1 : hadd2(tot[np][1], v1);
    }
}

// per-step ldmatrix / store offsets of one branch's T ring (shared by the warp's channels: a channel adds its plane offset)
template <int K>
__device__ __forceinline__ void sdm_ring_addrs(int base, int lag, int lane, int g, int t, uint32_t& st0, uint32_t& st1, uint32_t (&va)[3]) {
    using C = Br<K>;
    constexpr int P = K / 2;
    const int s0 = ring_slot<C::RING>(base, g), s1 = ring_slot<C::RING>(base, g + 8);
    st0 = s0 * M_TPITCH + t * 4 + (((s0 >> 1) & 3) << 4);
    st1 = s1 * M_TPITCH + t * 4 + (((s1 >> 1) & 3) << 4);
    const int cbit = (lane >> 3) & 1;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i <= C::NM) {
            const int blk = 2 * i + (i < C::NM ? (lane >> 4) : 0);
            const int slot = ring_slot<C::RING>(base, -lag - P + 8 * blk + (lane & 7));
            va[i] = slot * M_TPITCH + ((cbit ^ ((slot >> 1) & 3)) << 4);
        } else {
            va[i] = 0;
        }
    }
}

// CPW = channels per warp: 1 -> 16 warps (128 registers per thread), 2 -> 8 warps (255 registers).
template <int K0, int K1, int K2, int CPW, bool IDENT>
__global__ void __launch_bounds__(M_THREADS / CPW, 1)
msca_sdm_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ w0, const float* __restrict__ b0,
                const float* __restrict__ w1cat, const float* __restrict__ b1cat, const float* __restrict__ w2cat,
                const float* __restrict__ b2cat, const float* __restrict__ res, const SdmParams p) {
    using L = Lay<K0, K1, K2>;
    constexpr int NB = L::NB, NT = M_THREADS / CPW, NW = 16 / CPW;
    constexpr int XR = (2 * M_XROWS + NW - 1) / NW;                // x staging rounds: warp task = (row, channel half), 40 tasks
    constexpr int XRA = (XR + 1) / 2;                               // rounds of the first batch (fetched at B1, stored after hpass)
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sb = ptx::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int strip = blockIdx.x % p.strips, cg = blockIdx.x / p.strips, b = blockIdx.y;
    const int w0c = strip * 32, c0 = cg * 16;
    const int H = p.H, W = p.W, C = p.C;
    const __nv_bfloat16* xb = x + (size_t)b * H * W * C;
    __nv_bfloat16* yb = y + (size_t)b * H * W * C;

    // ---- x staging: warp task wt = warp + 8 r = (row, channel half hf), lane = column pair: pixels 2 lane, 2 lane + 1 ----
    const int xcol = w0c - 13 + 2 * lane;
    const bool xc0 = xcol >= 0 && xcol < W, xc1 = xcol + 1 >= 0 && xcol + 1 < W;
    uint4 xr[XR][2];
    auto x_fetch = [&](int s, int ra, int rb) {
        const int r0 = 16 * s - 2;
#pragma unroll
        for (int r = 0; r < XR; ++r) {
            if (r < ra || r >= rb) continue;
            const int wt = warp + NW * r, row = r0 + (wt >> 1), hf = wt & 1;
            const bool rin = row >= 0 && row < H && wt < 2 * M_XROWS;
            const __nv_bfloat16* src = xb + ((ptrdiff_t)row * W + xcol) * C + c0 + 8 * hf;
            xr[r][0] = (rin && xc0) ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0);
            xr[r][1] = (rin && xc1) ? __ldg(reinterpret_cast<const uint4*>(src + C)) : make_uint4(0, 0, 0, 0);
        }
    };
    auto x_store = [&](int ra, int rb) {
#pragma unroll
        for (int r = 0; r < XR; ++r) {
            if (r < ra || r >= rb) continue;
            const int wt = warp + NW * r;
            if (wt >= 2 * M_XROWS) continue;
            const uint32_t a = sb + L::X_OFF + 8 * (wt & 1) * M_XPLANE + (wt >> 1) * M_XPITCH + lane * 4;
            const uint32_t u0[4] = {xr[r][0].x, xr[r][0].y, xr[r][0].z, xr[r][0].w}, u1[4] = {xr[r][1].x, xr[r][1].y, xr[r][1].z, xr[r][1].w};
#pragma unroll
            for (int c = 0; c < 8; ++c) sts32(a + c * M_XPLANE, prmt(u0[c >> 1], u1[c >> 1], (c & 1) ? 0x7632u : 0x5410u));
        }
    };
    x_fetch(0, 0, XR);

    // ---- zero the a0 and T rings (rows above the image read as zero padding; nothing an MMA can touch before it is written is
    //      ever NaN); X and O are always written before they are read ----
    for (int i = tid; i < (L::O_OFF - L::A_OFF) / 16; i += NT) reinterpret_cast<uint4*>(smem + L::A_OFF)[i] = make_uint4(0, 0, 0, 0);
    if (res != nullptr && p.res_p > 0) {
        float* rs = reinterpret_cast<float*>(smem + L::R_OFF);
        for (int i = tid; i < 2 * 16 * 16; i += NT) {
            const int half = i >> 8, rr = (i >> 4) & 15, c = i & 15;
            rs[i] = rr < p.res_p ? res[((size_t)half * C + c0 + c) * p.res_p + rr] : 0.f;
        }
    }
    __syncthreads();

    // ---- weights of the CTA's 16 channels -> shared memory (in the O region), then each warp builds its Toeplitz fragments ----
    constexpr int KS[3] = {K0, K1, K2};
    constexpr int KOFF[3] = {0, K0, K0 + K1};
    constexpr int WO_B0 = 25 * 16, WO_BR0 = WO_B0 + 16;
    constexpr int WO_BR[3] = {WO_BR0, WO_BR0 + (2 * K0 + 2) * 16, WO_BR0 + (2 * K0 + 2) * 16 + (2 * K1 + 2) * 16};   // [w1 k][b1][w2 k][b2]
    float* wst = reinterpret_cast<float*>(smem + L::O_OFF);
    for (int e = tid; e < 25 * 16; e += NT) wst[e] = w0[(size_t)(e >> 4) * C + c0 + (e & 15)];
    if (tid < 16) wst[WO_B0 + tid] = b0 ? b0[c0 + tid] : 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int k = KS[i];
        float* d = wst + WO_BR[i];
        for (int e = tid; e < k * 16; e += NT) {
            d[e] = w1cat[(size_t)(KOFF[i] + (e >> 4)) * C + c0 + (e & 15)];
            d[(k + 1) * 16 + e] = w2cat[(size_t)(KOFF[i] + (e >> 4)) * C + c0 + (e & 15)];
        }
        if (tid < 16) {
            d[k * 16 + tid] = b1cat ? b1cat[(size_t)i * C + c0 + tid] : 0.f;
            d[(2 * k + 1) * 16 + tid] = b2cat ? b2cat[(size_t)i * C + c0 + tid] : 0.f;
        }
    }
    __syncthreads();
    unsigned long long bw0[CPW][5], bh[CPW][3][2], bv[CPW][3][2];
    float bias0[CPW], bias1[CPW][3], bias2[CPW][3];
#pragma unroll
    for (int cw = 0; cw < CPW; ++cw) {
        const int ch = warp * CPW + cw;
        bias0[cw] = wst[WO_B0 + ch];
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) toep_frag<5>(bw0[cw][dy], wst + dy * 5 * 16, ch, 0, 0, g, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int q = 0; q < 2; ++q) { bh[cw][i][q] = 0ull; bv[cw][i][q] = 0ull; }
            bias1[cw][i] = bias2[cw][i] = 0.f;
        }
        if constexpr (K0 > 0) {
            const float* d = wst + WO_BR[0];
#pragma unroll
            for (int q = 0; q < Br<K0>::NM; ++q) { toep_frag<K0>(bh[cw][0][q], d, ch, q, Br<K0>::HOFF, g, t); toep_frag<K0>(bv[cw][0][q], d + (K0 + 1) * 16, ch, q, 0, g, t); }
            bias1[cw][0] = d[K0 * 16 + ch]; bias2[cw][0] = d[(2 * K0 + 1) * 16 + ch];
        }
        if constexpr (K1 > 0) {
            const float* d = wst + WO_BR[1];
#pragma unroll
            for (int q = 0; q < Br<K1>::NM; ++q) { toep_frag<K1>(bh[cw][1][q], d, ch, q, Br<K1>::HOFF, g, t); toep_frag<K1>(bv[cw][1][q], d + (K1 + 1) * 16, ch, q, 0, g, t); }
            bias1[cw][1] = d[K1 * 16 + ch]; bias2[cw][1] = d[(2 * K1 + 1) * 16 + ch];
        }
        if constexpr (K2 > 0) {
            const float* d = wst + WO_BR[2];
#pragma unroll
            for (int q = 0; q < Br<K2>::NM; ++q) { toep_frag<K2>(bh[cw][2][q], d, ch, q, Br<K2>::HOFF, g, t); toep_frag<K2>(bv[cw][2][q], d + (K2 + 1) * 16, ch, q, 0, g, t); }
            bias1[cw][2] = d[K2 * 16 + ch]; bias2[cw][2] = d[(2 * K2 + 1) * 16 + ch];
        }
    }
    x_store(0, XR);
    __syncthreads();                                    // weights consumed (O is free), x rows of step 0 staged

    // ---- per-lane address constants (first channel of the warp; channel cw adds cw * plane size as an immediate) ----
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;          // ldmatrix.x4: row / chunk half supplied by this lane
    const uint32_t x_ld = sb + L::X_OFF + warp * CPW * M_XPLANE + lrow * M_XPITCH + lhi * 16;     // + dy * pitch + 32 q
    const uint32_t a_ld = sb + L::A_OFF + warp * CPW * M_APLANE + lrow * M_APITCH + lhi * 16;     // + block * 16 pitch + 32 q (+16)
    const uint32_t a_st = sb + L::A_OFF + warp * CPW * M_APLANE + g * M_APITCH + t * 4;           // + block * 16 pitch + 16 j (+ 8 pitch)
    const uint32_t a_id = sb + L::A_OFF + warp * CPW * M_APLANE + (g + 11) * 2;                   // + slot * pitch + 2 (16 mp + 8 hc)
    const int osw = 4 * ((g >> 2) & 1);
    const uint32_t o_st = sb + L::O_OFF + warp * CPW * M_OPLANE + g * 32;                         // + 32 (16 mp + 8 hc) + 4 ((4 np + t) ^ osw)
    const uint32_t t_w0 = sb + L::T0_OFF + warp * CPW * Br<K0>::RING * M_TPITCH, t_w1 = sb + L::T1_OFF + warp * CPW * Br<K1>::RING * M_TPITCH,
                   t_w2 = sb + L::T2_OFF + warp * CPW * Br<K2>::RING * M_TPITCH;
    const int p2 = (res != nullptr && p.res_p > 0) ? min(H, p.res_p) : 0;
    int base40 = 0;                                     // (16 s) mod 40
    const bool cols_in = w0c - 11 >= 0 && w0c + 45 <= W;        // every a0 position of the CTA is an image column
    uint32_t cmask = 0;                                 // bit 2j + e: a0 position 8j + 2t + e is an image column
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        const int c = w0c - 11 + 8 * (i >> 1) + 2 * t + (i & 1);
        if (c >= 0 && c < W) cmask |= 1u << i;
    }

    for (int s = 0; s < p.steps; ++s) {
        const int base32 = (s & 1) * 16;
        // ================= conv0: x rows [16s-2, 16s+18) -> a0 rows [16s, 16s+16), positions 0..55 =================
        if (16 * s < H) {
            const bool rows_in = 16 * s + 16 <= H;
            const bool r0in = 16 * s + g < H, r1in = 16 * s + g + 8 < H;
#pragma unroll
            for (int cw = 0; cw < CPW; ++cw) {
                const float cq[4] = {bias0[cw], bias0[cw], bias0[cw], bias0[cw]};
                float acc[7][4];
#pragma unroll
                for (int dy = 0; dy < 5; ++dy) {
                    // odd tiles start on an odd chunk: their A quads are assembled with register moves (cheaper than 3 more ldmatrix)
                    uint32_t q4[4][4];                  // q4[q] = {chunk 2q rows g | g+8, chunk 2q+1 rows g | g+8}
#pragma unroll
                    for (int q = 0; q < 4; ++q) ldsm4(q4[q], x_ld + cw * M_XPLANE + dy * M_XPITCH + q * 32);
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const int q = j >> 1;
                        const uint32_t a0 = (j & 1) ? q4[q][2] : q4[q][0], a1 = (j & 1) ? q4[q][3] : q4[q][1];
                        const uint32_t a2 = (j & 1) ? q4[q + 1][0] : q4[q][2], a3 = (j & 1) ? q4[q + 1][1] : q4[q][3];
                        if (dy == 0) mma16816c(acc[j], a0, a1, a2, a3, bw0[cw][dy], cq);
                        else mma16816(acc[j], a0, a1, a2, a3, bw0[cw][dy]);
                    }
                }
                // outside the image (rows or columns) the strip convs see ZERO padding, not conv0 of a padded input
                const uint32_t d0 = a_st + cw * M_APLANE + base32 * M_APITCH, d1 = d0 + 8 * M_APITCH;
                if (rows_in && cols_in) {
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        sts32(d0 + j * 16, pack2(acc[j][0], acc[j][1]));
                        sts32(d1 + j * 16, pack2(acc[j][2], acc[j][3]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const uint32_t m = (((cmask >> (2 * j)) & 1u) * 0xffffu) | (((cmask >> (2 * j + 1)) & 1u) * 0xffff0000u);
                        sts32(d0 + j * 16, r0in ? (pack2(acc[j][0], acc[j][1]) & m) : 0u);
                        sts32(d1 + j * 16, r1in ? (pack2(acc[j][2], acc[j][3]) & m) : 0u);
                    }
                }
            }
        }
        __syncthreads();                                 // B1: X is free, O is free (last step's write-out is done)
        const bool more = s + 1 < p.steps && 16 * (s + 1) - 2 < H;
        if (more) x_fetch(s + 1, 0, XRA);

        // per-step ring addresses (shared by the warp's channels)
        uint32_t st0[3] = {0, 0, 0}, st1[3] = {0, 0, 0}, va[3][3];
        if constexpr (K0 > 0) sdm_ring_addrs<K0>(Br<K0>::RING == 32 ? base32 : base40, p.lag, lane, g, t, st0[0], st1[0], va[0]);
        if constexpr (K1 > 0) sdm_ring_addrs<K1>(Br<K1>::RING == 32 ? base32 : base40, p.lag, lane, g, t, st0[1], st1[1], va[1]);
        if constexpr (K2 > 0) sdm_ring_addrs<K2>(Br<K2>::RING == 32 ? base32 : base40, p.lag, lane, g, t, st0[2], st1[2], va[2]);
        const bool rin0 = 16 * s + g < H, rin1 = 16 * s + g + 8 < H;

        // ================= hpass: a0 block -> T_k rows [16s, 16s+16) of every branch =================
#pragma unroll
        for (int cw = 0; cw < CPW; ++cw) {
            uint32_t ae[3][4], ao[3][4];
            const uint32_t ra = a_ld + cw * M_APLANE + base32 * M_APITCH;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                ldsm4(ae[q], ra + q * 32);
                ldsm4(ao[q], ra + q * 32 + 16);
            }
            if constexpr (K0 > 0) sdm_hpass<K0>(t_w0 + cw * Br<K0>::RING * M_TPITCH + st0[0], t_w0 + cw * Br<K0>::RING * M_TPITCH + st1[0], ae, ao, bh[cw][0], bias1[cw][0], rin0, rin1);
            if constexpr (K1 > 0) sdm_hpass<K1>(t_w1 + cw * Br<K1>::RING * M_TPITCH + st0[1], t_w1 + cw * Br<K1>::RING * M_TPITCH + st1[1], ae, ao, bh[cw][1], bias1[cw][1], rin0, rin1);
            if constexpr (K2 > 0) sdm_hpass<K2>(t_w2 + cw * Br<K2>::RING * M_TPITCH + st0[2], t_w2 + cw * Br<K2>::RING * M_TPITCH + st1[2], ae, ao, bh[cw][2], bias1[cw][2], rin0, rin1);
        }
        __syncwarp();

        // ================= vpass: T_k rings -> output rows [16s-10, 16s+6) into O (+ identity, + border residues) =================
        {
            uint32_t idoff[2][2];
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                idoff[np][0] = ((base32 - p.lag + 8 * np + 2 * t) & 31) * M_APITCH;
                idoff[np][1] = ((base32 - p.lag + 8 * np + 2 * t + 1) & 31) * M_APITCH;
            }
#pragma unroll
            for (int cw = 0; cw < CPW; ++cw) {
#pragma unroll
                for (int mp = 0; mp < 2; ++mp) {
                    if (more && cw == CPW - 1 && mp == (CPW == 1 ? 1 : 0)) { x_store(0, XRA); x_fetch(s + 1, XRA, XR); }
                    uint32_t tot[2][2];
                    uint32_t v3[3];
                    if constexpr (K0 > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v3[i] = (t_w0 + cw * Br<K0>::RING * M_TPITCH + va[0][i]) ^ (mp << 5);
                        sdm_vbranch<K0>(v3, bv[cw][0], bias2[cw][0], true, tot);
                    }
                    if constexpr (K1 > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v3[i] = (t_w1 + cw * Br<K1>::RING * M_TPITCH + va[1][i]) ^ (mp << 5);
                        sdm_vbranch<K1>(v3, bv[cw][1], bias2[cw][1], false, tot);
                    }
                    if constexpr (K2 > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v3[i] = (t_w2 + cw * Br<K2>::RING * M_TPITCH + va[2][i]) ^ (mp << 5);
                        sdm_vbranch<K2>(v3, bv[cw][2], bias2[cw][2], false, tot);
                    }
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            uint32_t v = tot[np][hc];
                            if constexpr (IDENT) {
                                const uint32_t ia = a_id + cw * M_APLANE + (16 * mp + 8 * hc) * 2;
                                v = hadd2(v, lds16(ia + idoff[np][0]) | (lds16(ia + idoff[np][1]) << 16));
                            }
                            sts32(o_st + cw * M_OPLANE + (16 * mp + 8 * hc) * 32 + ((((4 * np) ^ osw) + t) << 2), v);
                        }
                    }
                }
            }
        }
        if (more) x_store(XRA, XR);
        __syncthreads();                                 // B2: O complete, X holds the rows of step s + 1
        // ================= write-out: O -> y, 32 bytes (16 channels) per pixel, both rows of a word =================
        {
            const int col = 4 * (warp / (NW / 8)) + (lane >> 3), rp = lane & 7, icol = w0c + col;
            const int row_lo = 16 * s - p.lag + 2 * rp;
            if (icol < W && row_lo + 1 >= 0 && row_lo < H) {
                const uint32_t oa = sb + L::O_OFF + col * 32 + ((rp ^ (((col >> 2) & 1) * 4)) << 2);
                uint32_t w[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) w[c] = lds32(oa + c * M_OPLANE);
                if (p2 > 0 && (row_lo < p2 || row_lo + 1 >= H - p2)) {
                    // FixPaddingBias (MscaRep's border residues): top rows then bottom rows, each add materialised
                    const float* rs_top = reinterpret_cast<const float*>(smem + L::R_OFF);
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        float f[2] = {__uint_as_float(w[c] << 16), __uint_as_float(w[c] & 0xffff0000u)};
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int rr = row_lo + e;
                            if (rr >= 0 && rr < H) {
                                if (rr < p2) f[e] = __bfloat162float(__float2bfloat16_rn(f[e] + rs_top[rr * 16 + c]));
                                if (rr >= H - p2) f[e] += rs_top[256 + (p.res_p - (H - rr)) * 16 + c];
                            }
                        }
                        w[c] = pack2(f[0], f[1]);
                    }
                }
                __nv_bfloat16* dst = yb + ((ptrdiff_t)row_lo * W + icol) * C + c0;
#pragma unroll
                for (int rs = 0; rs < 2; ++rs) {
                    if (CPW == 1 && rs != (warp & 1)) continue;          // 16 warps: two warps share a column group, one row each
                    const int row = row_lo + rs;
                    if (row < 0 || row >= H) continue;
                    const uint32_t sel = rs ? 0x7632u : 0x5410u;
                    uint4 o0, o1;
                    o0.x = prmt(w[0], w[1], sel); o0.y = prmt(w[2], w[3], sel); o0.z = prmt(w[4], w[5], sel); o0.w = prmt(w[6], w[7], sel);
                    o1.x = prmt(w[8], w[9], sel); o1.y = prmt(w[10], w[11], sel); o1.z = prmt(w[12], w[13], sel); o1.w = prmt(w[14], w[15], sel);
                    uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)rs * W * C);
                    d4[0] = o0;
                    d4[1] = o1;
                }
            }
        }
        base40 += 16;
        base40 -= base40 >= 40 ? 40 : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Warp-specialised generation of the same kernel: 8 COMPUTE warps (2 channels each, 216 registers after setmaxnreg) run
// conv0 -> hpass -> vpass with no CTA barrier at all — a warp only synchronises with itself — and 4 IO warps (64 registers) do the
// NHWC <-> planar transposes: they fetch the x rows of step s + 1 while step s computes, store them into X as soon as every
// compute warp has left conv0(s), and write the output block of step s - 1 out of O.  Hand-offs are four mbarriers:
//     full_x  (IO -> compute)   X holds the rows of step s            empty_x (compute -> IO)   conv0(s) has read X
//     full_o  (compute -> IO)   O holds the output block of step s    empty_o (IO -> compute)   O has been written out
// Measured reason (ncu source view of the barrier-synchronised kernel above): with every warp in the same phase between two
// __syncthreads the tensor / LSU / issue pipes take turns instead of overlapping (issue 28 %, HMMA 21 %), and the prefetched x
// registers share scoreboards with the shared-memory loads of the MMA code, so LDSM consumers wait for global loads.
constexpr int W_CT = 256;                       // compute threads; NIO = IO warps (4: 64 / 216 registers, 8: 48 / 208)

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {          // pinned where it is written (volatile): issue order = program order
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int K0, int K1, int K2, bool IDENT, int NIO>
__global__ void __launch_bounds__(W_CT + 32 * NIO, 1)
msca_sdw_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, const float* __restrict__ w0, const float* __restrict__ b0,
                const float* __restrict__ w1cat, const float* __restrict__ b1cat, const float* __restrict__ w2cat,
                const float* __restrict__ b2cat, const float* __restrict__ res, const SdmParams p) {
    using L = Lay<K0, K1, K2>;
    constexpr int NB = L::NB, CPW = 2, W_IO = 32 * NIO, W_NT = W_CT + W_IO;
    // setmaxnreg: an increase draws only from what this CTA released.  launch registers = 65536 / W_NT rounded down to 8:
    //   NIO 4: 168 -> IO 64 (frees 128 x 104 = 13312), compute 216 (needs 256 x 48 = 12288)
    //   NIO 8: 128 -> IO 48 (frees 256 x 80 = 20480),  compute 208 (needs 256 x 80 = 20480)
    constexpr int REG_IO = NIO == 4 ? 64 : 48, REG_CT = NIO == 4 ? 216 : 208;
    constexpr int XT = 40 / NIO;                                    // x staging tasks per IO warp (40 = 20 rows x 2 channel halves)
    constexpr int XB = NIO == 4 ? 4 : 3;                            // rounds per batch (8 registers each)
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sb = ptx::smem_u32(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int strip = blockIdx.x % p.strips, cg = blockIdx.x / p.strips, b = blockIdx.y;
    const int w0c = strip * 32, c0 = cg * 16;
    const int H = p.H, W = p.W, C = p.C;
    const uint32_t bar_fx = sb + L::BAR_OFF, bar_ex = bar_fx + 8, bar_fo = bar_fx + 16, bar_eo = bar_fx + 24;

    // ---- prologue (all 12 warps): zero the a0 / T rings, stage border residues and weights, init barriers ----
    for (int i = tid; i < (L::O_OFF - L::A_OFF) / 16; i += W_NT) reinterpret_cast<uint4*>(smem + L::A_OFF)[i] = make_uint4(0, 0, 0, 0);
    if (res != nullptr && p.res_p > 0) {
        float* rs = reinterpret_cast<float*>(smem + L::R_OFF);
        for (int i = tid; i < 2 * 16 * 16; i += W_NT) {
            const int half = i >> 8, rr = (i >> 4) & 15, c = i & 15;
            rs[i] = rr < p.res_p ? res[((size_t)half * C + c0 + c) * p.res_p + rr] : 0.f;
        }
    }
    constexpr int KS[3] = {K0, K1, K2};
    constexpr int KOFF[3] = {0, K0, K0 + K1};
    constexpr int WO_B0 = 25 * 16, WO_BR0 = WO_B0 + 16;
    constexpr int WO_BR[3] = {WO_BR0, WO_BR0 + (2 * K0 + 2) * 16, WO_BR0 + (2 * K0 + 2) * 16 + (2 * K1 + 2) * 16};   // [w1 k][b1][w2 k][b2]
    float* wst = reinterpret_cast<float*>(smem + L::O_OFF);
    for (int e = tid; e < 25 * 16; e += W_NT) wst[e] = w0[(size_t)(e >> 4) * C + c0 + (e & 15)];
    if (tid < 16) wst[WO_B0 + tid] = b0 ? b0[c0 + tid] : 0.f;
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        const int k = KS[i];
        float* d = wst + WO_BR[i];
        for (int e = tid; e < k * 16; e += W_NT) {
            d[e] = w1cat[(size_t)(KOFF[i] + (e >> 4)) * C + c0 + (e & 15)];
            d[(k + 1) * 16 + e] = w2cat[(size_t)(KOFF[i] + (e >> 4)) * C + c0 + (e & 15)];
        }
        if (tid < 16) {
            d[k * 16 + tid] = b1cat ? b1cat[(size_t)i * C + c0 + tid] : 0.f;
            d[(2 * k + 1) * 16 + tid] = b2cat ? b2cat[(size_t)i * C + c0 + tid] : 0.f;
        }
    }
    if (tid == 0) {
        ptx::mbar_init(bar_fx, W_IO);
        ptx::mbar_init(bar_ex, W_CT);
        ptx::mbar_init(bar_fo, W_CT);
        ptx::mbar_init(bar_eo, W_IO);
        ptx::fence_mbar_init();
    }
    __syncthreads();

    if (warp >= W_CT / 32) {
        // =========================================== IO warps ===========================================
        // Fragments are built by the compute warps from the O region: the first write-out happens long after (full_o of step 0).
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REG_IO));
        const int iw = warp - W_CT / 32, itid = tid - W_CT;
        const __nv_bfloat16* xb = x + (size_t)b * H * W * C;
        __nv_bfloat16* yb = y + (size_t)b * H * W * C;
        const int p2 = (res != nullptr && p.res_p > 0) ? min(H, p.res_p) : 0;
        // x staging: warp task wt = iw + NIO r = (row, channel half hf), lane = column pair: pixels 2 lane, 2 lane + 1
        const int xcol = w0c - 13 + 2 * lane;
        const bool xc0 = xcol >= 0 && xcol < W, xc1 = xcol + 1 >= 0 && xcol + 1 < W;
        auto write_out = [&](int s) {
#pragma unroll 1
            for (int k = 0; k < 8 / NIO; ++k) {
                const int col = 4 * (iw + NIO * k) + (lane >> 3), rp = lane & 7, icol = w0c + col;
                const int row_lo = 16 * s - p.lag + 2 * rp;
                if (icol < W && row_lo + 1 >= 0 && row_lo < H) {
                    const uint32_t oa = sb + L::O_OFF + col * 32 + ((rp ^ (((col >> 2) & 1) * 4)) << 2);
                    uint32_t w[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) w[c] = lds32(oa + c * M_OPLANE);
                    if (p2 > 0 && (row_lo < p2 || row_lo + 1 >= H - p2)) {
                        // FixPaddingBias (MscaRep's border residues): top rows then bottom rows, each add materialised
                        const float* rs_top = reinterpret_cast<const float*>(smem + L::R_OFF);
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            float f[2] = {__uint_as_float(w[c] << 16), __uint_as_float(w[c] & 0xffff0000u)};
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int rr = row_lo + e;
                                if (rr >= 0 && rr < H) {
                                    if (rr < p2) f[e] = __bfloat162float(__float2bfloat16_rn(f[e] + rs_top[rr * 16 + c]));
                                    if (rr >= H - p2) f[e] += rs_top[256 + (p.res_p - (H - rr)) * 16 + c];
                                }
                            }
                            w[c] = pack2(f[0], f[1]);
                        }
                    }
                    __nv_bfloat16* dst = yb + ((ptrdiff_t)row_lo * W + icol) * C + c0;
#pragma unroll
                    for (int rs = 0; rs < 2; ++rs) {
                        const int row = row_lo + rs;
                        if (row < 0 || row >= H) continue;
                        const uint32_t sel = rs ? 0x7632u : 0x5410u;
                        uint4 o0, o1;
                        o0.x = prmt(w[0], w[1], sel); o0.y = prmt(w[2], w[3], sel); o0.z = prmt(w[4], w[5], sel); o0.w = prmt(w[6], w[7], sel);
                        o1.x = prmt(w[8], w[9], sel); o1.y = prmt(w[10], w[11], sel); o1.z = prmt(w[12], w[13], sel); o1.w = prmt(w[14], w[15], sel);
                        ptx::st_global_v8(dst + (size_t)rs * W * C, o0, o1);
                    }
                }
            }
        };
        (void)itid;
        for (int s = 0; s < p.steps; ++s) {
            const int r0 = 16 * s - 2;
#pragma unroll 1
            for (int bt = 0; bt < XT; bt += XB) {
                uint4 xr[XB][2];
#pragma unroll
                for (int r = 0; r < XB; ++r) {
                    const int wt = iw + NIO * (bt + r), row = r0 + (wt >> 1), hf = wt & 1;
                    const bool rin = row >= 0 && row < H && bt + r < XT;
                    const __nv_bfloat16* src = xb + ((ptrdiff_t)row * W + xcol) * C + c0 + 8 * hf;
                    xr[r][0] = (rin && xc0) ? ldg_nc_v4(src) : make_uint4(0, 0, 0, 0);
                    xr[r][1] = (rin && xc1) ? ldg_nc_v4(src + C) : make_uint4(0, 0, 0, 0);
                }
                if (bt == 0 && s > 0) ptx::mbar_wait(bar_ex, (s - 1) & 1);      // every compute warp has left conv0(s - 1)
#pragma unroll
                for (int r = 0; r < XB; ++r) {
                    if (bt + r < XT) {
                        const int wt = iw + NIO * (bt + r);
                        const uint32_t a = sb + L::X_OFF + 8 * (wt & 1) * M_XPLANE + (wt >> 1) * M_XPITCH + lane * 4;
                        const uint32_t u0[4] = {xr[r][0].x, xr[r][0].y, xr[r][0].z, xr[r][0].w}, u1[4] = {xr[r][1].x, xr[r][1].y, xr[r][1].z, xr[r][1].w};
#pragma unroll
                        for (int c = 0; c < 8; ++c) sts32(a + c * M_XPLANE, prmt(u0[c >> 1], u1[c >> 1], (c & 1) ? 0x7632u : 0x5410u));
                    }
                }
            }
            ptx::mbar_arrive(bar_fx);                                         // X holds the rows of step s
            if (s > 0) {
                ptx::mbar_wait(bar_fo, (s - 1) & 1);
                write_out(s - 1);
                ptx::mbar_arrive(bar_eo);
            }
        }
        ptx::mbar_wait(bar_fo, (p.steps - 1) & 1);
        write_out(p.steps - 1);
        return;
    }

    // =========================================== compute warps ===========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REG_CT));
    unsigned long long bw0[CPW][5], bh[CPW][3][2], bv[CPW][3][2];
    float bias0[CPW], bias1[CPW][3], bias2[CPW][3];
#pragma unroll
    for (int cw = 0; cw < CPW; ++cw) {
        const int ch = warp * CPW + cw;
        bias0[cw] = wst[WO_B0 + ch];
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) toep_frag<5>(bw0[cw][dy], wst + dy * 5 * 16, ch, 0, 0, g, t);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int q = 0; q < 2; ++q) { bh[cw][i][q] = 0ull; bv[cw][i][q] = 0ull; }
            bias1[cw][i] = bias2[cw][i] = 0.f;
        }
        if constexpr (K0 > 0) {
            const float* d = wst + WO_BR[0];
#pragma unroll
            for (int q = 0; q < Br<K0>::NM; ++q) { toep_frag<K0>(bh[cw][0][q], d, ch, q, Br<K0>::HOFF, g, t); toep_frag<K0>(bv[cw][0][q], d + (K0 + 1) * 16, ch, q, 0, g, t); }
            bias1[cw][0] = d[K0 * 16 + ch]; bias2[cw][0] = d[(2 * K0 + 1) * 16 + ch];
        }
        if constexpr (K1 > 0) {
            const float* d = wst + WO_BR[1];
#pragma unroll
            for (int q = 0; q < Br<K1>::NM; ++q) { toep_frag<K1>(bh[cw][1][q], d, ch, q, Br<K1>::HOFF, g, t); toep_frag<K1>(bv[cw][1][q], d + (K1 + 1) * 16, ch, q, 0, g, t); }
            bias1[cw][1] = d[K1 * 16 + ch]; bias2[cw][1] = d[(2 * K1 + 1) * 16 + ch];
        }
        if constexpr (K2 > 0) {
            const float* d = wst + WO_BR[2];
#pragma unroll
            for (int q = 0; q < Br<K2>::NM; ++q) { toep_frag<K2>(bh[cw][2][q], d, ch, q, Br<K2>::HOFF, g, t); toep_frag<K2>(bv[cw][2][q], d + (K2 + 1) * 16, ch, q, 0, g, t); }
            bias1[cw][2] = d[K2 * 16 + ch]; bias2[cw][2] = d[(2 * K2 + 1) * 16 + ch];
        }
    }
    // O may be overwritten (vpass of step 0) only after EVERY compute warp has read its weights: one named barrier among them
    asm volatile("bar.sync 1, %0;" ::"n"(W_CT) : "memory");

    // ---- per-lane address constants (first channel of the warp; channel cw adds cw * plane size as an immediate) ----
    const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;          // ldmatrix.x4: row / chunk half supplied by this lane
    const uint32_t x_ld = sb + L::X_OFF + warp * CPW * M_XPLANE + lrow * M_XPITCH + lhi * 16;     // + dy * pitch + 32 q
    const uint32_t a_ld = sb + L::A_OFF + warp * CPW * M_APLANE + lrow * M_APITCH + lhi * 16;     // + block * 16 pitch + 32 q (+16)
    const uint32_t a_st = sb + L::A_OFF + warp * CPW * M_APLANE + g * M_APITCH + t * 4;           // + block * 16 pitch + 16 j (+ 8 pitch)
    const uint32_t a_id = sb + L::A_OFF + warp * CPW * M_APLANE + (g + 11) * 2;                   // + slot * pitch + 2 (16 mp + 8 hc)
    const int osw = 4 * ((g >> 2) & 1);
    const uint32_t o_st = sb + L::O_OFF + warp * CPW * M_OPLANE + g * 32;                         // + 32 (16 mp + 8 hc) + 4 ((4 np + t) ^ osw)
    const uint32_t t_w0 = sb + L::T0_OFF + warp * CPW * Br<K0>::RING * M_TPITCH, t_w1 = sb + L::T1_OFF + warp * CPW * Br<K1>::RING * M_TPITCH,
                   t_w2 = sb + L::T2_OFF + warp * CPW * Br<K2>::RING * M_TPITCH;
    int base40 = 0;                                     // (16 s) mod 40
    const bool cols_in = w0c - 11 >= 0 && w0c + 45 <= W;        // every a0 position of the CTA is an image column
    uint32_t cmask = 0;                                 // bit 2j + e: a0 position 8j + 2t + e is an image column
#pragma unroll
    for (int i = 0; i < 14; ++i) {
        const int c = w0c - 11 + 8 * (i >> 1) + 2 * t + (i & 1);
        if (c >= 0 && c < W) cmask |= 1u << i;
    }

    for (int s = 0; s < p.steps; ++s) {
        const int base32 = (s & 1) * 16;
        // ================= conv0: x rows [16s-2, 16s+18) -> a0 rows [16s, 16s+16), positions 0..55 =================
        ptx::mbar_wait(bar_fx, s & 1);
        if (16 * s < H) {
            const bool rows_in = 16 * s + 16 <= H;
            const bool r0in = 16 * s + g < H, r1in = 16 * s + g + 8 < H;
#pragma unroll
            for (int cw = 0; cw < CPW; ++cw) {
                const float cq[4] = {bias0[cw], bias0[cw], bias0[cw], bias0[cw]};
                float acc[7][4];
#pragma unroll
                for (int dy = 0; dy < 5; ++dy) {
                    // odd tiles start on an odd chunk: their A quads are assembled with register moves (cheaper than 3 more ldmatrix)
                    uint32_t q4[4][4];                  // q4[q] = {chunk 2q rows g | g+8, chunk 2q+1 rows g | g+8}
#pragma unroll
                    for (int q = 0; q < 4; ++q) ldsm4(q4[q], x_ld + cw * M_XPLANE + dy * M_XPITCH + q * 32);
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const int q = j >> 1;
                        const uint32_t a0 = (j & 1) ? q4[q][2] : q4[q][0], a1 = (j & 1) ? q4[q][3] : q4[q][1];
                        const uint32_t a2 = (j & 1) ? q4[q + 1][0] : q4[q][2], a3 = (j & 1) ? q4[q + 1][1] : q4[q][3];
                        if (dy == 0) mma16816c(acc[j], a0, a1, a2, a3, bw0[cw][dy], cq);
                        else mma16816(acc[j], a0, a1, a2, a3, bw0[cw][dy]);
                    }
                }
                if (cw == CPW - 1) ptx::mbar_arrive(bar_ex);      // this warp has read X (every ldmatrix result has been consumed)
                // outside the image (rows or columns) the strip convs see ZERO padding, not conv0 of a padded input
                const uint32_t d0 = a_st + cw * M_APLANE + base32 * M_APITCH, d1 = d0 + 8 * M_APITCH;
                if (rows_in && cols_in) {
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        sts32(d0 + j * 16, pack2(acc[j][0], acc[j][1]));
                        sts32(d1 + j * 16, pack2(acc[j][2], acc[j][3]));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const uint32_t m = (((cmask >> (2 * j)) & 1u) * 0xffffu) | (((cmask >> (2 * j + 1)) & 1u) * 0xffff0000u);
                        sts32(d0 + j * 16, r0in ? (pack2(acc[j][0], acc[j][1]) & m) : 0u);
                        sts32(d1 + j * 16, r1in ? (pack2(acc[j][2], acc[j][3]) & m) : 0u);
                    }
                }
            }
        } else {
            ptx::mbar_arrive(bar_ex);
        }
        __syncwarp();

        // per-step ring addresses (shared by the warp's channels)
        uint32_t st0[3] = {0, 0, 0}, st1[3] = {0, 0, 0}, va[3][3];
        if constexpr (K0 > 0) sdm_ring_addrs<K0>(Br<K0>::RING == 32 ? base32 : base40, p.lag, lane, g, t, st0[0], st1[0], va[0]);
        if constexpr (K1 > 0) sdm_ring_addrs<K1>(Br<K1>::RING == 32 ? base32 : base40, p.lag, lane, g, t, st0[1], st1[1], va[1]);
        if constexpr (K2 > 0) sdm_ring_addrs<K2>(Br<K2>::RING == 32 ? base32 : base40, p.lag, lane, g, t, st0[2], st1[2], va[2]);
        const bool rin0 = 16 * s + g < H, rin1 = 16 * s + g + 8 < H;

        // ================= hpass: a0 block -> T_k rows [16s, 16s+16) of every branch =================
#pragma unroll
        for (int cw = 0; cw < CPW; ++cw) {
            uint32_t ae[3][4], ao[3][4];
            const uint32_t ra = a_ld + cw * M_APLANE + base32 * M_APITCH;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                ldsm4(ae[q], ra + q * 32);
                ldsm4(ao[q], ra + q * 32 + 16);
            }
            if constexpr (K0 > 0) sdm_hpass<K0>(t_w0 + cw * Br<K0>::RING * M_TPITCH + st0[0], t_w0 + cw * Br<K0>::RING * M_TPITCH + st1[0], ae, ao, bh[cw][0], bias1[cw][0], rin0, rin1);
            if constexpr (K1 > 0) sdm_hpass<K1>(t_w1 + cw * Br<K1>::RING * M_TPITCH + st0[1], t_w1 + cw * Br<K1>::RING * M_TPITCH + st1[1], ae, ao, bh[cw][1], bias1[cw][1], rin0, rin1);
            if constexpr (K2 > 0) sdm_hpass<K2>(t_w2 + cw * Br<K2>::RING * M_TPITCH + st0[2], t_w2 + cw * Br<K2>::RING * M_TPITCH + st1[2], ae, ao, bh[cw][2], bias1[cw][2], rin0, rin1);
        }
        __syncwarp();
        if (s > 0) ptx::mbar_wait(bar_eo, (s - 1) & 1);                  // the output block of step s - 1 has left O

        // ================= vpass: T_k rings -> output rows [16s-10, 16s+6) into O (+ identity) =================
        {
            uint32_t idoff[2][2];
#pragma unroll
            for (int np = 0; np < 2; ++np) {
                idoff[np][0] = ((base32 - p.lag + 8 * np + 2 * t) & 31) * M_APITCH;
                idoff[np][1] = ((base32 - p.lag + 8 * np + 2 * t + 1) & 31) * M_APITCH;
            }
#pragma unroll
            for (int cw = 0; cw < CPW; ++cw) {
#pragma unroll
                for (int mp = 0; mp < 2; ++mp) {
                    uint32_t tot[2][2];
                    uint32_t v3[3];
                    if constexpr (K0 > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v3[i] = (t_w0 + cw * Br<K0>::RING * M_TPITCH + va[0][i]) ^ (mp << 5);
                        sdm_vbranch<K0>(v3, bv[cw][0], bias2[cw][0], true, tot);
                    }
                    if constexpr (K1 > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v3[i] = (t_w1 + cw * Br<K1>::RING * M_TPITCH + va[1][i]) ^ (mp << 5);
                        sdm_vbranch<K1>(v3, bv[cw][1], bias2[cw][1], false, tot);
                    }
                    if constexpr (K2 > 0) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) v3[i] = (t_w2 + cw * Br<K2>::RING * M_TPITCH + va[2][i]) ^ (mp << 5);
                        sdm_vbranch<K2>(v3, bv[cw][2], bias2[cw][2], false, tot);
                    }
#pragma unroll
                    for (int np = 0; np < 2; ++np) {
#pragma unroll
                        for (int hc = 0; hc < 2; ++hc) {
                            uint32_t v = tot[np][hc];
                            if constexpr (IDENT) {
                                const uint32_t ia = a_id + cw * M_APLANE + (16 * mp + 8 * hc) * 2;
                                v = hadd2(v, lds16(ia + idoff[np][0]) | (lds16(ia + idoff[np][1]) << 16));
                            }
                            sts32(o_st + cw * M_OPLANE + (16 * mp + 8 * hc) * 32 + ((((4 * np) ^ osw) + t) << 2), v);
                        }
                    }
                }
            }
        }
        ptx::mbar_arrive(bar_fo);                                         // this warp's planes of the output block are in O
        base40 += 16;
        base40 -= base40 >= 40 ? 40 : 0;
    }
}

template <int K0, int K1, int K2, bool IDENT, int NIO>
int sdw_launch_k(const void* x, void* y, const float* w0, const float* b0, const float* w1cat, const float* b1cat, const float* w2cat,
                 const float* b2cat, const float* res, const SdmParams& p, cudaStream_t st) {
    using L = Lay<K0, K1, K2>;
    static std::atomic<unsigned long long> done{0};
    CNA_CUDA(ensure_dyn_smem(msca_sdw_kernel<K0, K1, K2, IDENT, NIO>, L::TOTAL, done));
    dim3 grid(p.strips * (p.C / 16), p.B);
    msca_sdw_kernel<K0, K1, K2, IDENT, NIO><<<grid, W_CT + 32 * NIO, L::TOTAL, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, w0, b0, w1cat,
                                                                                    b1cat, w2cat, b2cat, res, p);
    CNA_LAUNCH_OK();
    return 0;
}

template <int K0, int K1, int K2, int CPW, bool IDENT>
int sdm_launch_k(const void* x, void* y, const float* w0, const float* b0, const float* w1cat, const float* b1cat, const float* w2cat,
                 const float* b2cat, const float* res, const SdmParams& p, cudaStream_t st) {
    using L = Lay<K0, K1, K2>;
    static std::atomic<unsigned long long> done{0};
    CNA_CUDA(ensure_dyn_smem(msca_sdm_kernel<K0, K1, K2, CPW, IDENT>, L::TOTAL, done));
    dim3 grid(p.strips * (p.C / 16), p.B);
    msca_sdm_kernel<K0, K1, K2, CPW, IDENT><<<grid, M_THREADS / CPW, L::TOTAL, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, w0, b0, w1cat,
                                                                                    b1cat, w2cat, b2cat, res, p);
    CNA_LAUNCH_OK();
    return 0;
}

template <int K0, int K1, int K2>
int sdm_launch_t(const void* x, void* y, const float* w0, const float* b0, const float* w1cat, const float* b1cat, const float* w2cat,
                 const float* b2cat, const float* res, const SdmParams& p, int max_smem, cudaStream_t st) {
    using L = Lay<K0, K1, K2>;
    if (L::TOTAL > max_smem) return -1;
    // Measured on B200 (13 MSCAN-T blocks, batch 32): the warp-specialised kernel wins (three cascades: 0.98 vs 1.15 ms for the
    // barrier-synchronised one; one cascade with 8 IO warps: 0.92 vs 1.02 ms).  Diagnostic A/B switch: CNA_MSCA_WS=0 runs the barrier kernel.
    static const int force = [] { const char* e = getenv("CNA_MSCA_WS"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
    const bool ws = force >= 0 ? force == 1 : true;
    // IO warps: a single cascade leaves the compute warps so little to do that four IO warps bind it (d1+fix, 13 blocks: 1.12 ms with 4,
    // 0.92 ms with 8); three cascades are indifferent (0.98 / 0.99).  CNA_MSCA_NIO=4 | 8 forces one.
    static const int nio_force = [] { const char* e = getenv("CNA_MSCA_NIO"); return e ? (e[0] == '8' ? 8 : 4) : 0; }();
    const int nio = nio_force ? nio_force : (L::NB == 1 ? 8 : 4);
    if (ws && nio == 8) {
        return p.identity ? sdw_launch_k<K0, K1, K2, true, 8>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, st)
                          : sdw_launch_k<K0, K1, K2, false, 8>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, st);
    }
    if (ws) {
        return p.identity ? sdw_launch_k<K0, K1, K2, true, 4>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, st)
                          : sdw_launch_k<K0, K1, K2, false, 4>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, st);
    }
    return p.identity ? sdm_launch_k<K0, K1, K2, 2, true>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, st)
                      : sdm_launch_k<K0, K1, K2, 2, false>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, st);
}

}  // namespace

// Launches the tensor-path kernel; returns -1 when the configuration is not covered (caller falls back), 0 / 1 like the C ABI.
int msca_sdm_launch(const void* x, void* y, const float* w0, const float* b0, const float* w1cat, const float* b1cat, const float* w2cat,
                    const float* b2cat, const float* res, int res_p, const int* ks_host, int nb, int identity, int B, int H, int W, int C,
                    int dtype, cudaStream_t st) {
    if (dtype != CNA_BF16 || nb < 1 || nb > 3 || B > 65535 || C % 16 != 0) return -1;
    if (res && res_p > 16) return -1;
    if ((long long)H * W * C >= (1LL << 30)) return -1;                  // 32-bit offsets inside an image
    SdmParams p{};
    p.B = B; p.H = H; p.W = W; p.C = C; p.identity = identity; p.res_p = res ? res_p : 0;
    p.strips = ceil_div(W, 32);
    p.lag = H <= 16 ? 0 : M_P;                                           // a one-step image needs no lag: rows below it are the zeroed ring
    p.steps = ceil_div(H + p.lag, 16);                                   // last step's output block [16s-lag, 16s-lag+16) covers row H-1
    DevInfo di;
    if (dev_info(&di)) return 1;
    const int k0 = ks_host[0], k1 = nb > 1 ? ks_host[1] : 0, k2 = nb > 2 ? ks_host[2] : 0;
#define CNA_SDM_CASE(A, Bk, Ck) \
    if (k0 == A && k1 == Bk && k2 == Ck) return sdm_launch_t<A, Bk, Ck>(x, y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p, di.max_smem, st);
    CNA_SDM_CASE(7, 11, 21)       // the MSCA block as shipped (layers/msca.py:18-24)
    CNA_SDM_CASE(21, 0, 0)        // MscaRep rank 1 (core/msca_rep.py:154-171)
    CNA_SDM_CASE(21, 21, 0)       // MscaRep rank 2
#undef CNA_SDM_CASE
    return -1;
}

}  // namespace cna
