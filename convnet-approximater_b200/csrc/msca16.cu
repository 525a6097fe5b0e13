// Fused MSCA "sd" chain, 16-row march:   s = [conv0(x)] + sum_k cascade_k(conv0(x))   in ONE pass over HBM.
//
// Contract of the reference modules it replaces (src/approx/layers/msca.py:18-26, depth_seperable_conv.py:12-16, 45-47, 57-61;
// MscaRep's rank-r forms core/msca_rep.py:44-55): read x once, write s once, every tensor the reference's bf16 modules
// materialise (conv0 out, each 1xk out, each branch out, every partial sum) rounded to bf16 at the same point.
//
// The chain is FP32-pipe bound (103 MAC per element against 4 bytes), and on this part an FP32-bound kernel is really an
// ISSUE-bound kernel: the packed FFMA2 needs half the issue slots of its pipe time, everything else competes for the other
// half.  So the kernel is built to spend instructions on FMAs:
//   * a CTA (512 threads) owns (image, 16 bf16 / 8 fp32 channels, 32 output columns) and marches down the image 16 rows per step;
//   * thread tasks are register tiles along the filter axis — 2 rows x 7 columns for the 5x5 conv0 (5.3 FFMA2 per shared-memory
//     load), runs of 8 outputs for the strips (4-6 per load) — with the weights of the running tap (row) in registers;
//   * all sliding windows are rings of aligned 16-row BLOCKS (a0: 2, T_k: 3 for k = 21, else 2) addressed as
//     `block base + immediate`: no modulo, no per-row index math inside a task;
//   * accumulators start at the bias; branch outputs, partial sums and the identity branch are combined with PACKED bf16 adds
//     (one HADD2.BF16 per add and rounding, the a0 word is used as it sits in shared memory) instead of unpack / add / round chains;
//   * the x rows of the NEXT step are prefetched into registers before the current step's arithmetic starts; their addresses
//     and the column masks are computed once per kernel.
//
// Per step s (rows are image rows; a0 = conv0 output, T_k = 1xk output):
//     conv0 : x rows [16s-2, 16s+18)            -> a0 rows [16s, 16s+16), 52 columns (32 + 10 halo each side)
//     hpass : a0 rows [16s, 16s+16)             -> T_k rows [16s, 16s+16), 32 columns, every branch k
//     vpass : T_k rows [16s-10-k/2, 16s+6+k/2)  -> output rows [16s-10, 16s+6)   (+ identity a0, + border residues)
// Lanes run over the 8 four-byte units of a pixel and 4 neighbouring tasks; row pitches are 8 words mod 32 (and conv0's runs are
// 7 columns = 24 words mod 32 apart), so every warp access is one conflict-free 128-byte wavefront.
#include "common.cuh"

namespace cna {

constexpr int S16_THREADS = 512;
constexpr int S16_RB = 16;         // rows per step
constexpr int S16_WT = 32;         // output columns per CTA
constexpr int S16_P = 10;          // output lag / halo of the widest strip (21 taps)
constexpr int S16_UNITS = 8;       // 4-byte units per pixel held by a CTA
constexpr int S16_MAXB = 3;
constexpr int S16_XC = S16_WT + 2 * S16_P + 4;   // 56 x columns
constexpr int S16_AC = S16_WT + 2 * S16_P;       // 52 a0 columns
constexpr int S16_XR = S16_RB + 4;               // 20 x rows per step

__host__ __device__ constexpr int s16_pitch(int cols) { return ((cols * S16_UNITS + 31) / 32) * 32 + 8; }   // words
__host__ __device__ constexpr int s16_nblk(int k) { return k / 2 > 6 ? 3 : 2; }                               // ring blocks of T_k
constexpr int S16_XP = s16_pitch(S16_XC), S16_AP = s16_pitch(S16_AC), S16_TP = s16_pitch(S16_WT);
constexpr int S16_OFF_A0 = S16_XR * S16_XP;                       // word offsets inside dynamic shared memory
constexpr int S16_OFF_T = S16_OFF_A0 + 2 * S16_RB * S16_AP;

struct S16Params {
    int B, H, W, C;
    int nb;
    int ks[S16_MAXB];
    int identity;
    int res_p;
    int strips;
    int steps;
};

template <typename T> struct U16;
template <> struct U16<__nv_bfloat16> {
    static constexpr int NCH = 2;
    __device__ static __forceinline__ float2 unpack(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
    __device__ static __forceinline__ uint32_t pack(float2 f) {
        __nv_bfloat162 v = __floats2bfloat162_rn(f.x, f.y);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    // a + b with the sum rounded to bf16, both channels at once (HADD2.BF16)
    __device__ static __forceinline__ uint32_t addw(uint32_t a, uint32_t b) {
        __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
        return *reinterpret_cast<uint32_t*>(&r);
    }
    __device__ static __forceinline__ float2 ldw(const float* w) { return *reinterpret_cast<const float2*>(w); }   // 2 channels
};
template <> struct U16<float> {
    static constexpr int NCH = 1;
    __device__ static __forceinline__ float2 unpack(uint32_t u) { return make_float2(__uint_as_float(u), 0.f); }
    __device__ static __forceinline__ uint32_t pack(float2 f) { return __float_as_uint(f.x); }
    __device__ static __forceinline__ uint32_t addw(uint32_t a, uint32_t b) { return __float_as_uint(__uint_as_float(a) + __uint_as_float(b)); }
    __device__ static __forceinline__ float2 ldw(const float* w) { return make_float2(*w, 0.f); }
};

// acc += v * w on a channel pair: one packed FFMA2 (fma.rn.f32x2) for bf16 units, a scalar FFMA for fp32 units
template <int NCH>
__device__ __forceinline__ void fma_u(float2& acc, float2 v, float2 w) {
    if constexpr (NCH == 2) {
        unsigned long long ra = *reinterpret_cast<unsigned long long*>(&acc), rx = *reinterpret_cast<unsigned long long*>(&v),
                           rw = *reinterpret_cast<unsigned long long*>(&w), rd;
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(rx), "l"(rw), "l"(ra));
        acc = *reinterpret_cast<float2*>(&rd);
    } else {
        acc.x = fmaf(v.x, w.x, acc.x);
    }
}

// Horizontal 1xK pass: T_k row (16s + r), columns [8 run, 8 run + 8)  <-  a0 row r of the current a0 block.
// a0row / trow: word offsets (into sm) of that a0 row / T row, unit `u` and the run's first column already applied.
template <typename T, int K>
__device__ __forceinline__ void s16_hpass(uint32_t* __restrict__ sm, int a0row, int trow, const float* __restrict__ w1,
                                          const float* __restrict__ b1, bool row_in, int u) {
    using U = U16<T>;
    constexpr int NCH = U::NCH, PK = K / 2;
    if (!row_in) {                      // outside the image: the vertical conv sees zero padding
#pragma unroll
        for (int i = 0; i < 8; ++i) sm[trow + i * S16_UNITS] = 0u;
        return;
    }
    float2 wt[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wt[j] = U::ldw(w1 + (j * S16_UNITS + u) * NCH);
    const float2 bz = U::ldw(b1 + u * NCH);
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bz;
    const int src = a0row + (S16_P - PK) * S16_UNITS;
#pragma unroll
    for (int m = 0; m < K + 7; ++m) {
        const float2 v = U::unpack(sm[src + m * S16_UNITS]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (m - i >= 0 && m - i < K) fma_u<NCH>(acc[i], v, wt[m - i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) sm[trow + i * S16_UNITS] = U::pack(acc[i]);
}

// Vertical Kx1 pass for output rows 16s - 10 + 8 HALF + i (i < 8) at one column: packed branch outputs (bias added, rounded).
// ring0/1/2: word offsets (lane offset applied) of the T_k blocks holding rows [16 (s - (NBLK-1) + q), +16), q = 0, 1, 2.
template <typename T, int K, int HALF>
__device__ __forceinline__ void s16_vpass(const uint32_t* __restrict__ sm, int ring0, int ring1, int ring2, const float* __restrict__ w2,
                                          const float* __restrict__ b2, int u, uint32_t (&out)[8]) {
    using U = U16<T>;
    constexpr int NCH = U::NCH, PK = K / 2, NBLK = s16_nblk(K);
    constexpr int OFF0 = 16 * (NBLK - 1) - S16_P - PK + 8 * HALF;      // ring-relative row of (i = 0, tap 0)
    float2 wt[K];
#pragma unroll
    for (int j = 0; j < K; ++j) wt[j] = U::ldw(w2 + (j * S16_UNITS + u) * NCH);
    const float2 bz = U::ldw(b2 + u * NCH);
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = bz;
#pragma unroll
    for (int m = 0; m < K + 7; ++m) {
        const int o = OFF0 + m;                                          // compile-time after unrolling
        const int base = o / 16 == 0 ? ring0 : (o / 16 == 1 ? ring1 : ring2);
        const float2 v = U::unpack(sm[base + (o % 16) * S16_TP]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (m - i >= 0 && m - i < K) fma_u<NCH>(acc[i], v, wt[m - i]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = U::pack(acc[i]);
}

template <typename T, int HALF>
__device__ __forceinline__ void s16_vphase(const S16Params& p, const uint32_t* __restrict__ sm, const float* wsm, int ob0, int ob1, int ob2,
                                           int tb0, int tb1, int tb2, const float* __restrict__ res, T* __restrict__ yb, int s, int u,
                                           int col, int w0c, int c0) {
    using U = U16<T>;
    constexpr int NCH = U::NCH, CB = S16_UNITS * NCH;
    const int lane_off = col * S16_UNITS + u;
    uint32_t total[8];
#pragma unroll 1
    for (int br = 0; br < p.nb; ++br) {        // one copy of each strip size's code: the dispatch is a runtime loop
        const int k = br == 0 ? p.ks[0] : (br == 1 ? p.ks[1] : p.ks[2]);
        const int obr = br == 0 ? ob0 : (br == 1 ? ob1 : ob2);
        const int tb = (br == 0 ? tb0 : (br == 1 ? tb1 : tb2)) + lane_off;
        const float* w2 = wsm + obr + k * CB + CB;
        const float* b2 = w2 + k * CB;
        constexpr int BLK = S16_RB * S16_TP;
        uint32_t o[8];
        if (k == 21) {            // 3-block ring: block (s - 2 + q) sits in slot (s + 1 + q) mod 3
            const int r0 = (s + 1) % 3, r1 = (s + 2) % 3, r2 = s % 3;
            s16_vpass<T, 21, HALF>(sm, tb + r0 * BLK, tb + r1 * BLK, tb + r2 * BLK, w2, b2, u, o);
        } else {                  // 2-block ring: block (s - 1 + q) sits in slot (s + 1 + q) & 1
            const int r0 = (s + 1) & 1, r1 = s & 1;
            if (k == 7) s16_vpass<T, 7, HALF>(sm, tb + r0 * BLK, tb + r1 * BLK, tb, w2, b2, u, o);
            else s16_vpass<T, 11, HALF>(sm, tb + r0 * BLK, tb + r1 * BLK, tb, w2, b2, u, o);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) total[i] = br == 0 ? o[i] : U::addw(total[i], o[i]);       // running sum, rounded per add
    }
    const int colw = w0c + col;
    const int H = p.H, W = p.W;
    if (colw >= W) return;
    const int row0 = 16 * s - S16_P + 8 * HALF;
    const size_t rstride = (size_t)W * p.C;
    T* yp = yb + ((ptrdiff_t)row0 * W + colw) * p.C + c0 + u * NCH;      // row0 may be negative: only dereferenced for rows >= 0
    const int p2 = p.res_p > 0 ? min(H, p.res_p) : 0;
    const int a0lane = S16_OFF_A0 + (col + S16_P) * S16_UNITS + u;
    const int a0b0 = a0lane + ((s + 1) & 1) * (S16_RB * S16_AP), a0b1 = a0lane + (s & 1) * (S16_RB * S16_AP);
    const bool all_in = row0 >= 0 && row0 + 7 < H;
    const bool border = p2 > 0 && (row0 < p2 || row0 + 7 >= H - p2);
    if (all_in && !border) {
        // interior rows: identity add on the packed words, one 4-byte store per output, a running row pointer
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t v = total[i];
            if (p.identity) {
                const int oa = 16 - S16_P + 8 * HALF + i;                // a0 row relative to block s-1: 6 .. 21 (compile time)
                v = U::addw(v, sm[(oa / 16 == 0 ? a0b0 : a0b1) + (oa % 16) * S16_AP]);
            }
            *reinterpret_cast<uint32_t*>(yp) = v;
            yp += rstride;
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i, yp += rstride) {
        const int row = row0 + i;
        if (row < 0 || row >= H) continue;
        uint32_t v = total[i];
        if (p.identity) {
            const int oa = 16 - S16_P + 8 * HALF + i;
            v = U::addw(v, sm[(oa / 16 == 0 ? a0b0 : a0b1) + (oa % 16) * S16_AP]);
        }
        if (p2 > 0 && (row < p2 || row >= H - p2)) {      // FixPaddingBias: top rows then bottom rows, each add materialised
            float2 f = U::unpack(v);
            float vv[2] = {f.x, f.y};
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int ch = c0 + u * NCH + c;
                if (row < p2) vv[c] += res[(size_t)ch * p.res_p + row];
            }
            f = U::unpack(U::pack(make_float2(vv[0], vv[1])));
            vv[0] = f.x; vv[1] = f.y;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int ch = c0 + u * NCH + c;
                if (row >= H - p2) vv[c] += res[((size_t)p.C + ch) * p.res_p + (p.res_p - (H - row))];
            }
            v = U::pack(make_float2(vv[0], vv[1]));
        }
        *reinterpret_cast<uint32_t*>(yp) = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(S16_THREADS, 1)
msca_sd16_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ w0, const float* __restrict__ b0,
                 const float* __restrict__ w1cat, const float* __restrict__ b1cat, const float* __restrict__ w2cat,
                 const float* __restrict__ b2cat, const float* __restrict__ res, const S16Params p) {
    using U = U16<T>;
    constexpr int NCH = U::NCH, CB = S16_UNITS * NCH;
    extern __shared__ uint32_t sm[];
    const int tid = threadIdx.x, u = tid & 7, pos = tid >> 3;            // pos 0..63
    const int strip = blockIdx.x % p.strips, cg = blockIdx.x / p.strips, b = blockIdx.y;
    const int w0c = strip * S16_WT, c0 = cg * CB;

    // ---- shared memory (word offsets): X | A0 (2 blocks) | T_k rings | weights ----
    int tb[S16_MAXB] = {0, 0, 0};                                        // fully unrolled uses only: stays in registers
    int wbase = S16_OFF_T;
#pragma unroll
    for (int i = 0; i < S16_MAXB; ++i) {
        tb[i] = wbase;
        if (i < p.nb) wbase += s16_nblk(p.ks[i]) * S16_RB * S16_TP;
    }
    float* wsm = reinterpret_cast<float*>(sm + wbase);
    const int o_w0 = 0, o_b0 = 25 * CB;
    int o_br[S16_MAXB] = {0, 0, 0};
    {
        int o = o_b0 + CB, ko = 0;
#pragma unroll
        for (int i = 0; i < S16_MAXB; ++i) {
            if (i >= p.nb) break;
            o_br[i] = o;
            const int k = p.ks[i];
            float* d = wsm + o;
            for (int e = tid; e < k * CB; e += S16_THREADS) {
                d[e] = w1cat[(size_t)(ko + e / CB) * p.C + c0 + e % CB];
                d[k * CB + CB + e] = w2cat[(size_t)(ko + e / CB) * p.C + c0 + e % CB];
            }
            for (int e = tid; e < CB; e += S16_THREADS) {
                d[k * CB + e] = b1cat ? b1cat[i * p.C + c0 + e] : 0.f;
                d[2 * k * CB + CB + e] = b2cat ? b2cat[i * p.C + c0 + e] : 0.f;
            }
            o += 2 * (k * CB + CB);
            ko += k;
        }
        for (int e = tid; e < 25 * CB; e += S16_THREADS) wsm[o_w0 + e] = w0[(size_t)(e / CB) * p.C + c0 + e % CB];
        for (int e = tid; e < CB; e += S16_THREADS) wsm[o_b0 + e] = b0 ? b0[c0 + e] : 0.f;
    }
    // rows above the image must read as zero padding: zero every ring once
    for (int q = tid; q < wbase; q += S16_THREADS) sm[q] = 0u;

    const int H = p.H, W = p.W;
    const T* xb = x + (size_t)b * H * W * p.C;
    T* yb = y + (size_t)b * H * W * p.C;

    // ---- x tile of a step: image rows [16s - 2, 16s + 18) x columns [w0c - 12, w0c + 44), two 16-byte chunks per pixel.
    // Per-thread chunk slots are fixed for the whole kernel: shared-memory offset, global offset inside a row block, column mask.
    constexpr int XCHUNKS = S16_XR * S16_XC * 2;                        // 2240
    constexpr int XPT = (XCHUNKS + S16_THREADS - 1) / S16_THREADS;      // 5 per thread
    int xs_off[XPT], xg_off[XPT], xrow[XPT];
#pragma unroll
    for (int q = 0; q < XPT; ++q) {
        const int i = tid + q * S16_THREADS;
        const int rr = i / (S16_XC * 2), rem = i - rr * (S16_XC * 2), cc = rem >> 1, half = rem & 1;
        const int col = w0c - S16_P - 2 + cc;
        xs_off[q] = rr * S16_XP + cc * S16_UNITS + half * 4;
        xg_off[q] = (rr * W + col) * p.C + c0 + half * (16 / (int)sizeof(T));
        xrow[q] = (i < XCHUNKS && col >= 0 && col < W) ? rr : -1000000;   // invalid slots never pass the row test below
    }
    uint4 xr[XPT];
    auto x_fetch = [&](int s) {
        const int r0 = 16 * s - 2;
        const T* base = xb + (ptrdiff_t)r0 * W * p.C;
#pragma unroll
        for (int q = 0; q < XPT; ++q) {
            const int row = r0 + xrow[q];
            xr[q] = (row >= 0 && row < H) ? __ldg(reinterpret_cast<const uint4*>(base + xg_off[q])) : make_uint4(0, 0, 0, 0);
        }
    };
    auto x_store = [&]() {
#pragma unroll
        for (int q = 0; q < XPT; ++q)
            if (tid + q * S16_THREADS < XCHUNKS) *reinterpret_cast<uint4*>(sm + xs_off[q]) = xr[q];
    };

    // ---- conv0 task: a0 rows 2rp, 2rp+1 of the block, columns [7 run, 7 run + 7); 4 neighbouring runs per warp
    const int c_run = pos & 7, c_rp = pos >> 3;
    uint32_t c_mask = 0;                                                 // bit i: a0 column 7 run + i is inside the image
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const int ac = c_run * 7 + i, col = w0c - S16_P + ac;
        if (ac < S16_AC && col >= 0 && col < W) c_mask |= 1u << i;
    }
    auto conv0 = [&](int s) {
        const int row = 16 * s + 2 * c_rp;                               // first of the two a0 rows (>= 0)
        const int dst = S16_OFF_A0 + ((s & 1) * S16_RB + 2 * c_rp) * S16_AP + (c_run * 7) * S16_UNITS + u;
        const float2 bz = U::ldw(wsm + o_b0 + u * NCH);
        float2 acc[2][7];
#pragma unroll
        for (int i = 0; i < 7; ++i) { acc[0][i] = bz; acc[1][i] = bz; }
        if (row < H) {
            const int src = (2 * c_rp) * S16_XP + (c_run * 7) * S16_UNITS + u;
            float2 wprev[5];
#pragma unroll
            for (int ri = 0; ri < 6; ++ri) {                             // x row ri of the task feeds tap row ri (upper a0 row) and ri-1 (lower)
                float2 wcur[5];
                if (ri < 5) {
#pragma unroll
                    for (int j = 0; j < 5; ++j) wcur[j] = U::ldw(wsm + o_w0 + ((ri * 5 + j) * S16_UNITS + u) * NCH);
                }
#pragma unroll
                for (int m = 0; m < 11; ++m) {
                    const float2 v = U::unpack(sm[src + ri * S16_XP + m * S16_UNITS]);
#pragma unroll
                    for (int i = 0; i < 7; ++i) {
                        if (m - i >= 0 && m - i < 5) {
                            if (ri < 5) fma_u<NCH>(acc[0][i], v, wcur[m - i]);
                            if (ri >= 1) fma_u<NCH>(acc[1][i], v, wprev[m - i]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 5; ++j) wprev[j] = wcur[j];
            }
        }
        // outside the image (rows or columns) the strip convs see ZERO padding, not conv0 of a padded input
        const bool in0 = row < H, in1 = row + 1 < H;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            if (c_run * 7 + i < S16_AC) {
                const bool cin = (c_mask >> i) & 1u;
                sm[dst + i * S16_UNITS] = (in0 && cin) ? U::pack(acc[0][i]) : 0u;
                sm[dst + S16_AP + i * S16_UNITS] = (in1 && cin) ? U::pack(acc[1][i]) : 0u;
            }
        }
    };

    x_fetch(0);
    __syncthreads();                     // rings zeroed, weights staged
    x_store();
    __syncthreads();
    const int h_r = pos & 15, h_run = pos >> 4;                          // hpass task: row, run of 8 columns
    for (int s = 0; s < p.steps; ++s) {
        const bool more = s + 1 < p.steps && 16 * (s + 1) - 2 < H;       // the next step still has x rows inside the image
        if (more) x_fetch(s + 1);
        if (16 * s < H) conv0(s);                                         // a0 rows below the image are never read
        __syncthreads();
        {   // horizontal passes of every branch on the new a0 block; the prefetched x rows go to shared memory afterwards
            const bool row_in = 16 * s + h_r < H;
            const int a0row = S16_OFF_A0 + ((s & 1) * S16_RB + h_r) * S16_AP + (h_run * 8) * S16_UNITS + u;
#pragma unroll 1
            for (int br = 0; br < p.nb; ++br) {
                const int k = br == 0 ? p.ks[0] : (br == 1 ? p.ks[1] : p.ks[2]);
                const int slot = s16_nblk(k) == 3 ? s % 3 : (s & 1);
                const int trow = (br == 0 ? tb[0] : (br == 1 ? tb[1] : tb[2])) + (slot * S16_RB + h_r) * S16_TP + (h_run * 8) * S16_UNITS + u;
                const float* w1 = wsm + (br == 0 ? o_br[0] : (br == 1 ? o_br[1] : o_br[2]));
                const float* b1 = w1 + k * CB;
                if (k == 7) s16_hpass<T, 7>(sm, a0row, trow, w1, b1, row_in, u);
                else if (k == 11) s16_hpass<T, 11>(sm, a0row, trow, w1, b1, row_in, u);
                else s16_hpass<T, 21>(sm, a0row, trow, w1, b1, row_in, u);
            }
            if (more) x_store();
        }
        __syncthreads();
        if (pos < 32) s16_vphase<T, 0>(p, sm, wsm, o_br[0], o_br[1], o_br[2], tb[0], tb[1], tb[2], res, yb, s, u, pos, w0c, c0);
        else s16_vphase<T, 1>(p, sm, wsm, o_br[0], o_br[1], o_br[2], tb[0], tb[1], tb[2], res, yb, s, u, pos - 32, w0c, c0);
        __syncthreads();
    }
}

size_t s16_smem_bytes(const S16Params& p, int cb) {
    size_t words = (size_t)S16_OFF_T;
    size_t wfl = 26 * cb;
    for (int i = 0; i < p.nb; ++i) {
        words += (size_t)s16_nblk(p.ks[i]) * S16_RB * S16_TP;
        wfl += 2 * (p.ks[i] * cb + cb);
    }
    return (words + wfl) * 4 + 256;      // + slack: conv0's last run reads a few words past an x row (values never stored)
}

// Launches the 16-row-march kernel; returns -1 when the configuration is not covered (caller falls back), 0 / 1 like the C ABI.
int msca_sd16_launch(const void* x, void* y, const float* w0, const float* b0, const float* w1cat, const float* b1cat,
                     const float* w2cat, const float* b2cat, const float* res, int res_p, const int* ks_host, int nb, int identity,
                     int B, int H, int W, int C, int dtype, cudaStream_t st) {
    if (nb < 1 || nb > S16_MAXB || B > 65535) return -1;
    if ((long long)H * W * C >= (1LL << 30)) return -1;                  // 32-bit offsets inside an image
    S16Params p{};
    p.B = B; p.H = H; p.W = W; p.C = C; p.nb = nb; p.identity = identity; p.res_p = res ? res_p : 0;
    for (int i = 0; i < nb; ++i) {
        if (ks_host[i] != 7 && ks_host[i] != 11 && ks_host[i] != 21) return -1;
        p.ks[i] = ks_host[i];
    }
    const int cb = dtype == CNA_BF16 ? 16 : 8;
    if (C % cb != 0) return -1;
    p.strips = ceil_div(W, S16_WT);
    p.steps = (H > 6 ? ceil_div(H - 6, 16) : 0) + 1;                    // last step's output block [16s-10, 16s+6) covers row H-1
    const size_t smem = s16_smem_bytes(p, cb);
    DevInfo di;
    if (dev_info(&di)) return 1;
    if (smem > (size_t)di.max_smem) return -1;
    dim3 grid(p.strips * (C / cb), B);
    if (dtype == CNA_BF16) {
        static std::atomic<unsigned long long> done{0};
        CNA_CUDA(ensure_dyn_smem(msca_sd16_kernel<__nv_bfloat16>, di.max_smem, done));
        msca_sd16_kernel<__nv_bfloat16><<<grid, S16_THREADS, smem, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, w0, b0, w1cat, b1cat,
                                                                        w2cat, b2cat, res, p);
    } else {
        static std::atomic<unsigned long long> done{0};
        CNA_CUDA(ensure_dyn_smem(msca_sd16_kernel<float>, di.max_smem, done));
        msca_sd16_kernel<float><<<grid, S16_THREADS, smem, st>>>((const float*)x, (float*)y, w0, b0, w1cat, b1cat, w2cat, b2cat, res, p);
    }
    CNA_LAUNCH_OK();
    return 0;
}

}  // namespace cna
