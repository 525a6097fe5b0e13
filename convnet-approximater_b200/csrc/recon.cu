// Reconstruction step of the layer-wise fit: output-L2 loss and its gradient, weight gradients of the
// dense tap convolutions (scheme-2 v_conv / h_conv, 1x1 convs), bias gradient and a fused AdamW update.
//
//   reference: src/approx/hooks/finetune.py:80-82   per-sample  || Y_new - Y_old ||_2   (NOT squared)
//              src/approx/hooks/finetune.py:315-331 loss = l2_w * mean_b(sum_layers norm / L); backward; optimizer.step
// The gradient w.r.t. the layer input of a tap convolution is itself a tap convolution with flipped taps and
// transposed weights, so it reuses cna_conv_taps_forward on weights packed by cna_pack_taps_weight_dgrad.
#include "common.cuh"

namespace cna {

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sumsq[b] += sum over a slab of sample b of (y_new - y_old)^2      (grid: (slabs, B))
template <typename T>
__global__ void __launch_bounds__(256) l2_sumsq_kernel(const T* __restrict__ yn, const T* __restrict__ yo,
                                                      float* __restrict__ sumsq, long long per_sample) {
    const int b = blockIdx.y;
    const T* a = yn + (size_t)b * per_sample;
    const T* c = yo + (size_t)b * per_sample;
    float s = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x) {
        float d = to_f32(ldg(a + i)) - to_f32(ldg(c + i));
        s = fmaf(d, d, s);
    }
    __shared__ float red[8];
    s = warp_sum_f(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0) atomicAdd(sumsq + b, s);
    }
}

// norms[b] = sqrt(sumsq[b]);  g = (y_new - y_old) * scale / norm_b   (0 where the norm is 0, like torch.norm's subgradient)
template <typename T>
__global__ void __launch_bounds__(256) l2_grad_kernel(const T* __restrict__ yn, const T* __restrict__ yo,
                                                     const float* __restrict__ sumsq, float* __restrict__ norms,
                                                     T* __restrict__ g, long long per_sample, float scale) {
    const int b = blockIdx.y;
    const float nrm = sqrtf(sumsq[b]);
    if (blockIdx.x == 0 && threadIdx.x == 0) norms[b] = nrm;
    const float f = nrm > 0.f ? scale / nrm : 0.f;
    const size_t off = (size_t)b * per_sample;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x)
        g[off + i] = from_f32<T>((to_f32(ldg(yn + off + i)) - to_f32(ldg(yo + off + i))) * f);
}

// Backward of the per-sample norm for autograd callers: g[b] = (y_new[b] - y_old[b]) * upstream[b] / norms[b]   (0 where norm = 0)
template <typename T>
__global__ void __launch_bounds__(256) l2_norm_backward_kernel(const T* __restrict__ yn, const T* __restrict__ yo,
                                                              const float* __restrict__ norms, const float* __restrict__ upstream,
                                                              T* __restrict__ g, long long per_sample) {
    const int b = blockIdx.y;
    const float nrm = norms ? norms[b] : 1.f;
    const float f = nrm > 0.f ? upstream[b] / nrm : 0.f;
    const size_t off = (size_t)b * per_sample;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample; i += (long long)gridDim.x * blockDim.x)
        g[off + i] = from_f32<T>((to_f32(ldg(yn + off + i)) - (yo ? to_f32(ldg(yo + off + i)) : 0.f)) * f);
}

// ---- bf16, 16-byte vectorised variants (per_sample % 8 == 0): the reconstruction step's loss kernels move 3 passes over y, so
// they are plain HBM streams and want 128-bit accesses; the gradient kernel also accumulates the BIAS gradient
// (db[n] = sum_p g[p][n]) when the grid stride is a multiple of N, so no separate pass over g is needed. ----
__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] = __uint_as_float(u[i] << 16); f[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}

__global__ void __launch_bounds__(256) l2_sumsq_v8_kernel(const uint4* __restrict__ yn, const uint4* __restrict__ yo,
                                                         float* __restrict__ sumsq, long long per_sample8) {
    const int b = blockIdx.y;
    const uint4* a = yn + (size_t)b * per_sample8;
    const uint4* c = yo + (size_t)b * per_sample8;
    float s = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample8; i += (long long)gridDim.x * blockDim.x) {
        float fa[8], fc[8];
        bf16x8_to_f32(__ldg(a + i), fa);
        bf16x8_to_f32(__ldg(c + i), fc);
#pragma unroll
        for (int k = 0; k < 8; ++k) { const float d = fa[k] - fc[k]; s = fmaf(d, d, s); }
    }
    __shared__ float red[8];
    s = warp_sum_f(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        s = red[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
        if (threadIdx.x == 0) atomicAdd(sumsq + b, s);
    }
}

// N8 = N / 8 (vectors per pixel); dbias != NULL requires (gridDim.x * 256) % N8 == 0 so a thread always sees the same 8 channels
__global__ void __launch_bounds__(256) l2_grad_v8_kernel(const uint4* __restrict__ yn, const uint4* __restrict__ yo,
                                                        const float* __restrict__ sumsq, float* __restrict__ norms,
                                                        uint4* __restrict__ g, long long per_sample8, float scale,
                                                        float* __restrict__ dbias, int N8) {
    extern __shared__ float sdb[];                 // [N] when dbias
    const int b = blockIdx.y;
    const float nrm = sqrtf(sumsq[b]);
    if (blockIdx.x == 0 && threadIdx.x == 0) norms[b] = nrm;
    const float f = nrm > 0.f ? scale / nrm : 0.f;
    const size_t off = (size_t)b * per_sample8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (dbias) {
        for (int i = threadIdx.x; i < N8 * 8; i += 256) sdb[i] = 0.f;
        __syncthreads();
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_sample8; i += (long long)gridDim.x * blockDim.x) {
        float fa[8], fc[8];
        bf16x8_to_f32(__ldg(yn + off + i), fa);
        bf16x8_to_f32(__ldg(yo + off + i), fc);
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            __nv_bfloat162 t = __floats2bfloat162_rn((fa[2 * k] - fc[2 * k]) * f, (fa[2 * k + 1] - fc[2 * k + 1]) * f);
            o[k] = *reinterpret_cast<uint32_t*>(&t);
            acc[2 * k] += __uint_as_float(o[k] << 16);            // the bias gradient sums the ROUNDED g, like a pass over g would
            acc[2 * k + 1] += __uint_as_float(o[k] & 0xffff0000u);
        }
        g[off + i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (dbias) {
        const int ch = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % N8) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(sdb + ch + k, acc[k]);
        __syncthreads();
        for (int i = threadIdx.x; i < N8 * 8; i += 256) atomicAdd(dbias + i, sdb[i]);
    }
}

// Weight gradient of a stride-1 tap convolution:  dw[o][i][t] += sum_p x[p shifted by tap t][i] * g[p][o]
// One CTA: a 32(i) x 64(o) block of one tap, over a slab of output pixels; fp32 atomics into dw.
constexpr int WG_I = 32, WG_O = 64, WG_P = 32;
template <typename T>
__global__ void __launch_bounds__(256)
conv_taps_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ g, float* __restrict__ dw,
                       int B, int H, int W, int C, int N, int taps, int axis, int pad, int pix_per_cta) {
    __shared__ float Xs[WG_P][WG_I + 1];
    __shared__ float Gs[WG_P][WG_O + 1];
    const int t = blockIdx.z % taps, slab = blockIdx.z / taps;
    const int i0 = blockIdx.x * WG_I, o0 = blockIdx.y * WG_O;
    const long long P = (long long)B * H * W;
    const long long p_beg = (long long)slab * pix_per_cta, p_end = min(P, p_beg + pix_per_cta);
    const int tid = threadIdx.x;
    const int ti = tid & 31, to = tid >> 5;     // thread computes i = ti, o = to*8 .. to*8+7
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (long long p0 = p_beg; p0 < p_end; p0 += WG_P) {
        // stage x (shifted) and g for 32 pixels
        for (int e = tid; e < WG_P * WG_I; e += 256) {
            int pp = e / WG_I, ii = e % WG_I;
            long long p = p0 + pp;
            float v = 0.f;
            if (p < p_end && i0 + ii < C) {
                int w = (int)(p % W), h = (int)((p / W) % H);
                long long bimg = p / ((long long)W * H);
                int hs = h + (axis == 0 ? t - pad : 0), ws = w + (axis == 1 ? t - pad : 0);
                if (hs >= 0 && hs < H && ws >= 0 && ws < W)
                    v = to_f32(ldg(x + ((bimg * H + hs) * W + ws) * C + i0 + ii));
            }
            Xs[pp][ii] = v;
        }
        for (int e = tid; e < WG_P * WG_O; e += 256) {
            int pp = e / WG_O, oo = e % WG_O;
            long long p = p0 + pp;
            Gs[pp][oo] = (p < p_end && o0 + oo < N) ? to_f32(ldg(g + p * N + o0 + oo)) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int pp = 0; pp < WG_P; ++pp) {
            float xv = Xs[pp][ti];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(xv, Gs[pp][to * 8 + k], acc[k]);
        }
        __syncthreads();
    }
    if (i0 + ti < C) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            int o = o0 + to * 8 + k;
            if (o < N) atomicAdd(dw + ((size_t)o * C + i0 + ti) * taps + t, acc[k]);
        }
    }
}

// db[n] += sum_p g[p][n]
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const T* __restrict__ g, float* __restrict__ db, long long P, int N,
                                                       int pix_per_cta) {
    const long long p_beg = (long long)blockIdx.x * pix_per_cta, p_end = min(P, p_beg + pix_per_cta);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float s = 0.f;
        for (long long p = p_beg; p < p_end; ++p) s += to_f32(ldg(g + p * N + n));
        atomicAdd(db + n, s);
    }
}

__global__ void __launch_bounds__(256) adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, long long n, long long n_decay, float lr, float beta1, float beta2,
                                                       float eps, float wd, const int* __restrict__ step_dev, float gscale) {
    const float step = (float)(*step_dev + 1);
    const float bc1 = 1.f - powf(beta1, step), bc2 = 1.f - powf(beta2, step);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float pi = p[i] * (1.f - lr * (i < n_decay ? wd : 0.f));
        float mi = beta1 * m[i] + (1.f - beta1) * gi;
        float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] = pi - (lr / bc1) * mi / denom;
    }
}
__global__ void bump_step_kernel(int* step_dev) { *step_dev += 1; }

// w[o][i][t] fp32 -> packed[taps-1-t][o_pad rows = input of the transposed conv][i_pad cols = its output]
template <typename T>
__global__ void pack_taps_dgrad_kernel(const float* __restrict__ w, T* __restrict__ out, int O, int I, int taps) {
    const long long total = (long long)taps * O * I;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        int i = (int)(idx % I);
        int o = (int)((idx / I) % O);
        int tt = (int)(idx / ((long long)I * O));
        out[idx] = from_f32<T>(w[((size_t)o * I + i) * taps + (taps - 1 - tt)]);
    }
}

// decoupled weight decay Adam (torch.optim.AdamW semantics), fp32 master parameters
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, long long n_decay, float lr, float beta1, float beta2,
                                                   float eps, float wd, float bc1, float bc2, float gscale) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        float pi = p[i] * (1.f - lr * (i < n_decay ? wd : 0.f));
        float mi = beta1 * m[i] + (1.f - beta1) * gi;
        float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi; v[i] = vi;
        float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] = pi - (lr / bc1) * mi / denom;
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

static int l2_recon_impl(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad, float* dbias, int N,
                         int B, long long per_sample, float scale, int dtype, void* stream, bool zero_first = true) {
    CNA_CHECK(y_new && y_old && sumsq_scratch && norms, "null pointer");
    CNA_CHECK(B > 0 && B <= 65535 && per_sample > 0, "bad dims");
    CNA_CHECK(dbias == nullptr || (grad != nullptr && N > 0 && per_sample % N == 0), "bias gradient needs the gradient tensor and N");
    cudaStream_t st = as_stream(stream);
    if (zero_first) CNA_CUDA(cudaMemsetAsync(sumsq_scratch, 0, sizeof(float) * B, st));
    const bool aligned = ((reinterpret_cast<uintptr_t>(y_new) | reinterpret_cast<uintptr_t>(y_old) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
    if (dtype == CNA_BF16 && per_sample % 8 == 0 && aligned) {
        const long long ps8 = per_sample / 8;
        // ~8 CTAs per SM in total: enough to stream at full rate, few enough that the per-CTA bias atomics stay cheap
        // and at least 8 vectors per thread, so the per-CTA reduction atomics (N per CTA for the bias gradient, one per CTA for the
        // sum of squares) do not outweigh the streaming on small per-sample sizes
        int slabs = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(ceil_div_ll(ps8, 256 * 8), 64), ceil_div_ll(148 * 8, B)));
        const int N8 = N / 8;
        // bias-gradient fusion: every thread must keep the same 8 channels across its grid-stride iterations
        bool fuse_db = dbias != nullptr && N % 8 == 0 && N8 <= 2048;
        if (fuse_db) {
            while (slabs > 1 && ((long long)slabs * 256) % N8 != 0) --slabs;
            fuse_db = ((long long)slabs * 256) % N8 == 0;
        }
        dim3 grid(slabs, B);
        l2_sumsq_v8_kernel<<<grid, 256, 0, st>>>((const uint4*)y_new, (const uint4*)y_old, sumsq_scratch, ps8);
        CNA_LAUNCH_OK();
        if (grad) {
            if (fuse_db && zero_first) CNA_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * N, st));
            l2_grad_v8_kernel<<<grid, 256, fuse_db ? sizeof(float) * N : 0, st>>>((const uint4*)y_new, (const uint4*)y_old, sumsq_scratch, norms,
                                                                                (uint4*)grad, ps8, scale, fuse_db ? dbias : nullptr, N8);
            CNA_LAUNCH_OK();
            if (dbias && !fuse_db) return cna_bias_grad(grad, dbias, (long long)B * (per_sample / N), N, dtype, stream);
        }
        return 0;
    }
    int slabs = (int)std::min<long long>(ceil_div_ll(per_sample, 256 * 8), 64);
    dim3 grid(slabs, B);
    if (dtype == CNA_F32) l2_sumsq_kernel<float><<<grid, 256, 0, st>>>((const float*)y_new, (const float*)y_old, sumsq_scratch, per_sample);
    else if (dtype == CNA_BF16) l2_sumsq_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y_new, (const __nv_bfloat16*)y_old, sumsq_scratch, per_sample);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    if (grad) {
        if (dtype == CNA_F32) l2_grad_kernel<float><<<grid, 256, 0, st>>>((const float*)y_new, (const float*)y_old, sumsq_scratch, norms, (float*)grad, per_sample, scale);
        else l2_grad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y_new, (const __nv_bfloat16*)y_old, sumsq_scratch, norms, (__nv_bfloat16*)grad, per_sample, scale);
        CNA_LAUNCH_OK();
        if (dbias) return cna_bias_grad(grad, dbias, (long long)B * (per_sample / N), N, dtype, stream);
    }
    return 0;
}

int cna_l2_recon_loss_grad(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad,
                           int B, long long per_sample, float scale, int dtype, void* stream) {
    return l2_recon_impl(y_new, y_old, sumsq_scratch, norms, grad, nullptr, 0, B, per_sample, scale, dtype, stream);
}

int cna_l2_recon_loss_grad_bias(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad, float* dbias,
                                int N, int B, long long per_sample, float scale, int dtype, void* stream) {
    CNA_CHECK(dbias != nullptr, "null pointer");
    return l2_recon_impl(y_new, y_old, sumsq_scratch, norms, grad, dbias, N, B, per_sample, scale, dtype, stream);
}

/* Same as cna_l2_recon_loss_grad_bias, but sumsq_scratch and dbias must have been zeroed by the caller (one memset for all layers of a
 * multi-layer reconstruction step instead of two per layer). */
int cna_l2_recon_loss_grad_bias_acc(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad, float* dbias,
                                    int N, int B, long long per_sample, float scale, int dtype, void* stream) {
    CNA_CHECK(dbias != nullptr, "null pointer");
    return l2_recon_impl(y_new, y_old, sumsq_scratch, norms, grad, dbias, N, B, per_sample, scale, dtype, stream, false);
}

int cna_l2_norm_backward(const void* y_new, const void* y_old, const float* norms, const float* upstream, void* grad, int B,
                         long long per_sample, int dtype, void* stream) {
    CNA_CHECK(y_new && upstream && grad, "null pointer");          // y_old == NULL: zero; norms == NULL: one (plain per-sample scaling)
    CNA_CHECK(B > 0 && B <= 65535 && per_sample > 0, "bad dims");
    dim3 grid((unsigned)std::min<long long>(ceil_div_ll(per_sample, 256 * 8), 148 * 4), B);
    cudaStream_t st = as_stream(stream);
    if (dtype == CNA_F32) l2_norm_backward_kernel<float><<<grid, 256, 0, st>>>((const float*)y_new, (const float*)y_old, norms, upstream, (float*)grad, per_sample);
    else if (dtype == CNA_BF16) l2_norm_backward_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)y_new, (const __nv_bfloat16*)y_old, norms, upstream, (__nv_bfloat16*)grad, per_sample);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_conv_taps_wgrad(const void* x, const void* g, float* dw, float* dbias, int B, int H, int W, int C, int N,
                        int taps, int axis, int pad, int dtype, void* stream) {
    CNA_CHECK(x && g && dw, "null pointer");
    CNA_CHECK(B > 0 && H > 0 && W > 0 && C > 0 && N > 0 && taps > 0 && (axis == 0 || axis == 1), "bad dims");
    CNA_CHECK(2 * pad == taps - 1, "wgrad is implemented for stride-1 'same' tap convolutions");
    cudaStream_t st = as_stream(stream);
    const long long P = (long long)B * H * W;
    CNA_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)N * C * taps, st));
    int tiles = ceil_div(C, WG_I) * ceil_div(N, WG_O) * taps;
    int want_slabs = std::max(1, (148 * 4) / tiles);
    int pix_per_cta = (int)std::max<long long>(WG_P, ceil_div_ll(ceil_div_ll(P, want_slabs), WG_P) * WG_P);
    int slabs = (int)ceil_div_ll(P, pix_per_cta);
    CNA_CHECK((long long)slabs * taps <= 65535, "too many slabs");
    dim3 grid(ceil_div(C, WG_I), ceil_div(N, WG_O), slabs * taps);
    if (dtype == CNA_F32) conv_taps_wgrad_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (const float*)g, dw, B, H, W, C, N, taps, axis, pad, pix_per_cta);
    else if (dtype == CNA_BF16) conv_taps_wgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)g, dw, B, H, W, C, N, taps, axis, pad, pix_per_cta);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    if (dbias) {
        CNA_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * N, st));
        int ppc = (int)std::max<long long>(64, ceil_div_ll(P, 148 * 4));
        int grid_b = (int)ceil_div_ll(P, ppc);
        if (dtype == CNA_F32) bias_grad_kernel<float><<<grid_b, 256, 0, st>>>((const float*)g, dbias, P, N, ppc);
        else bias_grad_kernel<__nv_bfloat16><<<grid_b, 256, 0, st>>>((const __nv_bfloat16*)g, dbias, P, N, ppc);
        CNA_LAUNCH_OK();
    }
    return 0;
}

int cna_bias_grad(const void* g, float* dbias, long long P, int N, int dtype, void* stream) {
    CNA_CHECK(g && dbias && P > 0 && N > 0, "bad arguments");
    cudaStream_t st = as_stream(stream);
    CNA_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * N, st));
    int ppc = (int)std::max<long long>(64, ceil_div_ll(P, 148 * 4));
    int grid_b = (int)ceil_div_ll(P, ppc);
    if (dtype == CNA_F32) bias_grad_kernel<float><<<grid_b, 256, 0, st>>>((const float*)g, dbias, P, N, ppc);
    else if (dtype == CNA_BF16) bias_grad_kernel<__nv_bfloat16><<<grid_b, 256, 0, st>>>((const __nv_bfloat16*)g, dbias, P, N, ppc);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_pack_taps_weight_dgrad(const float* w_oit, void* packed, int O, int I, int taps, int dtype, void* stream) {
    CNA_CHECK(w_oit && packed && O > 0 && I > 0 && taps > 0, "bad arguments");
    long long total = (long long)taps * O * I;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 8);
    if (dtype == CNA_F32) pack_taps_dgrad_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(w_oit, (float*)packed, O, I, taps);
    else if (dtype == CNA_BF16) pack_taps_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(w_oit, (__nv_bfloat16*)packed, O, I, taps);
    else CNA_FAIL("unsupported dtype");
    CNA_LAUNCH_OK();
    return 0;
}

int cna_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, long long n_decay, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    CNA_CHECK(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1 && n_decay >= 0, "bad arguments");
    float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    int grid = (int)std::min<long long>(ceil_div_ll(n, 256), 148LL * 8);
    adamw_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, n_decay, lr, beta1, beta2, eps, weight_decay,
                                                     bc1, bc2, grad_scale);
    CNA_LAUNCH_OK();
    return 0;
}

/* Graph-capturable AdamW: the step counter lives on the device (bias corrections are computed in the kernel from *step_dev + 1,
 * a second one-thread kernel then increments it), so a captured reconstruction step stays correct across replays. */
int cna_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, long long n_decay, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int* step_dev, float grad_scale, void* stream) {
    CNA_CHECK(param && grad && exp_avg && exp_avg_sq && step_dev && n > 0 && n_decay >= 0, "bad arguments");
    int grid = (int)std::min<long long>(ceil_div_ll(n, 256), 148LL * 8);
    adamw_dev_kernel<<<grid, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, n_decay, lr, beta1, beta2, eps, weight_decay,
                                                         step_dev, grad_scale);
    CNA_LAUNCH_OK();
    bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
