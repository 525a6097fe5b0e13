// Device-side fitting step: the small dense linear algebra behind the approximater apps.
//
//   * batched one-sided Jacobi SVD of small matrices (d x d basis filters, 21 x 21 merged MSCA kernels,
//     q x q Gram matrices), one warp per matrix, fp64 in shared memory
//       - LowRankExpConvV1.decomp            reference src/approx/layers/low_rank_conv.py:58-66
//       - MscaRep.optimize (decomp >= 1)     reference src/approx/core/msca_rep.py:154-168
//   * Gram-matrix route for the tall-skinny SVD of the (N*C) x d^2 filter matrix
//       - LowRankExpV1 SVD initialisation    reference src/approx/core/low_rank_exp.py:110-120,148-151,181-184
//   * MSCA equivalent-kernel merge (weight, bias, border residues)
//       - MscaRep.get_equivalent_kernel      reference src/approx/core/msca_rep.py:57-143
#include "common.cuh"

#include <algorithm>
#include <vector>

namespace cna {

constexpr int kSvdMax = 64;   // max rows / cols handled by the in-smem Jacobi
constexpr int kMaxBranches = 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp.  G: m x n (row-major, ld = n) working copy of A, V: n x n.  On exit the columns of G are
// mutually orthogonal (G = U * diag(S)), V holds the right singular vectors as columns.
__device__ void jacobi_one_sided(double* G, double* V, int m, int n, int lane) {
    for (int i = lane; i < n * n; i += 32) V[i] = (i / n == i % n) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 60; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < n - 1; ++p) {
            for (int q = p + 1; q < n; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = lane; i < m; i += 32) {
                    double gp = G[i * n + p], gq = G[i * n + q];
                    a += gp * gp; b += gq * gq; g += gp * gq;
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                if (fabs(g) <= 1e-15 * sqrt(a * b) || g == 0.0) continue;
                rotated = 1;
                double zeta = (b - a) / (2.0 * g);
                double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = lane; i < m; i += 32) {
                    double gp = G[i * n + p], gq = G[i * n + q];
                    G[i * n + p] = c * gp - s * gq;
                    G[i * n + q] = s * gp + c * gq;
                }
                for (int i = lane; i < n; i += 32) {
                    double vp = V[i * n + p], vq = V[i * n + q];
                    V[i * n + p] = c * vp - s * vq;
                    V[i * n + q] = s * vp + c * vq;
                }
                __syncwarp();
            }
        }
        if (!rotated) break;
    }
}

// A (batch, m, n) with m >= n -> U (batch, m, n), S (batch, n) descending, Vh (batch, n, n).
template <typename TI, typename TO>
__global__ void __launch_bounds__(32) svd_small_kernel(const TI* __restrict__ A, TO* __restrict__ U, TO* __restrict__ S,
                                                      TO* __restrict__ Vh, int m, int n) {
    extern __shared__ double sm[];
    double* G = sm;             // m*n
    double* V = G + m * n;      // n*n
    double* sv = V + n * n;     // n
    int* order = reinterpret_cast<int*>(sv + n);
    const int lane = threadIdx.x;
    const size_t bidx = blockIdx.x;
    for (int i = lane; i < m * n; i += 32) G[i] = (double)A[bidx * m * n + i];
    __syncwarp();
    jacobi_one_sided(G, V, m, n, lane);
    for (int j = 0; j < n; ++j) {
        double a = 0;
        for (int i = lane; i < m; i += 32) a += G[i * n + j] * G[i * n + j];
        a = warp_sum(a);
        if (lane == 0) { sv[j] = sqrt(a); order[j] = j; }
    }
    __syncwarp();
    if (lane == 0) {  // selection sort, n <= 64
        for (int i = 0; i < n - 1; ++i) {
            int best = i;
            for (int j = i + 1; j < n; ++j) if (sv[order[j]] > sv[order[best]]) best = j;
            int t = order[i]; order[i] = order[best]; order[best] = t;
        }
    }
    __syncwarp();
    for (int idx = lane; idx < m * n; idx += 32) {
        int i = idx / n, j = idx % n, src = order[j];
        double s = sv[src];
        U[bidx * m * n + idx] = (TO)(s > 0 ? G[i * n + src] / s : 0.0);
    }
    for (int j = lane; j < n; j += 32) S[bidx * n + j] = (TO)sv[order[j]];
    for (int idx = lane; idx < n * n; idx += 32) {
        int j = idx / n, i = idx % n;
        Vh[bidx * n * n + idx] = (TO)V[i * n + order[j]];
    }
}

// G (q x q, fp64) += W^T W over a slab of rows.  W: R x q fp32.
__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ W, double* __restrict__ G, int R, int q,
                                                  int rows_per_block) {
    extern __shared__ float tile[];   // 32 rows x q
    const int r0 = blockIdx.x * rows_per_block;
    const int r1 = min(R, r0 + rows_per_block);
    const int npair = q * q;
    // each thread owns pairs tid, tid+256, ... (at most ceil(q*q/256) <= 16 for q <= 64)
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (int rb = r0; rb < r1; rb += 32) {
        int nr = min(32, r1 - rb);
        for (int i = threadIdx.x; i < nr * q; i += 256) tile[i] = W[(size_t)rb * q + i];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            int pr = threadIdx.x + k * 256;
            if (pr < npair) {
                int i = pr / q, j = pr % q;
                double a = 0;
                for (int r = 0; r < nr; ++r) a += (double)tile[r * q + i] * (double)tile[r * q + j];
                acc[k] += a;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int pr = threadIdx.x + k * 256;
        if (pr < npair) atomicAdd(G + pr, acc[k]);
    }
}

// weights[r][m] = (W[r,:] . v_m) / sqrt(s_m) ;  bases[m][:] = sqrt(s_m) * v_m   with s_m = sqrt(lambda_m)
// Vh (q x q fp64, rows = right singular vectors), lam (q) = eigenvalues of W^T W, descending.
__global__ void __launch_bounds__(256) lowrank1_factors_kernel(const float* __restrict__ W, const double* __restrict__ Vh,
                                                              const double* __restrict__ lam, float* __restrict__ weights,
                                                              float* __restrict__ bases, int R, int q, int M) {
    const long long total = (long long)R * M;
    const int upb = min(M, q);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int mm = (int)(idx % M);
        long long r = idx / M;
        double out = 0.0;
        if (mm < upb) {
            double s = sqrt(fmax(lam[mm], 0.0));  // singular value
            if (s > 0) {
                double a = 0;
                for (int j = 0; j < q; ++j) a += (double)W[r * q + j] * Vh[mm * q + j];
                out = a / sqrt(s);
            }
        }
        weights[idx] = (float)out;
    }
    if (blockIdx.x == 0) {
        for (int i = threadIdx.x; i < M * q; i += blockDim.x) {
            int mm = i / q, j = i % q;
            double v = 0.0;
            if (mm < upb) v = sqrt(sqrt(fmax(lam[mm], 0.0))) * Vh[mm * q + j];
            bases[i] = (float)v;
        }
    }
}

// ---- one-sided Jacobi SVD of a large matrix, one launch per tournament round ---------------------------------
// Work arrays (fp64): Wk [n][m] (row j = j-th column of A), Vk [n][n] (row j = j-th column of V).
// Round `r` of the round-robin tournament pairs n/2 disjoint (p,q); each CTA orthogonalises one pair.
__global__ void __launch_bounds__(256) jacobi_round_kernel(double* __restrict__ Wk, double* __restrict__ Vk, int m, int n,
                                                          int n_even, int round, int* __restrict__ rotations) {
    // round-robin schedule on n_even players (player n_even-1 fixed)
    const int i = blockIdx.x;                 // 0 .. n_even/2 - 1
    const int mod = n_even - 1;
    int p, q;
    if (i == 0) { p = n_even - 1; q = round % mod; }
    else { p = (round + i) % mod; q = (round - i + mod) % mod; }
    if (p >= n || q >= n) return;             // padding player
    if (p > q) { int t = p; p = q; q = t; }
    double* wp = Wk + (size_t)p * m; double* wq = Wk + (size_t)q * m;
    double a = 0, b = 0, g = 0;
    for (int k = threadIdx.x; k < m; k += 256) { double x = wp[k], y = wq[k]; a += x * x; b += y * y; g += x * y; }
    __shared__ double red[3][8];
    a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; red[2][threadIdx.x >> 5] = g; }
    __syncthreads();
    a = b = g = 0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b += red[1][w]; g += red[2][w]; }
    if (fabs(g) <= 1e-14 * sqrt(a * b) || g == 0.0) return;
    if (threadIdx.x == 0) atomicAdd(rotations, 1);
    double zeta = (b - a) / (2.0 * g);
    double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
    for (int k = threadIdx.x; k < m; k += 256) { double x = wp[k], y = wq[k]; wp[k] = c * x - sn * y; wq[k] = sn * x + c * y; }
    double* vp = Vk + (size_t)p * n; double* vq = Vk + (size_t)q * n;
    for (int k = threadIdx.x; k < n; k += 256) { double x = vp[k], y = vq[k]; vp[k] = c * x - sn * y; vq[k] = sn * x + c * y; }
}

// A (m x n fp32 row-major) -> Wk [n][m] fp64 ; Vk = I
__global__ void jacobi_init_kernel(const float* __restrict__ A, double* __restrict__ Wk, double* __restrict__ Vk, int m, int n) {
    const long long tot = (long long)m * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tot; i += (long long)gridDim.x * blockDim.x) {
        int j = (int)(i / m), k = (int)(i % m);
        Wk[i] = (double)A[(size_t)k * n + j];
    }
    const long long tv = (long long)n * n;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < tv; i += (long long)gridDim.x * blockDim.x)
        Vk[i] = (i / n == i % n) ? 1.0 : 0.0;
}

__global__ void jacobi_norms_kernel(const double* __restrict__ Wk, double* __restrict__ sv, int m, int n) {
    const int j = blockIdx.x;
    double a = 0;
    for (int k = threadIdx.x; k < m; k += 256) { double x = Wk[(size_t)j * m + k]; a += x * x; }
    __shared__ double red[8];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) { double t = 0; for (int w = 0; w < 8; ++w) t += red[w]; sv[j] = sqrt(t); }
}

// gather the top-r triplets given the descending order: U (m x r), S (r), Vh (r x n), fp32
__global__ void jacobi_gather_kernel(const double* __restrict__ Wk, const double* __restrict__ Vk, const double* __restrict__ sv,
                                     const int* __restrict__ order, float* __restrict__ U, float* __restrict__ S,
                                     float* __restrict__ Vh, int m, int n, int r) {
    const int j = blockIdx.x, src = order[j];
    const double s = sv[src];
    for (int k = threadIdx.x; k < m; k += 256) U[(size_t)k * r + j] = (float)(s > 0 ? Wk[(size_t)src * m + k] / s : 0.0);
    for (int k = threadIdx.x; k < n; k += 256) Vh[(size_t)j * n + k] = (float)Vk[(size_t)src * n + k];
    if (threadIdx.x == 0) S[j] = (float)s;
}

struct EquivParams {
    int nb, C, K;
    int ks[kMaxBranches];
    int off[kMaxBranches];   // element offset of branch i inside w1cat / w2cat (C * k_i each)
};

// one block per channel
__global__ void __launch_bounds__(256) msca_equiv_kernel(const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                                        float* __restrict__ weight, float* __restrict__ bias,
                                                        float* __restrict__ res, EquivParams p) {
    const int c = blockIdx.x, K = p.K, P = K / 2;
    for (int idx = threadIdx.x; idx < K * K; idx += blockDim.x) {
        int i = idx / K, j = idx % K;
        float v = (i == K / 2 && j == K / 2) ? 1.f : 0.f;
        for (int br = 0; br < p.nb; ++br) {
            int k = p.ks[br], qo = (K - k) / 2;
            int ii = i - qo, jj = j - qo;
            if (ii >= 0 && ii < k && jj >= 0 && jj < k)
                v += w2[p.off[br] + c * k + ii] * w1[p.off[br] + c * k + jj];
        }
        weight[(size_t)c * K * K + idx] = v;
    }
    if (threadIdx.x == 0) {
        float bsum = 0.f;
        for (int br = 0; br < p.nb; ++br) {
            int k = p.ks[br];
            float s = 0.f;
            for (int i = 0; i < k; ++i) s += w2[p.off[br] + c * k + i];
            float bb1 = b1 ? b1[br * p.C + c] : 0.f, bb2 = b2 ? b2[br * p.C + c] : 0.f;
            bsum += s * bb1 + bb2;
        }
        bias[c] = bsum;
    }
    // border residues: res[0][c][.] top, res[1][c][.] bottom, P entries each
    for (int idx = threadIdx.x; idx < 2 * P; idx += blockDim.x) {
        int side = idx / P, e = idx % P;
        float v = 0.f;
        for (int br = 0; br < p.nb; ++br) {
            int k = p.ks[br], pb = k / 2;
            float bb1 = b1 ? b1[br * p.C + c] : 0.f;
            const float* w = w2 + p.off[br] + c * k;
            if (side == 0) {
                // branch residue r0[pb - i] = -(sum_{t<i} w[t]) * b1, i = 1..pb, merged at index (pb - i) (top aligned)
                if (e < pb) {
                    int i = pb - e;
                    float s = 0.f;
                    for (int t = 0; t < i; ++t) s += w[t];
                    v -= s * bb1;
                }
            } else {
                // r1[i - 1] = -(sum of last i taps) * b1, merged at index P - pb + (i - 1) (bottom aligned)
                int i = e - (P - pb) + 1;
                if (i >= 1 && i <= pb) {
                    float s = 0.f;
                    for (int t = k - i; t < k; ++t) s += w[t];
                    v -= s * bb1;
                }
            }
        }
        res[((size_t)side * p.C + c) * P + e] = v;
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_svd_small_batched(const float* A, float* U, float* S, float* Vh, int batch, int m, int n, void* stream) {
    CNA_CHECK(A && U && S && Vh && batch > 0, "bad arguments");
    CNA_CHECK(m >= n && n >= 1 && m <= kSvdMax, "need 1 <= n <= m <= 64 (transpose on the caller side otherwise)");
    size_t smem = (size_t)(m * n + n * n + n) * sizeof(double) + n * sizeof(int);
    CNA_CUDA(cudaFuncSetAttribute(svd_small_kernel<float, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    svd_small_kernel<float, float><<<batch, 32, smem, as_stream(stream)>>>(A, U, S, Vh, m, n);
    CNA_LAUNCH_OK();
    return 0;
}

int cna_lowrank1_svd_fit(const float* W, int R, int q, int M, float* weights, float* bases, void* stream) {
    CNA_CHECK(W && weights && bases && R > 0 && q > 0 && M > 0, "bad arguments");
    CNA_CHECK(q <= kSvdMax, "filter matrices with more than 64 columns (d > 8) are not supported yet");
    cudaStream_t st = as_stream(stream);
    size_t nG = (size_t)q * q;
    AsyncScratch scratch(st);                                           // freed on every return path
    CNA_CUDA(scratch.alloc((3 * nG + q) * sizeof(double)));             // G | U | Vh | lam
    double* G = static_cast<double*>(scratch.ptr);
    double *U = G + nG, *Vh = U + nG, *lam = Vh + nG;
    CNA_CUDA(cudaMemsetAsync(G, 0, nG * sizeof(double), st));
    int rows_per_block = 1024;
    gram_kernel<<<ceil_div(R, rows_per_block), 256, 32 * q * sizeof(float), st>>>(W, G, R, q, rows_per_block);
    CNA_LAUNCH_OK();
    size_t smem = (size_t)(2 * q * q + q) * sizeof(double) + q * sizeof(int);
    CNA_CUDA(cudaFuncSetAttribute(svd_small_kernel<double, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    svd_small_kernel<double, double><<<1, 32, smem, st>>>(G, U, lam, Vh, q, q);   // symmetric PSD: S = eigenvalues
    CNA_LAUNCH_OK();
    long long total = (long long)R * M;
    int grid = (int)std::min<long long>(ceil_div_ll(total, 256), 148LL * 16);
    lowrank1_factors_kernel<<<grid, 256, 0, st>>>(W, Vh, lam, weights, bases, R, q, M);
    CNA_LAUNCH_OK();
    return 0;
}

// Truncated SVD of a large fp32 matrix A (m x n row-major, n <= m): the r leading triplets.
// One-sided (Hestenes) Jacobi in fp64 with a round-robin tournament: n-1 launches per sweep, n/2 independent column
// pairs per launch; stops when a sweep makes no rotation.  Used for the scheme-2 closed-form fit, where A is the
// (C*d) x (d*N) reshaping of the filter bank (Jaderberg et al. 2014, section 3.2).
int cna_svd_jacobi_large(const float* A, int m, int n, int r, float* U, float* S, float* Vh, int max_sweeps, void* stream) {
    CNA_CHECK(A && U && S && Vh, "null pointer");
    CNA_CHECK(m >= n && n >= 1 && r >= 1 && r <= n, "need 1 <= r <= n <= m (transpose on the caller side otherwise)");
    cudaStream_t st = as_stream(stream);
    const size_t nW = (size_t)n * m, nV = (size_t)n * n;
    AsyncScratch s_w(st), s_rot(st);                                    // freed on every return path
    CNA_CUDA(s_w.alloc((nW + nV + n) * sizeof(double)));
    CNA_CUDA(s_rot.alloc(sizeof(int) * (n + 1)));
    double* Wk = static_cast<double*>(s_w.ptr);
    int* d_rot = static_cast<int*>(s_rot.ptr);
    double *Vk = Wk + nW, *sv = Vk + nV;
    int* d_order = d_rot + 1;
    jacobi_init_kernel<<<148 * 4, 256, 0, st>>>(A, Wk, Vk, m, n);
    CNA_LAUNCH_OK();
    const int n_even = n + (n & 1);
    int sweeps = 0;
    if (n > 1) {
        for (; sweeps < max_sweeps; ++sweeps) {
            CNA_CUDA(cudaMemsetAsync(d_rot, 0, sizeof(int), st));
            for (int round = 0; round < n_even - 1; ++round)
                jacobi_round_kernel<<<n_even / 2, 256, 0, st>>>(Wk, Vk, m, n, n_even, round, d_rot);
            count_launch(n_even - 1);
            int h_rot = 0;
            CNA_CUDA(cudaMemcpyAsync(&h_rot, d_rot, sizeof(int), cudaMemcpyDeviceToHost, st));
            CNA_CUDA(cudaStreamSynchronize(st));
            if (h_rot == 0) break;
        }
    }
    jacobi_norms_kernel<<<n, 256, 0, st>>>(Wk, sv, m, n);
    CNA_LAUNCH_OK();
    // order by singular value (host side: n <= a few thousand values)
    std::vector<double> h_sv(n);
    CNA_CUDA(cudaMemcpyAsync(h_sv.data(), sv, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CNA_CUDA(cudaStreamSynchronize(st));
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h_sv[a] > h_sv[b]; });
    CNA_CUDA(cudaMemcpyAsync(d_order, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
    jacobi_gather_kernel<<<r, 256, 0, st>>>(Wk, Vk, sv, d_order, U, S, Vh, m, n, r);
    CNA_LAUNCH_OK();
    CNA_CUDA(cudaStreamSynchronize(st));   // `order` (host) must outlive the copy
    return 0;
}

int cna_msca_equiv_kernel(const float* w1cat, const float* b1, const float* w2cat, const float* b2,
                          const int* ks_host, int nb, int C, float* weight, float* bias, float* res, void* stream) {
    CNA_CHECK(w1cat && w2cat && ks_host && weight && bias && res, "null pointer");
    CNA_CHECK(nb >= 1 && nb <= kMaxBranches && C > 0, "bad branch/channel count");
    EquivParams p{};
    p.nb = nb; p.C = C; p.K = 0;
    int off = 0;
    for (int i = 0; i < nb; ++i) {
        CNA_CHECK(ks_host[i] >= 1 && ks_host[i] % 2 == 1, "strip sizes must be odd");
        p.ks[i] = ks_host[i];
        p.off[i] = off;
        off += C * ks_host[i];
        p.K = std::max(p.K, ks_host[i]);
    }
    // the reference takes the LAST branch as the largest (core/msca_rep.py:123: `hw = max(w1_lst[-1].shape[2:])`)
    CNA_CHECK(p.K == ks_host[nb - 1], "k_sizes must be ascending (the reference assumes the last branch is the largest)");
    msca_equiv_kernel<<<C, 256, 0, as_stream(stream)>>>(w1cat, b1, w2cat, b2, weight, bias, res, p);
    CNA_LAUNCH_OK();
    return 0;
}

}  // extern "C"
