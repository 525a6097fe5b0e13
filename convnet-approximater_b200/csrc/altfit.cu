// Device-side alternating fit of scheme-1 basis filters with a nuclear-norm penalty — the role cvxpy plays in the reference
// (src/approx/core/low_rank_exp.py:69-180, Jaderberg et al. 2014 eq. (3)):
//
//     minimise over  A (R x M "weights"),  B (M x d^2 "bases")
//         f(A, B) = sum_r || W_r - A_r B ||_2   +   lambda * sum_m || reshape(B_m, d, d) ||_*          W: (R = N*C) x d^2
//
// alternately in B (A fixed) and in A (B fixed), lambda swept over an increasing list with warm starts, each lambda iterated
// until |f - f_prev| < epsilon or max_iter.  The reference hands both convex sub-problems to a generic conic solver on the CPU
// (minutes for one AlexNet layer).  Here:
//   * A-step: the objective separates over rows and ||w - a B||_2 has the minimiser of its square: A = W B^T (B B^T)^+  — one tiny
//     pseudo-inverse (Jacobi eigen-decomposition of the M x M Gram matrix, one warp) and one row-parallel product;
//   * B-step: majorise-minimise on the un-squared row norms (sum_r ||e_r|| <= sum_r ||e_r||^2 / (2 u_r) + u_r / 2 at u_r = ||e_r||),
//     which leaves a weighted least-squares term  1/2 tr(B^T G B) - tr(B^T H)  with  G = A^T D A,  H = A^T D W,  D = diag(1 / u_r)
//     — two tiny matrices reduced over all R rows by one pass — plus the nuclear norms: FISTA with singular-value soft-thresholding
//     of the M d x d bases (one warp per basis, Jacobi SVD in shared memory), all inside ONE single-CTA kernel per MM round.
// Everything that touches the R rows is a grid-wide kernel; everything M x M / M x d^2 is a one-CTA kernel; the host only reads one
// double per alternation for the stopping rule.  fp64 for the small algebra, fp32 for W / A as in the reference's write-back.
//
// cvxpy itself cannot run here and its solutions are not bit-reproducible, so parity for this entry point is "unpinned": the tests
// check the monotone decrease of f, the optimality of both block updates against an fp64 numpy restatement of the same objective,
// and lambda = 0 against the SVD subspace.
#include "common.cuh"

#include <vector>

namespace cna {

constexpr int kAfMaxM = 16;     // bases per layer
constexpr int kAfMaxD = 8;      // filter size (d x d), q = d*d <= 64
constexpr int kAfMaxQ = kAfMaxD * kAfMaxD;

__device__ __forceinline__ double af_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp: cyclic Jacobi on a symmetric n x n matrix S (row-major, destroyed): on exit diag(S) = eigenvalues, V = eigenvectors (columns).
__device__ void af_jacobi_sym(double* S, double* V, int n, int lane) {
    for (int i = lane; i < n * n; i += 32) V[i] = (i / n == i % n) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; ++sweep) {
        double off = 0;
        for (int i = lane; i < n * n; i += 32) if (i / n != i % n) off += S[i] * S[i];
        off = af_warp_sum(off);
        double diag = 0;
        for (int i = lane; i < n; i += 32) diag += S[i * n + i] * S[i * n + i];
        diag = af_warp_sum(diag);
        if (off <= 1e-30 * (diag + 1e-300)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = S[p * n + q];
                if (fabs(apq) < 1e-300) continue;
                const double app = S[p * n + p], aqq = S[q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                __syncwarp();
                for (int k = lane; k < n; k += 32) {            // columns p, q
                    const double kp = S[k * n + p], kq = S[k * n + q];
                    S[k * n + p] = c * kp - s * kq;
                    S[k * n + q] = s * kp + c * kq;
                }
                __syncwarp();
                for (int k = lane; k < n; k += 32) {            // rows p, q
                    const double pk = S[p * n + k], qk = S[q * n + k];
                    S[p * n + k] = c * pk - s * qk;
                    S[q * n + k] = s * pk + c * qk;
                }
                for (int k = lane; k < n; k += 32) {
                    const double vp = V[k * n + p], vq = V[k * n + q];
                    V[k * n + p] = c * vp - s * vq;
                    V[k * n + q] = s * vp + c * vq;
                }
                __syncwarp();
            }
    }
}

// One warp: one-sided Jacobi SVD of a d x d matrix Z (row-major, overwritten by U*diag(s)), V = right singular vectors (columns).
__device__ void af_jacobi_svd(double* Z, double* V, int d, int lane) {
    for (int i = lane; i < d * d; i += 32) V[i] = (i / d == i % d) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 40; ++sweep) {
        int rotated = 0;
        for (int p = 0; p < d - 1; ++p)
            for (int q = p + 1; q < d; ++q) {
                double a = 0, b = 0, g = 0;
                for (int i = lane; i < d; i += 32) { const double zp = Z[i * d + p], zq = Z[i * d + q]; a += zp * zp; b += zq * zq; g += zp * zq; }
                a = af_warp_sum(a); b = af_warp_sum(b); g = af_warp_sum(g);
                if (g == 0.0 || fabs(g) <= 1e-15 * sqrt(a * b)) continue;
                rotated = 1;
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int i = lane; i < d; i += 32) {
                    const double zp = Z[i * d + p], zq = Z[i * d + q];
                    Z[i * d + p] = c * zp - s * zq; Z[i * d + q] = s * zp + c * zq;
                    const double vp = V[i * d + p], vq = V[i * d + q];
                    V[i * d + p] = c * vp - s * vq; V[i * d + q] = s * vp + c * vq;
                }
                __syncwarp();
            }
        if (!rotated) break;
    }
}

// Pass over the rows: out[0] += sum_r ||W_r - A_r B||;  with want_gh also G (M*M) at out[1..] and H (M*q) behind it, weighted by
// 1 / max(||e_r||, delta).  B (M x q, fp64) is staged in shared memory; G/H are reduced per warp, then per block, then atomically.
__global__ void __launch_bounds__(256) af_rows_kernel(const float* __restrict__ W, const float* __restrict__ A, const double* __restrict__ Bm,
                                                     int R, int M, int q, double delta, double* __restrict__ out, int want_gh) {
    extern __shared__ double af_sm[];
    double* sB = af_sm;                       // M*q
    double* sAcc = sB + M * q;                // 1 + M*M + M*q
    const int nacc = 1 + (want_gh ? M * M + M * q : 0);
    for (int i = threadIdx.x; i < M * q; i += 256) sB[i] = Bm[i];
    for (int i = threadIdx.x; i < nacc; i += 256) sAcc[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int rows_per_block = (R + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * rows_per_block, r1 = min(R, r0 + rows_per_block);
    for (int rb = r0 + (threadIdx.x >> 5) * 32; rb < r1; rb += 256) {      // a warp takes 32 consecutive rows
        const int r = rb + lane;
        const bool ok = r < r1;
        float a[kAfMaxM];
        float e[kAfMaxQ];
#pragma unroll
        for (int m = 0; m < kAfMaxM; ++m) a[m] = (ok && m < M) ? A[(size_t)r * M + m] : 0.f;
        double n2 = 0;
        for (int j = 0; j < q; ++j) {
            double p = 0;
#pragma unroll
            for (int m = 0; m < kAfMaxM; ++m) if (m < M) p += (double)a[m] * sB[m * q + j];
            const float w = ok ? W[(size_t)r * q + j] : 0.f;
            e[j] = w;                                     // keep W_r for H
            const double d = (double)w - p;
            n2 += d * d;
        }
        const double u = sqrt(n2);
        double tot = af_warp_sum(ok ? u : 0.0);
        if (lane == 0) atomicAdd(&sAcc[0], tot);
        if (want_gh) {
            const double wgt = ok ? 1.0 / fmax(u, delta) : 0.0;
            for (int i = 0; i < M; ++i) {
                for (int j = 0; j < M; ++j) {
                    const double v = af_warp_sum(wgt * (double)a[i] * (double)a[j]);
                    if (lane == 0) atomicAdd(&sAcc[1 + i * M + j], v);
                }
                for (int j = 0; j < q; ++j) {
                    const double v = af_warp_sum(wgt * (double)a[i] * (double)e[j]);
                    if (lane == 0) atomicAdd(&sAcc[1 + M * M + i * q + j], v);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nacc; i += 256) atomicAdd(out + i, sAcc[i]);
}

// ONE CTA, one warp per basis.  FISTA on  1/2 tr(B^T G B) - tr(B^T H) + lambda sum_m ||B_m||_*  from the current B.
// Writes B back and nuc_out[0] = sum of the nuclear norms of the result.
__global__ void __launch_bounds__(32 * kAfMaxM) af_bases_kernel(const double* __restrict__ GH, double* __restrict__ Bm, int M, int d,
                                                               double lmda, int iters, double* __restrict__ nuc_out) {
    const int q = d * d;
    extern __shared__ double bs[];
    double* G = bs;                 // M*M
    double* H = G + M * M;          // M*q
    double* Bc = H + M * q;         // current iterate
    double* Y = Bc + M * q;         // extrapolated point
    double* Z = Y + M * q;          // per-basis SVD work (M * q)
    double* V = Z + M * q;          // per-basis V (M * q)
    double* S = V + M * q;          // M*M copy of G for the eigen-solve
    double* Vg = S + M * M;         // M*M
    __shared__ double sL, sNuc[kAfMaxM];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nthr = blockDim.x;
    for (int i = threadIdx.x; i < M * M; i += nthr) { G[i] = GH[i]; S[i] = GH[i]; }
    for (int i = threadIdx.x; i < M * q; i += nthr) { H[i] = GH[M * M + i]; Bc[i] = Bm[i]; Y[i] = Bm[i]; }
    __syncthreads();
    if (warp == 0) {
        af_jacobi_sym(S, Vg, M, lane);
        double mx = 0;
        for (int i = lane; i < M; i += 32) mx = fmax(mx, S[i * M + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) sL = fmax(mx, 1e-300);
    }
    __syncthreads();
    const double L = sL, thr = lmda / L;
    double t = 1.0;
    for (int it = 0; it <= iters; ++it) {
        const bool last = it == iters;                 // last round: only the nuclear norms of the result
        if (warp < M) {
            double* z = Z + warp * q;
            double* v = V + warp * q;
            if (!last) {
                for (int j = lane; j < q; j += 32) {   // gradient step on basis `warp`
                    double g = -H[warp * q + j];
                    for (int k = 0; k < M; ++k) g += G[warp * M + k] * Y[k * q + j];
                    z[j] = Y[warp * q + j] - g / L;
                }
            } else {
                for (int j = lane; j < q; j += 32) z[j] = Bc[warp * q + j];
            }
            __syncwarp();
            af_jacobi_svd(z, v, d, lane);
            if (last) {                                      // nuclear norm of the result = sum of the column norms of U*diag(s)
                double nuc = 0;
                for (int c = 0; c < d; ++c) {
                    double a = 0;
                    for (int i = lane; i < d; i += 32) a += z[i * d + c] * z[i * d + c];
                    nuc += sqrt(af_warp_sum(a));
                }
                if (lane == 0) sNuc[warp] = nuc;
            }
        }
        if (last) break;
        __syncthreads();
        // second pass: singular-value soft-thresholding, reconstruction and the momentum step (after the barrier: no warp may
        // overwrite Y while another still reads it for its gradient)
        const double tn = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0, mom = (t - 1.0) / tn;
        if (warp < M) {
            const double* z = Z + warp * q;
            const double* v = V + warp * q;
            double sv[kAfMaxD];
            for (int c = 0; c < d; ++c) {
                double a = 0;
                for (int i = lane; i < d; i += 32) a += z[i * d + c] * z[i * d + c];
                sv[c] = sqrt(af_warp_sum(a));
            }
            for (int j = lane; j < q; j += 32) {
                const int i = j / d, k = j % d;
                double acc = 0;
                for (int c = 0; c < d; ++c) {
                    const double sc = sv[c] > thr ? (sv[c] - thr) / sv[c] : 0.0;
                    acc += sc * z[i * d + c] * v[k * d + c];
                }
                const double old = Bc[warp * q + j];
                Bc[warp * q + j] = acc;
                Y[warp * q + j] = acc + mom * (acc - old);
            }
        }
        t = tn;
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < M * q; i += nthr) Bm[i] = Bc[i];
    if (threadIdx.x == 0) {
        double s = 0;
        for (int m = 0; m < M; ++m) s += sNuc[m];
        nuc_out[0] = s;
    }
}

// One warp: P (q x M) = B^T (B B^T)^+ with eigenvalues below 1e-12 * max treated as zero.
__global__ void __launch_bounds__(32) af_pinv_kernel(const double* __restrict__ Bm, double* __restrict__ P, int M, int q) {
    __shared__ double S[kAfMaxM * kAfMaxM], V[kAfMaxM * kAfMaxM], Pi[kAfMaxM * kAfMaxM];
    const int lane = threadIdx.x;
    for (int idx = lane; idx < M * M; idx += 32) {
        const int i = idx / M, j = idx % M;
        double a = 0;
        for (int k = 0; k < q; ++k) a += Bm[i * q + k] * Bm[j * q + k];
        S[idx] = a;
    }
    __syncwarp();
    af_jacobi_sym(S, V, M, lane);
    double mx = 0;
    for (int i = 0; i < M; ++i) mx = fmax(mx, S[i * M + i]);
    for (int idx = lane; idx < M * M; idx += 32) {
        const int i = idx / M, j = idx % M;
        double a = 0;
        for (int k = 0; k < M; ++k) {
            const double ev = S[k * M + k];
            if (ev > 1e-12 * mx) a += V[i * M + k] * V[j * M + k] / ev;
        }
        Pi[idx] = a;
    }
    __syncwarp();
    for (int idx = lane; idx < q * M; idx += 32) {
        const int k = idx / M, j = idx % M;
        double a = 0;
        for (int i = 0; i < M; ++i) a += Bm[i * q + k] * Pi[i * M + j];
        P[idx] = a;
    }
}

__global__ void __launch_bounds__(256) af_weights_kernel(const float* __restrict__ W, const double* __restrict__ P, float* __restrict__ A,
                                                        int R, int M, int q) {
    extern __shared__ double sP[];
    for (int i = threadIdx.x; i < q * M; i += 256) sP[i] = P[i];
    __syncthreads();
    for (long long idx = blockIdx.x * 256LL + threadIdx.x; idx < (long long)R * M; idx += (long long)gridDim.x * 256) {
        const int m = (int)(idx % M);
        const long long r = idx / M;
        double a = 0;
        for (int k = 0; k < q; ++k) a += (double)W[r * q + k] * sP[k * M + m];
        A[idx] = (float)a;
    }
}

__global__ void af_f2d_kernel(const float* __restrict__ s, double* __restrict__ d, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = (double)s[i];
}
__global__ void af_d2f_kernel(const double* __restrict__ s, float* __restrict__ d, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) d[i] = (float)s[i];
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_lowrank1_alt_fit(const float* W, int R, int d, int M, float* weights, float* bases, const double* lmda_list_host, int n_lmda,
                         int max_iter, double epsilon, int mm_rounds, int fista_iters, double* obj_trace_host, int trace_cap,
                         int* trace_len_host, void* stream) {
    CNA_CHECK(W && weights && bases && lmda_list_host, "null pointer");
    CNA_CHECK(R > 0 && d >= 1 && d <= kAfMaxD && M >= 1 && M <= kAfMaxM, "need d <= 8 and num_bases <= 16");
    CNA_CHECK(n_lmda >= 1 && max_iter >= 1 && mm_rounds >= 1 && fista_iters >= 1, "bad iteration counts");
    const int q = d * d;
    cudaStream_t st = as_stream(stream);
    DevInfo di;
    if (dev_info(&di)) return 1;
    AsyncScratch scr(st);
    const size_t nacc = 1 + (size_t)M * M + (size_t)M * q;
    // Bm (M*q) | P (q*M) | acc (nacc) | nuc (1)
    CNA_CUDA(scr.alloc(((size_t)2 * M * q + nacc + 1) * sizeof(double)));
    double* Bm = static_cast<double*>(scr.ptr);
    double* P = Bm + M * q;
    double* acc = P + q * M;
    double* nuc = acc + nacc;
    af_f2d_kernel<<<1, 256, 0, st>>>(bases, Bm, M * q);
    CNA_LAUNCH_OK();
    const int grid_rows = std::min(di.num_sms * 2, std::max(1, ceil_div(R, 256)));
    const size_t smem_rows = ((size_t)M * q + nacc) * sizeof(double);
    const size_t smem_bases = ((size_t)2 * M * M + (size_t)5 * M * q + M * M) * sizeof(double) + 1024;
    static std::atomic<unsigned long long> done1{0}, done2{0};
    CNA_CUDA(ensure_dyn_smem(af_rows_kernel, 96 * 1024, done1));
    CNA_CUDA(ensure_dyn_smem(af_bases_kernel, 96 * 1024, done2));
    CNA_CHECK(smem_rows <= 96 * 1024 && smem_bases <= 96 * 1024, "shared-memory plan does not fit");
    const int grid_w = (int)std::min<long long>(ceil_div_ll((long long)R * M, 256), di.num_sms * 8LL);
    double h_err = 0, h_nuc = 0, last = 0;
    int ntrace = 0;
    auto objective = [&](double lmda, bool with_nuc_from_kernel, double* out) -> int {
        CNA_CUDA(cudaMemsetAsync(acc, 0, sizeof(double), st));
        af_rows_kernel<<<grid_rows, 256, smem_rows, st>>>(W, weights, Bm, R, M, q, 0.0, acc, 0);
        CNA_LAUNCH_OK();
        CNA_CUDA(cudaMemcpyAsync(&h_err, acc, sizeof(double), cudaMemcpyDeviceToHost, st));
        if (with_nuc_from_kernel) CNA_CUDA(cudaMemcpyAsync(&h_nuc, nuc, sizeof(double), cudaMemcpyDeviceToHost, st));
        CNA_CUDA(cudaStreamSynchronize(st));
        *out = h_err + lmda * h_nuc;
        return 0;
    };
    for (int li = 0; li < n_lmda; ++li) {
        const double lmda = lmda_list_host[li];
        for (int it = 1; it <= max_iter; ++it) {
            // ---- B-step: majorise-minimise rounds, each one pass over the rows + one single-CTA FISTA solve
            for (int mm = 0; mm < mm_rounds; ++mm) {
                CNA_CUDA(cudaMemsetAsync(acc, 0, nacc * sizeof(double), st));
                // smoothing floor of the row norms: relative to the mean residual of the previous pass (1e-9 of it; absolute 1e-30 at start)
                const double delta = std::max(1e-30, 1e-9 * (h_err / R));
                af_rows_kernel<<<grid_rows, 256, smem_rows, st>>>(W, weights, Bm, R, M, q, delta, acc, 1);
                CNA_LAUNCH_OK();
                af_bases_kernel<<<1, 32 * M, smem_bases, st>>>(acc + 1, Bm, M, d, lmda, fista_iters, nuc);
                CNA_LAUNCH_OK();
            }
            double f1 = 0;
            if (objective(lmda, true, &f1)) return 1;
            if (obj_trace_host && ntrace < trace_cap) obj_trace_host[ntrace++] = f1;
            // ---- A-step: closed form
            af_pinv_kernel<<<1, 32, 0, st>>>(Bm, P, M, q);
            CNA_LAUNCH_OK();
            af_weights_kernel<<<grid_w, 256, (size_t)q * M * sizeof(double), st>>>(W, P, weights, R, M, q);
            CNA_LAUNCH_OK();
            double f2 = 0;
            if (objective(lmda, false, &f2)) return 1;
            if (obj_trace_host && ntrace < trace_cap) obj_trace_host[ntrace++] = f2;
            if (fabs(last - f2) < epsilon) { last = f2; break; }
            last = f2;
        }
    }
    af_d2f_kernel<<<1, 256, 0, st>>>(Bm, bases, M * q);
    CNA_LAUNCH_OK();
    CNA_CUDA(cudaStreamSynchronize(st));
    if (trace_len_host) *trace_len_host = ntrace;
    return 0;
}

}  // extern "C"
