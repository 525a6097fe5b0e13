// Peer (NVLink / NVSwitch) memory and the fused gradient exchange + optimizer step of the layer-wise reconstruction.
//
// Reference: the data-parallel gradient all-reduce DDP performs inside `total_loss.backward()` followed by `optimizer.step()`
// (src/approx/hooks/finetune.py:153-160, 329-331).  There the collective and the optimizer are separate launches over separate
// buffers (NCCL ring/tree all-reduce of every gradient bucket, then a multi-tensor AdamW on EVERY rank over ALL parameters).
// Here both are ONE kernel over peer memory:
//
//   barrier A   every rank's factor gradients (flat fp32 buffer, all layers) are complete
//   phase 1     rank r owns slice r of the flat buffer: it pulls that slice of all `world` gradient buffers over NVLink
//               (fixed rank order -> bitwise identical sums everywhere), runs AdamW on its slice only (optimizer state is
//               sharded: exp_avg / exp_avg_sq exist only for the owned slice's range), and PUSHES the updated fp32 parameters
//               into every rank's parameter buffer (reduce-scatter + sharded optimizer + all-gather, one-shot each way:
//               2 x (world-1)/world x n x 4 bytes per rank instead of the all-reduce's n x 4 x world read by everyone)
//   barrier B   all pushes have landed (this also makes the gradient buffers reusable: every pull happened before it)
//   phase 2     every rank re-packs its (now complete) fp32 parameters into the bf16 operand layouts the tcgen05 kernels read
//               (wv [3][k_pad][c_pad], wh [3][N][k_pad], the transposed/flipped h_conv weight of the data-gradient conv) and the
//               fp32 bias vector, so the next step's forward needs no pack kernels
//
// Barriers are flag words in each rank's own buffer, written by the peers with system-scope release stores and polled locally;
// epochs only grow, so nothing is ever reset.  The step counter lives on the device: the launch is CUDA-graph capturable and a
// replayed graph stays correct.  A spin that sees no progress for ~2 s sets an error word instead of hanging the GPU.
//
// world == 1 runs the same kernel with itself as the only peer (no NVLink traffic, both barriers trivially satisfied).
#include "common.cuh"

#include <string.h>

namespace cna {

constexpr int kPeerMax = 8;
constexpr int kReconMaxLayers = 16;
constexpr int kSyncThreads = 1024;     // one block per SM (all blocks spin in the barriers, so all must be resident)
// header of a symmetric buffer (bytes): flags A [16] | flags B [16] | local block counter | error word | pad  -> 1024
constexpr int kHdrBytes = 1024;
constexpr int kOffFlagA = 0, kOffFlagB = 64, kOffCount = 128, kOffErr = 132;

struct ReconLayerDesc {
    long long off_v, off_h, off_b, end;   // element offsets into the flat buffer: [v | h | bias)
    int K, C, N, kp, cp;
    __nv_bfloat16 *wv, *wh, *wd;          // packed operand buffers (zero-initialised padding, never written here)
    float* bias;                          // fp32 bias vector read by the fused forward
};

struct ReconSyncParams {
    uint8_t* peer[kPeerMax];              // base of every rank's symmetric buffer (peer[rank] = the local one)
    int rank, world, nlayers;
    long long n;                          // floats in the flat buffer (multiple of 4)
    long long off_grad, off_param;        // byte offsets of the gradient / parameter region inside a symmetric buffer
    float* exp_avg;                       // local optimizer state, indexed like the flat buffer (only the owned slice is touched)
    float* exp_avg_sq;
    int* step_dev;                        // steps taken so far (incremented once per launch)
    float lr, beta1, beta2, eps, wd, gscale;
    ReconLayerDesc layer[kReconMaxLayers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

// thread 0 of the block polls the `world` flag words of the LOCAL buffer until all reached `epoch`
__device__ __forceinline__ void wait_flags(const uint32_t* flags, int world, uint32_t epoch, uint32_t* err) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < world; ++r) {
            while ((int32_t)(ld_acquire_sys(flags + r) - epoch) < 0) {
                if (clock64() - t0 > 4000000000LL) { atomicExch(err, 1u); break; }      // ~2 s: report instead of hanging
                __nanosleep(64);
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int find_layer(const ReconSyncParams& p, long long i) {
    int l = 0;
    while (l + 1 < p.nlayers && i >= p.layer[l].end) ++l;
    return l;
}

__global__ void __launch_bounds__(kSyncThreads)
recon_sync_kernel(const __grid_constant__ ReconSyncParams p) {
    uint8_t* me = p.peer[p.rank];
    uint32_t* flagA = reinterpret_cast<uint32_t*>(me + kOffFlagA);
    uint32_t* flagB = reinterpret_cast<uint32_t*>(me + kOffFlagB);
    uint32_t* count = reinterpret_cast<uint32_t*>(me + kOffCount);
    uint32_t* err = reinterpret_cast<uint32_t*>(me + kOffErr);
    const int step0 = *reinterpret_cast<volatile int*>(p.step_dev);
    const uint32_t epoch = (uint32_t)step0 + 1u;

    // ---- barrier A: my gradients (written by earlier kernels of this stream) are complete -> tell every peer, wait for all
    if (blockIdx.x == 0 && threadIdx.x < p.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<uint32_t*>(p.peer[threadIdx.x] + kOffFlagA) + p.rank, epoch);
    }
    wait_flags(flagA, p.world, epoch, err);

    // ---- phase 1: reduce my slice over all ranks, AdamW, push the new parameters to every rank
    const long long n4 = p.n >> 2;
    const long long per = (n4 + p.world - 1) / p.world;
    const long long lo4 = min(n4, per * p.rank), hi4 = min(n4, lo4 + per);
    const float stepf = (float)epoch;
    const float bc1 = 1.f - powf(p.beta1, stepf), bc2 = 1.f - powf(p.beta2, stepf);
    const float inv_sqrt_bc2 = rsqrtf(bc2), step_size = p.lr / bc1;
    const float4* myparam = reinterpret_cast<const float4*>(me + p.off_param);
    for (long long i4 = lo4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i4 < hi4; i4 += (long long)gridDim.x * blockDim.x) {
        float4 pv4[kPeerMax];
#pragma unroll
        for (int r = 0; r < kPeerMax; ++r)                        // all peer loads in flight before the first add (NVLink latency once)
            pv4[r] = r < p.world ? ld_volatile_f4(reinterpret_cast<const float4*>(p.peer[r] + p.off_grad) + i4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < kPeerMax; ++r) {                      // fixed rank order: every rank would compute the same bits
            if (r < p.world) { g[0] += pv4[r].x; g[1] += pv4[r].y; g[2] += pv4[r].z; g[3] += pv4[r].w; }
        }
        const float4 pv = myparam[i4];
        float4 mv = reinterpret_cast<const float4*>(p.exp_avg)[i4], vv = reinterpret_cast<const float4*>(p.exp_avg_sq)[i4];
        float pw[4] = {pv.x, pv.y, pv.z, pv.w}, m[4] = {mv.x, mv.y, mv.z, mv.w}, v2[4] = {vv.x, vv.y, vv.z, vv.w};
        const long long i = i4 << 2;
        const int l = find_layer(p, i);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // biases carry no weight decay (timm's create_optimizer_v2 default filter_bias_and_bn, reference hooks/finetune.py:188)
            int lq = l;
            while (lq + 1 < p.nlayers && i + q >= p.layer[lq].end) ++lq;
            const float wd = (i + q >= p.layer[lq].off_b) ? 0.f : p.wd;
            const float gi = g[q] * p.gscale;
            const float pi = pw[q] * (1.f - p.lr * wd);
            m[q] = p.beta1 * m[q] + (1.f - p.beta1) * gi;
            v2[q] = p.beta2 * v2[q] + (1.f - p.beta2) * gi * gi;
            const float denom = sqrtf(v2[q]) * inv_sqrt_bc2 + p.eps;
            pw[q] = pi - step_size * m[q] / denom;
        }
        reinterpret_cast<float4*>(p.exp_avg)[i4] = make_float4(m[0], m[1], m[2], m[3]);
        reinterpret_cast<float4*>(p.exp_avg_sq)[i4] = make_float4(v2[0], v2[1], v2[2], v2[3]);
        const float4 out = make_float4(pw[0], pw[1], pw[2], pw[3]);
        for (int r = 0; r < p.world; ++r) reinterpret_cast<float4*>(p.peer[r] + p.off_param)[i4] = out;
    }

    // ---- barrier B: all my pushes are visible system-wide -> tell every peer (last block of this grid), wait for all
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicInc(count, gridDim.x - 1) == gridDim.x - 1) {    // last block of this grid to arrive (the counter wraps back to 0)
            *reinterpret_cast<volatile int*>(p.step_dev) = step0 + 1;   // every block has read step0 before arriving here
            __threadfence_system();
            for (int r = 0; r < p.world; ++r) st_release_sys(reinterpret_cast<uint32_t*>(p.peer[r] + kOffFlagB) + p.rank, epoch);
        }
    }
    wait_flags(flagB, p.world, epoch, err);

    // ---- phase 2: re-pack ALL parameters (every rank needs all operand buffers) from the local fp32 copy.
    // Destination-ordered: consecutive threads write consecutive bf16 elements of a pack buffer (coalesced stores) and gather the
    // fp32 source (L2-resident, 12-byte stride); padded rows / columns of the packs were zeroed once by the caller and are skipped.
    const volatile float* param = reinterpret_cast<const volatile float*>(me + p.off_param);
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gstride = (long long)gridDim.x * blockDim.x;
    for (int l = 0; l < p.nlayers; ++l) {
        const ReconLayerDesc& L = p.layer[l];
        const long long nv = 3LL * L.K * L.C, nh = 3LL * L.N * L.K;
        for (long long d = gtid; d < nv + 2 * nh + L.N; d += gstride) {
            if (d < nv) {                                          // wv[t][k][c] <- v_conv.weight[k][c][t]
                const int c = (int)(d % L.C), k = (int)((d / L.C) % L.K), t = (int)(d / ((long long)L.C * L.K));
                L.wv[((size_t)t * L.kp + k) * L.cp + c] = __float2bfloat16_rn(param[L.off_v + ((long long)k * L.C + c) * 3 + t]);
            } else if (d < nv + nh) {                              // wh[t][n][k] <- h_conv.weight[n][k][t]
                const long long e = d - nv;
                const int k = (int)(e % L.K), nn = (int)((e / L.K) % L.N), t = (int)(e / ((long long)L.K * L.N));
                L.wh[((size_t)t * L.N + nn) * L.kp + k] = __float2bfloat16_rn(param[L.off_h + ((long long)nn * L.K + k) * 3 + t]);
            } else if (d < nv + 2 * nh) {                          // wd[t][k][n] <- h_conv.weight[n][k][2 - t]  (data-gradient conv)
                const long long e = d - nv - nh;
                const int nn = (int)(e % L.N), k = (int)((e / L.N) % L.K), t = (int)(e / ((long long)L.N * L.K));
                L.wd[((size_t)t * L.kp + k) * L.N + nn] = __float2bfloat16_rn(param[L.off_h + ((long long)nn * L.K + k) * 3 + (2 - t)]);
            } else {
                const long long e = d - nv - 2 * nh;
                L.bias[e] = param[L.off_b + e];
            }
        }
    }
}

}  // namespace cna

using namespace cna;

extern "C" {

int cna_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_host) {
    CNA_CHECK(bytes > 0 && ptr_out && handle64_host, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    CNA_CUDA(cudaMalloc(&p, (size_t)bytes));
    CNA_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        CNA_FAIL(std::string("cudaIpcGetMemHandle -> ") + cudaGetErrorString(e));
    }
    memcpy(handle64_host, &h, 64);
    *ptr_out = p;
    return 0;
}

int cna_peer_open(const unsigned char* handle64_host, void** ptr_out) {
    CNA_CHECK(handle64_host && ptr_out, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void* p = nullptr;
    CNA_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return 0;
}

int cna_peer_close(void* ptr) {
    CNA_CHECK(ptr != nullptr, "null pointer");
    CNA_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

int cna_peer_free(void* ptr) {
    CNA_CHECK(ptr != nullptr, "null pointer");
    CNA_CUDA(cudaFree(ptr));
    return 0;
}

int cna_recon_sync_header_bytes(void) { return kHdrBytes; }

int cna_recon_sync_step(void* const* peers_host, int rank, int world, long long n, long long off_grad, long long off_param,
                        float* exp_avg, float* exp_avg_sq, int* step_dev, const long long* layer_offsets_host,
                        const int* layer_dims_host, void* const* layer_ptrs_host, int nlayers, float lr, float beta1, float beta2,
                        float eps, float weight_decay, float grad_scale, void* stream) {
    CNA_CHECK(peers_host && exp_avg && exp_avg_sq && step_dev && layer_offsets_host && layer_dims_host && layer_ptrs_host, "null pointer");
    CNA_CHECK(world >= 1 && world <= kPeerMax && rank >= 0 && rank < world, "bad rank / world size (at most 8 peers)");
    CNA_CHECK(nlayers >= 1 && nlayers <= kReconMaxLayers, "1..16 layers per flat buffer");
    CNA_CHECK(n > 0 && n % 4 == 0 && off_grad % 16 == 0 && off_param % 16 == 0 && off_grad >= kHdrBytes && off_param >= kHdrBytes,
              "flat buffer must be a multiple of 4 floats, regions 16-byte aligned behind the header");
    ReconSyncParams p{};
    for (int r = 0; r < world; ++r) {
        CNA_CHECK(peers_host[r] != nullptr, "null peer pointer");
        p.peer[r] = static_cast<uint8_t*>(peers_host[r]);
    }
    p.rank = rank; p.world = world; p.nlayers = nlayers; p.n = n; p.off_grad = off_grad; p.off_param = off_param;
    p.exp_avg = exp_avg; p.exp_avg_sq = exp_avg_sq; p.step_dev = step_dev;
    p.lr = lr; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps; p.wd = weight_decay; p.gscale = grad_scale;
    for (int l = 0; l < nlayers; ++l) {
        ReconLayerDesc& L = p.layer[l];
        L.off_v = layer_offsets_host[4 * l]; L.off_h = layer_offsets_host[4 * l + 1]; L.off_b = layer_offsets_host[4 * l + 2];
        L.end = layer_offsets_host[4 * l + 3];
        L.K = layer_dims_host[5 * l]; L.C = layer_dims_host[5 * l + 1]; L.N = layer_dims_host[5 * l + 2];
        L.kp = layer_dims_host[5 * l + 3]; L.cp = layer_dims_host[5 * l + 4];
        CNA_CHECK(L.off_v <= L.off_h && L.off_h <= L.off_b && L.off_b <= L.end && L.end <= n, "layer offsets out of order");
        CNA_CHECK(L.off_h - L.off_v == 3LL * L.K * L.C && L.off_b - L.off_h == 3LL * L.N * L.K && L.end - L.off_b == L.N,
                  "layer extents do not match (K,C,3,1) | (N,K,1,3) | (N)");
        CNA_CHECK(L.kp >= L.K && L.cp >= L.C, "bad padded dims");
        L.wv = static_cast<__nv_bfloat16*>(layer_ptrs_host[4 * l]); L.wh = static_cast<__nv_bfloat16*>(layer_ptrs_host[4 * l + 1]);
        L.wd = static_cast<__nv_bfloat16*>(layer_ptrs_host[4 * l + 2]); L.bias = static_cast<float*>(layer_ptrs_host[4 * l + 3]);
        CNA_CHECK(L.wv && L.wh && L.wd && L.bias, "null pack pointer");
        CNA_CHECK(l == 0 || L.off_v == p.layer[l - 1].end, "layers must be contiguous in the flat buffer");
    }
    DevInfo di;
    if (dev_info(&di)) return 1;
    // every block spins in the barriers, so all of them must be resident at once: one block per SM at most
    const int grid = (int)std::max<long long>(1, std::min<long long>(di.num_sms, ceil_div_ll(n / 4, kSyncThreads)));
    recon_sync_kernel<<<grid, kSyncThreads, 0, as_stream(stream)>>>(p);
    CNA_LAUNCH_OK();
    return 0;
}

/* error word of a symmetric buffer (host copy): non-zero if a barrier timed out in any launch since allocation */
int cna_recon_sync_error(const void* local_buffer, int* err_host) {
    CNA_CHECK(local_buffer && err_host, "null pointer");
    uint32_t v = 0;
    CNA_CUDA(cudaMemcpy(&v, static_cast<const uint8_t*>(local_buffer) + kOffErr, 4, cudaMemcpyDeviceToHost));
    *err_host = (int)v;
    return 0;
}

}  // extern "C"
