/*
 * cna_b200.h — C ABI of the B200 (sm_100a) backend for ConvNet-Approximater's hot path.
 *
 * The reference (ashun989/ConvNet-Approximater) is pure Python/PyTorch and has no FFI of its
 * own; the "interface each entry point replaces" is therefore the `forward()` / `optimize()`
 * body of the reference module or app named beside it (paths relative to the reference root).
 * INTEGRATION.md shows the ctypes binding a maintainer adds to those bodies.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - activations are NHWC (channels-last) contiguous: x[b][h][w][c];
 *   - `dtype` selects the activation AND packed-weight element type: CNA_F32 or CNA_BF16
 *     (bf16 kernels accumulate in fp32 and round intermediates to bf16 exactly where the
 *     reference's bf16 modules materialise a tensor); biases / residues are always fp32;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous;
 *   - return value 0 = success, non-zero = error; cna_last_error() gives the message of the last
 *     failure on the calling thread.  There is NO CPU fallback: a missing/non-sm_100 device is an error.
 */
#ifndef CNA_B200_H
#define CNA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { CNA_F32 = 0, CNA_BF16 = 1 } cna_dtype;

/* ---- library ---------------------------------------------------------------------------- */
int         cna_version(void);
const char* cna_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (B200). */
int         cna_device_check(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
unsigned long long cna_launch_count(void);

/* ---- layout (boundary helpers: the reference's tensors are logical NCHW) ------------------ */
/* src NCHW (src_dtype) -> dst NHWC with channel count c_pad >= C (extra channels zero-filled). */
int cna_nchw_to_nhwc(const void* src, void* dst, int B, int C, int H, int W, int c_pad,
                     int src_dtype, int dst_dtype, void* stream);
/* uint8 NCHW images -> normalised bf16 NHWC, channels zero-padded to c_pad (8 or 16):  dst = (src - mean255[c]) * inv_std255[c].
 * This is the input boundary of an image pipeline: the reference's loaders (timm create_loader with its prefetcher,
 * src/approx/hooks/fps.py:38-49, hooks/finetune.py:164-175) ship uint8 batches to the GPU and normalise there with
 * mean * 255 / std * 255; here the normalisation is fused with the layout change, so a batch costs 1 byte per value over PCIe. */
int cna_u8_nchw_to_nhwc_norm(const void* src_u8, void* dst_bf16, const float* mean255, const float* inv_std255, int B, int C, int H,
                             int W, int c_pad, void* stream);
/* src NHWC with c_pad channels -> dst NCHW with C channels. */
int cna_nhwc_to_nchw(const void* src, void* dst, int B, int C, int H, int W, int c_pad,
                     int src_dtype, int dst_dtype, void* stream);

/* ---- weight packing ------------------------------------------------------------------------ */
/* Dense conv weight in the reference's OIHW layout, with kh==1 or kw==1 (a run of `taps` taps),
 * fp32:  w[o][i][t]  ->  packed[t][i_pad][o_pad] in `dtype` (zero-filled padding).
 * Replaces nothing in the reference: it is the cache behind `.weight.data` assignments
 * (src/approx/core/low_rank_exp.py:181-184, src/approx/core/msca_rep.py:149-168). */
int cna_pack_taps_weight(const float* w_oit, void* packed, int O, int I, int taps,
                         int o_pad, int i_pad, int dtype, void* stream);
/* Depthwise weight (Cout,1,kh,kw) fp32 -> packed[kh*kw][Cout] in `dtype`. */
int cna_pack_dw_weight(const float* w, void* packed, int Cout, int kh, int kw, int dtype, void* stream);

/* ---- generic dense "tap" convolution (FFMA path; exact fp32) -------------------------------- */
/* y[b,ho,wo,n] = act( sum_t sum_c x[b, ho*sh - pad_h + t*ah, wo*sw - pad_w + t*aw, c] * w[t][c][n] + bias[n] )
 *               (* gate[b,ho,wo,n] if gate != NULL)
 * axis 0: taps run along H (ah=1, aw=0, pad applies to H); axis 1: along W.  taps==1 is a 1x1 conv.
 * Replaces: nn.Conv2d calls at src/approx/layers/low_rank_conv.py:49,79-81 (d_conv, v_conv, dense h_conv)
 * and src/approx/layers/msca.py:23,26 (channel_mix and the `x *` gate). */
int cna_conv_taps_forward(const void* x, const void* w_packed, const float* bias, const void* gate, void* y,
                          int B, int H, int W, int C, int N, int taps, int axis,
                          int stride_h, int stride_w, int pad, int relu, int dtype, void* stream);

/* 1x1 convolution on tcgen05 (bf16 NHWC, fp32 accumulation): y[p][n] = act(sum_c a[p][c] * w[n][c] + bias[n]) [* gate[p][n]]
 * over P = B*H*W pixels.  w_packed = the (N, C, 1, 1) conv weight as bf16 (cna_conv1x1_pack).  gate may be NULL; with a gate
 * the conv output is rounded to bf16 before the multiply, as the reference's two materialised tensors are.
 * Replaces nn.Conv2d(C, N, 1) at src/approx/layers/msca.py:23,26 (channel_mix and the `x *` gate) and
 * src/approx/layers/low_rank_conv.py:49,52 (d_conv) for bf16 tensors; cna_conv_taps_forward stays the fp32 / any-shape path. */
int cna_conv1x1_tc_supported(int C, int N);
int cna_conv1x1_pack(const float* w_nc, void* packed, int N, int C, void* stream);
int cna_conv1x1_tc_forward(const void* a, const void* w_packed, const float* bias, const void* gate, void* y, long long P, int C,
                           int N, int relu, void* stream);
/* The same kernel with the epilogues an MSCAN attention block needs around MSCA (src/approx/models/mscan.py:65-80, 108-110):
 *   act: 0 none, 1 ReLU, 2 GELU (erf form, nn.GELU's default) on v = acc + bias[n];
 *   aux_mode 0: aux = multiplicative gate (or NULL), y = bf16(v) * aux            — the call above;
 *   aux_mode 1: aux = residual tensor, y = alpha[n] * v + beta[n] * aux + delta[n]  (fp32, one rounding; NULL arrays = 1, 1, 0):
 *               proj_2 + shortcut + layer_scale_1 + the outer residual of MultiScaleConvAttnModule.forward in one epilogue. */
int cna_conv1x1_tc_forward_ex(const void* a, const void* w_packed, const float* bias, const void* aux, const float* alpha,
                              const float* beta, const float* delta, void* y, long long P, int C, int N, int act, int aux_mode,
                              void* stream);

/* The same kernel on sub-matrices: a is [P][lda] with the C contracted channels at its start, w_packed [N][ldw], y and aux [P][ldy]
 * (element strides; lda, ldw multiples of 8, ldy of 16; pointers may be offset into a wider tensor).  This is how the FFN of an MSCAN
 * block (src/approx/models/mscan.py:27-62) runs on it: fc1 (N = hidden > 256 allowed here: the launch's CTAs split into groups per
 * slab of 256 output channels), fc2 contracts the hidden channels in slabs whose weight fits in shared memory, accumulating into y
 * through the residual epilogue (aux = y). */
int cna_conv1x1_tc_forward_strided(const void* a, long long lda, const void* w_packed, long long ldw, const float* bias, const void* aux,
                                   const float* alpha, const float* beta, const float* delta, void* y, long long ldy, long long P, int C,
                                   int N, int act, int aux_mode, void* stream);

/* ---- fused scheme-1 low-rank convolution (tcgen05) --------------------------------------------- */
/* y = d_conv(s_conv(x)) in ONE kernel for bf16 NHWC tensors, stride 1, 'same' padding: the (B, C*M, H, W) intermediate of
 * LowRankExpConvV1.forward (src/approx/layers/low_rank_conv.py:44-52; SeparableConv :17-25 after decomp()) is produced by CUDA
 * cores 128 pixels x 64 wide channels at a time straight into the UMMA A-operand layout in shared memory, rounded to bf16 where
 * the reference's bf16 modules materialise it, and consumed by tcgen05 against the streamed d_conv weight.
 *   x         (B,H,W,C) bf16, dense
 *   s_packed  fp32, per K-chunk of 64 wide channels (wide channel c*M+m):
 *               separable == 0: [d*d taps (i*d+j)][64] basis filters then [64] s_conv bias      (s_conv = nn.Conv2d, groups=C)
 *               separable == 1: [d v-taps][64] then [d h-taps][64]                                (SeparableConv, bias-free)
 *   dw_packed bf16 [N][C*M] (cna_conv1x1_pack of d_conv.weight), d_bias fp32 (N,) or NULL
 *   y         (B,H,W,N) bf16
 * Covered: d in {3,5}, M in {2,4,8,16}, C*M % 64 == 0, C % 8 == 0, N % 16 == 0, N <= 512 (N % 32 == 0 above 256), W <= 128;
 * everything else stays on cna_dwconv_forward + cna_conv1x1_tc_forward / cna_conv_taps_forward. */
int cna_lowrank1_fused_supported(int B, int H, int W, int C, int N, int M, int d, int separable);
int cna_lowrank1_fused_forward(const void* x, const float* s_packed, const void* dw_packed, const float* d_bias, void* y, int B, int H,
                               int W, int C, int N, int M, int d, int separable, void* stream);

/* ---- depthwise convolution (generic) -------------------------------------------------------- */
/* Grouped conv with groups=Cin and channel multiplier `mult`: output channel co reads input channel co/mult.
 * w_packed[kh*kw][Cin*mult]; bias may be NULL.
 * res/res_p: optional fused FixPaddingBias (src/approx/layers/depth_seperable_conv.py:57-61): res[2][Cout][res_p] fp32
 * added to the top/bottom min(Ho,res_p) output rows.
 * add: optional tensor (same shape as y) added to the result (ParallelConv's running sum,
 * src/approx/layers/depth_seperable_conv.py:45-47).
 * Replaces: nn.Conv2d(groups=C) at low_rank_conv.py:17-22,44-48; depth_seperable_conv.py:12-13; msca.py:18-19;
 * core/msca_rep.py:44-45. */
int cna_dwconv_forward(const void* x, const void* w_packed, const float* bias, const float* res, int res_p,
                       const void* add, void* y,
                       int B, int H, int W, int Cin, int mult, int kh, int kw,
                       int stride_h, int stride_w, int pad_h, int pad_w, int dtype, void* stream);

/* Depthwise 3x3 + bias + exact GELU in one pass (bf16 NHWC, stride 1, 'same' padding, C % 32 == 0) on the TMA-fed shared-memory
 * pipeline of csrc/dwstrip_tma.cu: FFN.dconv followed by FFN.act of an MSCAN block (src/approx/models/mscan.py:38-45, 58-59).
 * act: 0 = none, 2 = GELU.  w_packed [9][C] bf16 as for cna_dwconv_forward. */
int cna_dwconv3x3_act_forward(const void* x, const void* w_packed, const float* bias, void* y, int B, int H, int W, int C, int act,
                              void* stream);

/* out = a + b elementwise over `count` elements (ParallelConv's identity branch). */
int cna_add(const void* a, const void* b, void* out, long long count, int dtype, void* stream);

/* In-place FixPaddingBias on an NHWC tensor (depth_seperable_conv.py:57-61). */
int cna_fix_padding_bias(void* x, const float* res, int B, int H, int W, int C, int p, int dtype, void* stream);

/* ---- scheme-2 low-rank conv, fused (tcgen05 / TMEM / TMA) ------------------------------------ */
/* Fused  y = h_conv(v_conv(x)) + bias  for stride 1, odd d, pad d/2 ("same"), bf16 NHWC:
 *   T[b,h,w,k] = sum_dy sum_c x[b,h+dy-d/2,w,c] * wv[dy][k][c]        (rounded to bf16, stays in shared memory)
 *   y[b,h,w,n] = act( sum_dx sum_k T[b,h,w+dx-d/2,k] * wh[dx][n][k] + bias[n] )
 * wv_packed: [d][k_pad][c_pad] bf16 (c contiguous), wh_packed: [d][n][k_pad] bf16 (k contiguous);
 * build them with cna_lowrank2_pack. c_pad = C rounded up to 16, k_pad in {16,32,64}.
 * x has c_pad channels (NHWC), y has N channels.
 * Replaces: LowRankExpConvV2.forward, src/approx/layers/low_rank_conv.py:83-84 (true scheme 2: dense h_conv). */
int cna_lowrank2_pack(const float* v_w /*(K,C,d,1)*/, const float* h_w /*(N,K,1,d)*/,
                      void* wv_packed, void* wh_packed, int C, int K, int N, int d,
                      int c_pad, int k_pad, void* stream);
/* relu: fuse max(.,0); pool: additionally fuse a 2x2/stride-2 max-pool (H, W even) — y is then [B, H/2, W/2, N]. */
int cna_lowrank2_fused_forward(const void* x, const void* wv_packed, const void* wh_packed, const float* bias,
                               void* y, int B, int H, int W, int c_pad, int k_pad, int N, int d,
                               int relu, int pool, void* stream);
/* Same as cna_lowrank2_fused_forward (no pool), additionally writing the bf16 intermediate T = v_conv(x), [B, H, W, k_pad], from the
 * kernel's T-epilogue: the reconstruction step needs T for h_conv's weight gradient and gets it without a second pass over x. */
int cna_lowrank2_fused_forward_t(const void* x, const void* wv_packed, const void* wh_packed, const float* bias, void* y, void* t_out,
                                 int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu, void* stream);
/* 3-tap 'same' convolution C -> K <= 64 on tcgen05 (the fused kernel's GEMM1 alone): out[B,H,W,k_pad] bf16, taps along H (axis 0,
 * = v_conv) or along W (axis 1, = the data gradient through h_conv when w_packed holds its flipped, transposed taps).
 * w_packed [3][k_pad][c_pad] bf16 as produced by cna_lowrank2_pack for a (K, C, 3, 1) weight.
 * Replaces autograd's conv backward-data for h_conv and the v_conv recompute (src/approx/hooks/finetune.py:329-330). */
int cna_tapconv3_tc_forward(const void* x, const void* w_packed, void* out, int B, int H, int W, int c_pad, int k_pad, int axis,
                            void* stream);
/* Diagnostic variant: CTA 0 additionally records clock64() stamps, 20 per tile for its first `trace_tiles` tiles
 * (`trace` holds trace_tiles + 1 rows of 20; the last row = {clock64, globaltimer} at CTA 0's start and end)
 * (0 x issued, 1 GEMM1 turn, 2 x landed, 3 GEMM1 issued, 4 GEMM2 turn, 5 T ready, 6 T-epilogue start, 7 T stored,
 * 8 Y-epilogue start, 9 y store issued, 10 T loaded from TMEM, 11 A buffer free, 12 A copies issued, 13 Y loaded from TMEM,
 * 14 staging buffer free, 15 y staged, 16 GEMM2 got its accumulator, 17 GEMM2 issued).  Used by benchmarks/trace_lowrank2.py to read the pipeline's critical path. */
int cna_lowrank2_fused_forward_traced(const void* x, const void* wv_packed, const void* wh_packed, const float* bias,
                                      void* y, int B, int H, int W, int c_pad, int k_pad, int N, int d, int relu,
                                      void* stream, long long* trace, int trace_tiles);
/* 1 if the fused kernel supports this configuration, else 0 (caller falls back to two cna_conv_taps_forward). */
int cna_lowrank2_fused_supported(int H, int W, int c_pad, int k_pad, int N, int d);
/* The fused kernel's shared-memory plan for a layer, computed on the host without touching a device (for tests and tooling):
 * out[0..7] = {dynamic smem bytes, Wv resident (0/1), Wh resident (0/1), x-ring stages, Wh-ring slots, TMA-store buffers per epilogue
 * group, channel chunk CK, GEMM2 N-tile}.  smem_limit = the device's opt-in limit (232448 on B200).  Returns non-zero if the
 * configuration is not supported or does not fit. */
int cna_lowrank2_fused_plan(int c_pad, int k_pad, int N, int pool, int smem_limit, int* out);

/* ---- MSCA fused depthwise chain ------------------------------------------------------------------ */
/* s = [conv0(x)] + sum_b cascade_b(conv0(x))   (conv0: depthwise 5x5 + bias; cascade: depthwise 1xk + b1 then kx1 + b2,
 * k in {7,11,21}, up to 3 branches; `identity` adds conv0's output; res/res_p = fused FixPaddingBias) in ONE kernel
 * that reads x once and writes s once (sliding-window rings in shared memory).
 * w0 [25][C], b0 [C]; w1cat / w2cat [sum k][C]; b1cat / b2cat [nb][C] (NULL = no bias); all fp32.
 * Needs C % 16 == 0 (bf16) or C % 8 == 0 (fp32).
 * Replaces MSCA.forward's `sd_convs(conv0(x))` (src/approx/layers/msca.py:26, depth_seperable_conv.py:12-16,45-47) and
 * MscaRep's decomp>=1 forms (src/approx/core/msca_rep.py:46-55). */
int cna_msca_sd_supported(int C, int k1, int nb, const int* ks_host, int dtype);
/* The same chain as ONE host call that enqueues the register-tiled depthwise / strip kernels back to back (conv0 5x5 -> a0; per
 * cascade 1xk -> t, kx1 (+ running sum) -> s; identity add; border fix), every intermediate rounded exactly where the reference
 * materialises a tensor.  On B200 these HBM-bound launches beat the one-pass marching kernel (which is instruction-issue bound);
 * doing the sequencing here instead of in Python removes ~20 us of host work per launch.  Weights packed [taps][C] in the tensor
 * dtype (cna_pack_dw_weight); w1 / b1 / w2 / b2 are HOST arrays of nb device pointers (bias entries may be NULL);
 * a0, t, s: caller-provided tensors shaped like x (s is the result).  C % 32 == 0, k1 = 5, strips in {3,5,7,11,21}. */
int cna_msca_chain_supported(int C, int k1, int nb, const int* ks_host, int dtype);
int cna_msca_chain_forward(const void* x, void* a0, void* t, void* s, const void* w0, const float* b0, const void* const* w1,
                           const float* const* b1, const void* const* w2, const float* const* b2, const int* ks_host, int nb,
                           int identity, const float* res, int res_p, int B, int H, int W, int C, int dtype, void* stream);
int cna_msca_sd_forward(const void* x, void* y, const float* w0, const float* b0, const float* w1cat, const float* b1cat,
                        const float* w2cat, const float* b2cat, const float* res, int res_p, const int* ks_host, int nb,
                        int identity, int B, int H, int W, int C, int dtype, void* stream);

/* ---- fitting step (device-side small linear algebra) ------------------------------------------- */
/* Batched SVD of small matrices, A (batch,m,n) fp32 row-major with 1 <= n <= m <= 64:
 * U (batch,m,n), S (batch,n) descending, Vh (batch,n,n).  One-sided Jacobi in fp64, one warp per matrix.
 * Replaces torch.linalg.svd at src/approx/layers/low_rank_conv.py:61 and src/approx/core/msca_rep.py:154. */
int cna_svd_small_batched(const float* A, float* U, float* S, float* Vh, int batch, int m, int n, void* stream);
/* Scheme-1 SVD initialisation: W (R = N*C rows, q = d*d cols, fp32) -> weights (R,M) = U[:, :M]*sqrt(S),
 * bases (M,q) = sqrt(S)*Vh[:M]  (zero-filled beyond rank q).  Gram matrix + Jacobi eigen-solve in fp64.
 * Replaces numpy.linalg.svd + scaling at src/approx/core/low_rank_exp.py:110-120. */
int cna_lowrank1_svd_fit(const float* W, int R, int q, int M, float* weights, float* bases, void* stream);
/* Truncated SVD of a large fp32 matrix A (m x n row-major, n <= m): r leading triplets U (m,r), S (r), Vh (r,n).
 * One-sided Jacobi in fp64, one launch per tournament round; synchronises the stream (fit-time call).
 * Backs the scheme-2 closed-form fit W[n,c,i,j] ~ sum_k v[k,c,i] h[n,k,j] (no counterpart in the reference, which
 * never shipped a scheme-2 app; the scheme-1 analogue is numpy.linalg.svd at src/approx/core/low_rank_exp.py:111). */
int cna_svd_jacobi_large(const float* A, int m, int n, int r, float* U, float* S, float* Vh, int max_sweeps, void* stream);
/* MSCA equivalent kernel (src/approx/core/msca_rep.py:57-143).  nb strip branches with odd sizes ks_host[i]
 * (ascending); w1cat / w2cat hold, branch after branch, the (C x k_i) row-major taps of conv1 (1 x k) and
 * conv2 (k x 1); b1 / b2 are (nb, C) or NULL.  Outputs: weight (C,K,K), bias (C), res (2,C,K/2), K = max k. */
int cna_msca_equiv_kernel(const float* w1cat, const float* b1, const float* w2cat, const float* b2,
                          const int* ks_host, int nb, int C, float* weight, float* bias, float* res, void* stream);

/* ---- reconstruction step ------------------------------------------------------------------------- */
/* Per-sample output L2 error and the gradient of  scale * sum_b ||y_new_b - y_old_b||_2  w.r.t. y_new:
 *   norms[b] = ||y_new_b - y_old_b||_2 ; grad = (y_new - y_old) * scale / norms[b]   (grad may be NULL: norms only).
 * sumsq_scratch: B floats of workspace.  Replaces get_l2_error + the autograd of torch.norm / torch.mean at
 * src/approx/hooks/finetune.py:80-82,315-322 (scale = l2_weight / (L * B_global)). */
int cna_l2_recon_loss_grad(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad,
                           int B, long long per_sample, float scale, int dtype, void* stream);
/* Same, additionally producing the bias gradient dbias[n] = sum_p grad[p][n] (fp32, overwritten; N = channels of y) — fused into
 * the gradient kernel for bf16 tensors so the step needs no separate pass over dL/dy. */
int cna_l2_recon_loss_grad_bias(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad, float* dbias,
                                int N, int B, long long per_sample, float scale, int dtype, void* stream);
/* Weight (and bias) gradient of a stride-1 'same' tap convolution (see cna_conv_taps_forward):
 *   dw[o][i][t] = sum_p x[p shifted by tap t][i] * g[p][o]   (reference OIHW layout, fp32; overwritten)
 *   dbias[o]    = sum_p g[p][o]                               (NULL to skip)
 * Replaces autograd's conv backward-filter for v_conv / h_conv / 1x1 convs (src/approx/hooks/finetune.py:330). */
int cna_conv_taps_wgrad(const void* x, const void* g, float* dw, float* dbias, int B, int H, int W, int C, int N,
                        int taps, int axis, int pad, int dtype, void* stream);
/* Weight gradient of a 3-tap 'same' convolution on tcgen05 (bf16 NHWC operands, fp32 result, dw overwritten):
 *   wide_is_input = 1 (v_conv):  dw[k][c][t] = sum_p a[p + (t-1) e_axis][c] * bm[p][k]   a = x  [B,H,W,ca_stride], bm = dL/dT [B,H,W,kb_stride]
 *   wide_is_input = 0 (h_conv):  dw[n][k][t] = sum_p bm[p + (t-1) e_axis][k] * a[p][n]   a = dL/dy [B,H,W,ca_stride], bm = T [B,H,W,kb_stride]
 * in the reference's (O, I, 3, 1) / (O, I, 1, 3) layouts with I resp. O = Ca real wide channels and kb real narrow channels (<= 64).
 * The reduction over pixels runs on the tensor cores with both operands MN-major straight from the NHWC tensors.
 * Replaces autograd's conv backward-filter for v_conv / h_conv (src/approx/hooks/finetune.py:330); cna_conv_taps_wgrad is the
 * fp32 / any-shape path. */
int cna_wgrad3_tc(const void* a, const void* bm, float* dw, int B, int H, int W, int ca_stride, int Ca, int kb_stride, int kb, int axis,
                  int wide_is_input, void* stream);
/* ---- backward pieces of the depthwise layers (MSCA training; csrc/dw_backward.cu) ---------------------------------------
 * Replaces autograd through nn.Conv2d(groups=C), FixPaddingBias and the `x *` gate (src/approx/hooks/finetune.py:329-330 with an
 * MSCA block's sd_convs trainable).  The DATA gradient of a stride-1 'same' depthwise conv is cna_dwconv_forward on the flipped filter.
 * dw[co][i][j] = sum_{b,h,w} x[b,h+i-pad_h,w+j-pad_w,co/mult] * g[b,h,w,co]  (fp32, (Cin*mult,1,kh,kw) layout, overwritten). */
int cna_dwconv_wgrad(const void* x, const void* g, float* dw, int B, int H, int W, int Cin, int mult, int kh, int kw, int pad_h, int pad_w,
                     int dtype, void* stream);
/* Data gradient for any channel multiplier (scheme 1's s_conv): gx[.,ci] = sum_m sum_taps g[shifted, ci*mult+m] * w[ci*mult+m][tap];
 * w_packed [kh*kw][Cin*mult] as for cna_dwconv_forward (NOT flipped). */
int cna_dwconv_dgrad(const void* g, const void* w_packed, void* gx, int B, int H, int W, int Cin, int mult, int kh, int kw, int pad_h,
                     int pad_w, int dtype, void* stream);
/* dres[2][C][p] (fp32, overwritten): gradient of FixPaddingBias' residues = sums of g over (b, w) on the touched border rows. */
int cna_fix_padding_bias_grad(const void* g, float* dres, int B, int H, int W, int C, int p, int dtype, void* stream);
/* out = a * b elementwise (the gate and its gradients). */
int cna_mul(const void* a, const void* b, void* out, long long count, int dtype, void* stream);

/* dbias[n] = sum_p g[p][n] over P pixels (fp32, overwritten): the bias gradient of any conv whose output gradient is g. */
int cna_bias_grad(const void* g, float* dbias, long long P, int N, int dtype, void* stream);
/* Packs w[o][i][t] for the INPUT gradient of a tap conv: packed[taps-1-t][o][i]; feed to cna_conv_taps_forward with
 * C=O, N=I, pad = taps-1-pad to get dL/dx from dL/dy. */
int cna_pack_taps_weight_dgrad(const float* w_oit, void* packed, int O, int I, int taps, int dtype, void* stream);
/* Fused AdamW (torch.optim.AdamW semantics) on a flat fp32 buffer; grad is multiplied by grad_scale first
 * (1/world_size after an all-reduce(SUM)).  Weight decay applies to the first n_decay elements only: the reference builds its
 * optimizer with timm's create_optimizer_v2 (src/approx/hooks/finetune.py:188), whose default filter gives biases no decay, so
 * callers lay the flat buffer out as [weights | biases] and pass n_decay = number of weight elements (n_decay = n: uniform decay).
 * Replaces optimizer.step() at src/approx/hooks/finetune.py:331. */
int cna_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, long long n_decay, float lr, float beta1,
                   float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream);
/* Same, with the step counter on the device (*step_dev holds the number of steps taken so far and is incremented): the call is
 * CUDA-graph capturable, so a whole reconstruction step can be replayed without host work. */
int cna_adamw_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, long long n_decay, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int* step_dev, float grad_scale, void* stream);

/* Backward of the per-sample output-L2 norm for autograd callers (the loss hook get_l2_error, src/approx/hooks/finetune.py:80-82,
 * reached from total_loss.backward() at :329): grad[b] = (y_new[b] - y_old[b]) * upstream[b] / norms[b], 0 where the norm is 0.
 * y_old == NULL is read as zero and norms == NULL as one, i.e. grad[b] = y_new[b] * upstream[b] (scaling a saved unit gradient). */
int cna_l2_norm_backward(const void* y_new, const void* y_old, const float* norms, const float* upstream, void* grad, int B,
                         long long per_sample, int dtype, void* stream);

/* Accumulating variants for a multi-layer reconstruction step (one memset of the flat gradient buffer instead of several per layer):
 * cna_wgrad3_tc_acc adds into dw; cna_l2_recon_loss_grad_bias_acc expects sumsq_scratch and dbias zeroed by the caller. */
int cna_wgrad3_tc_acc(const void* a, const void* bm, float* dw, int B, int H, int W, int ca_stride, int Ca, int kb_stride, int kb, int axis,
                      int wide_is_input, void* stream);
int cna_l2_recon_loss_grad_bias_acc(const void* y_new, const void* y_old, float* sumsq_scratch, float* norms, void* grad, float* dbias,
                                    int N, int B, long long per_sample, float scale, int dtype, void* stream);

/* ---- peer memory + fused gradient exchange / optimizer step (multi-GPU reconstruction) -------------------------------
 * Replaces DistributedDataParallel's gradient all-reduce inside total_loss.backward() and the optimizer.step() that follows it
 * (src/approx/hooks/finetune.py:153-160, 329-331) by ONE kernel over NVLink peer memory.
 *
 * cna_peer_alloc: cudaMalloc + zero a "symmetric" buffer on the current device and export its CUDA-IPC handle (64 bytes, host).
 * cna_peer_open : map another process' buffer from its handle (peer access is enabled lazily); cna_peer_close unmaps it.
 * Every rank allocates the same layout:  [header, cna_recon_sync_header_bytes() bytes: barrier flags, zero-initialised]
 *                                        [off_grad : n fp32 gradients]  [off_param : n fp32 parameters]
 *
 * cna_recon_sync_step (all ranks launch it once per step, after their gradient kernels, on `stream`):
 *   barrier A over the peers' flags; rank r sums slice r = [r*ceil(n/4/world)*4, ...) of all `world` gradient buffers in rank
 *   order, runs AdamW (torch semantics, grad * grad_scale first, no weight decay on bias elements) on that slice with ITS slice of
 *   exp_avg / exp_avg_sq (sharded optimizer state, indexed like the flat buffer), writes the new fp32 parameters into EVERY
 *   rank's parameter region; barrier B; then every rank re-packs all parameters into the bf16 operand buffers of the fused
 *   kernels and the fp32 bias vectors.  *step_dev (device int, steps taken so far) is read for the bias correction / barrier epoch
 *   and incremented: the launch is CUDA-graph capturable.
 *   Flat layout, per layer l (contiguous, in order): v_conv.weight (K,C,3,1) | h_conv.weight (N,K,1,3) | h_conv.bias (N);
 *     layer_offsets_host[4l..] = {off_v, off_h, off_b, end} in floats, layer_dims_host[5l..] = {K, C, N, k_pad, c_pad},
 *     layer_ptrs_host[4l..] = {wv [3][k_pad][c_pad] bf16, wh [3][N][k_pad] bf16, wd [3][k_pad][N] bf16 (h_conv transposed, taps
 *     flipped: the weight of the data-gradient conv), bias fp32 (N)} — device pointers whose padding the caller zeroed once.
 *   peers_host[r] = base of rank r's symmetric buffer as mapped in THIS process (peers_host[rank] = the local allocation).
 *   n % 4 == 0; world <= 8; nlayers <= 16.  A barrier that sees no progress for ~2 s sets an error word instead of hanging:
 *   cna_recon_sync_error reads it back (0 = fine). */
int cna_peer_alloc(long long bytes, void** ptr_out, unsigned char* handle64_host);
int cna_peer_open(const unsigned char* handle64_host, void** ptr_out);
int cna_peer_close(void* ptr);
int cna_peer_free(void* ptr);
int cna_recon_sync_header_bytes(void);
int cna_recon_sync_step(void* const* peers_host, int rank, int world, long long n, long long off_grad, long long off_param,
                        float* exp_avg, float* exp_avg_sq, int* step_dev, const long long* layer_offsets_host,
                        const int* layer_dims_host, void* const* layer_ptrs_host, int nlayers, float lr, float beta1, float beta2,
                        float eps, float weight_decay, float grad_scale, void* stream);
int cna_recon_sync_error(const void* local_buffer, int* err_host);

/* Alternating fit of scheme-1 bases with a nuclear-norm penalty on the device — replaces the cvxpy loop of LowRankExpV1.optimize
 * (src/approx/core/low_rank_exp.py:127-135, 160-176):
 *     minimise  sum_r || W_r - A_r B ||_2 + lambda * sum_m || reshape(B_m, d, d) ||_*   over A = weights (R x M), B = bases (M x d*d),
 * alternately in B (majorise-minimise + FISTA with singular-value soft-thresholding, `mm_rounds` x `fista_iters`) and in A (closed
 * form W B^T (B B^T)^+), for every lambda of lmda_list_host in order (warm starts), at most max_iter alternations per lambda, stopping
 * when the objective moves by less than epsilon — the reference's schedule.  weights / bases hold the initial point on entry (device,
 * fp32) and the result on exit.  obj_trace_host (optional, capacity trace_cap) receives the objective after every B-step and A-step.
 * W: R x d*d fp32 device; d <= 8; M <= 16.  Synchronises the stream (the stopping rule is evaluated on the host). */
int cna_lowrank1_alt_fit(const float* W, int R, int d, int M, float* weights, float* bases, const double* lmda_list_host, int n_lmda,
                         int max_iter, double epsilon, int mm_rounds, int fista_iters, double* obj_trace_host, int trace_cap,
                         int* trace_len_host, void* stream);

/* 2x2/stride-2 max pool, NHWC (VGG trunk glue between the approximated convs). */
int cna_maxpool2x2(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CNA_B200_H */
