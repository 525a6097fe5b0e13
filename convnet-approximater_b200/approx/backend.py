"""ctypes binding of the C ABI (include/cna_b200.h) + NHWC tensor plumbing.

torch is used for device memory and streams only.  Every compute call goes to
lib/libcna_b200.so; there is no CPU path and no eager-PyTorch fallback: a missing
library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_void_p

import torch

F32, BF16 = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
# CNA_B200_LIB: load another build of the same library (A/B measurements of a kernel revision on one box)
LIB_PATH = os.environ.get("CNA_B200_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libcna_b200.so")

_lib = None

# name -> argtypes ('p' pointer, 'i' int); all return int unless noted
_SIGS = {
    "cna_device_check": "",
    "cna_nchw_to_nhwc": "ppiiiiiiip",
    "cna_nhwc_to_nchw": "ppiiiiiiip",
    "cna_u8_nchw_to_nhwc_norm": "ppppiiiiip",
    "cna_pack_taps_weight": "ppiiiiiip",
    "cna_pack_dw_weight": "ppiiiip",
    "cna_conv_taps_forward": "pppppiiiiiiiiiiiip",
    "cna_dwconv_forward": "ppppippiiiiiiiiiiiip",
    "cna_dwconv3x3_act_forward": "ppppiiiiip",
    "cna_conv1x1_tc_forward_strided": "pLpLppppppLLiiiip",
    "cna_conv1x1_tc_supported": "ii",
    "cna_conv1x1_pack": "ppiip",
    "cna_conv1x1_tc_forward": "pppppLiiip",
    "cna_conv1x1_tc_forward_ex": "ppppppppLiiiip",
    "cna_lowrank1_fused_supported": "iiiiiiii",
    "cna_lowrank1_fused_forward": "pppppiiiiiiiip",
    "cna_fix_padding_bias": "ppiiiiiip",
    "cna_add": "pppLip",
    "cna_lowrank2_pack": "ppppiiiiiip",
    "cna_lowrank2_fused_forward": "pppppiiiiiiiiip",
    "cna_lowrank2_fused_supported": "iiiiii",
    "cna_lowrank2_fused_plan": "iiiiip",
    "cna_lowrank2_fused_forward_t": "ppppppiiiiiiiip",
    "cna_tapconv3_tc_forward": "pppiiiiiip",
    "cna_wgrad3_tc": "pppiiiiiiiiip",
    "cna_wgrad3_tc_acc": "pppiiiiiiiiip",
    "cna_l2_recon_loss_grad_bias_acc": "ppppppiiLfip",
    "cna_l2_norm_backward": "pppppiLip",
    "cna_peer_alloc": "Lpp",
    "cna_peer_open": "pp",
    "cna_peer_close": "p",
    "cna_peer_free": "p",
    "cna_recon_sync_header_bytes": "",
    "cna_recon_sync_step": "piiLLLppppppiffffffp",
    "cna_recon_sync_error": "pp",
    "cna_bias_grad": "ppLiip",
    "cna_dwconv_wgrad": "pppiiiiiiiiiip",
    "cna_dwconv_dgrad": "pppiiiiiiiiiip",
    "cna_fix_padding_bias_grad": "ppiiiiiip",
    "cna_mul": "pppLip",
    "cna_lowrank2_fused_forward_traced": "pppppiiiiiiiippi",
    "cna_maxpool2x2": "ppiiiiip",
    "cna_msca_sd_supported": "iiipi",
    "cna_msca_chain_supported": "iiipi",
    "cna_msca_chain_forward": "pppppppppppiipiiiiiip",
    "cna_msca_sd_forward": "pppppppppipiiiiiiip",
    "cna_svd_small_batched": "ppppiiip",
    "cna_svd_jacobi_large": "piiipppip",
    "cna_l2_recon_loss_grad": "pppppiLfip",
    "cna_l2_recon_loss_grad_bias": "ppppppiiLfip",
    "cna_conv_taps_wgrad": "ppppiiiiiiiiip",
    "cna_pack_taps_weight_dgrad": "ppiiiip",
    "cna_adamw_step": "ppppLLfffffifp",
    "cna_adamw_step_dev": "ppppLLfffffpfp",
    "cna_lowrank1_svd_fit": "piiippp",
    "cna_msca_equiv_kernel": "pppppiipppp",
    "cna_lowrank1_alt_fit": "piiipppiiDiipipp",
}


def lib():
    """The loaded shared library (loads on first use; raises if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python convnet-approximater_b200/build.py` "
                "(this backend has no CPU or eager-PyTorch fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, sig in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = c_int
            fn.argtypes = [{"p": c_void_p, "i": c_int, "L": ctypes.c_longlong, "f": ctypes.c_float, "D": ctypes.c_double}[ch] for ch in sig]
        L.cna_last_error.restype = ctypes.c_char_p
        L.cna_launch_count.restype = ctypes.c_ulonglong
        L.cna_version.restype = c_int
        _lib = L
    return _lib


def exported_symbols():
    return list(_SIGS) + ["cna_last_error", "cna_launch_count", "cna_version"]


def _check(rc):
    if rc != 0:
        raise RuntimeError("cna_b200: " + lib().cna_last_error().decode())


def launch_count() -> int:
    return int(lib().cna_launch_count())


def dtype_code(dt) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"cna_b200 supports float32 and bfloat16 tensors, got {dt}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cna_b200: tensor is not on a CUDA device — this backend has no CPU fallback")


_checked_devices = set()


def _ready(t):
    require_cuda(t)
    dev = t.device.index
    if dev not in _checked_devices:
        with torch.cuda.device(dev):
            _check(lib().cna_device_check())
        _checked_devices.add(dev)


def _p(t):
    return 0 if t is None else t.data_ptr()


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


# ------------------------------------------------------------------------------------------- #
# layout
# ------------------------------------------------------------------------------------------- #
def to_nhwc(x: torch.Tensor, c_pad: int | None = None, dtype=None) -> torch.Tensor:
    """Logical-NCHW tensor -> contiguous (B,H,W,c_pad) tensor.  Zero-copy for channels_last input."""
    _ready(x)
    B, C, H, W = x.shape
    c_pad = C if c_pad is None else c_pad
    dtype = x.dtype if dtype is None else dtype
    v = x.permute(0, 2, 3, 1)
    if c_pad == C and dtype == x.dtype and v.is_contiguous():
        return v
    if not x.is_contiguous():
        x = x.contiguous()  # exotic strides: torch re-strides, the library transposes
    out = torch.empty((B, H, W, c_pad), device=x.device, dtype=dtype)
    with torch.cuda.device(x.device):
        _check(lib().cna_nchw_to_nhwc(_p(x), _p(out), B, C, H, W, c_pad, dtype_code(x.dtype), dtype_code(dtype),
                                      _stream(x)))
    return out


def u8_to_nhwc_norm(x_u8: torch.Tensor, mean255: torch.Tensor, inv_std255: torch.Tensor, c_pad: int = 16) -> torch.Tensor:
    """uint8 (B,C,H,W) images -> normalised bf16 (B,H,W,c_pad): (x - mean255[c]) * inv_std255[c], padding channels zero."""
    _ready(x_u8)
    assert x_u8.dtype == torch.uint8 and x_u8.dim() == 4 and x_u8.is_contiguous()
    Bn, C, H, W = x_u8.shape
    assert mean255.dtype == inv_std255.dtype == torch.float32 and mean255.numel() == inv_std255.numel() == C
    out = torch.empty((Bn, H, W, c_pad), device=x_u8.device, dtype=torch.bfloat16)
    with torch.cuda.device(x_u8.device):
        _check(lib().cna_u8_nchw_to_nhwc_norm(_p(x_u8), _p(out), _p(mean255), _p(inv_std255), Bn, C, H, W, c_pad, _stream(x_u8)))
    return out


def from_nhwc(y: torch.Tensor, C: int | None = None) -> torch.Tensor:
    """(B,H,W,c_pad) buffer -> logical NCHW view (channels_last strides); copies only to drop padding."""
    B, H, W, cp = y.shape
    if C is None or C == cp:
        return y.permute(0, 3, 1, 2)
    return y[..., :C].permute(0, 3, 1, 2)


def to_nchw_contiguous(y: torch.Tensor, C: int | None = None, dtype=None) -> torch.Tensor:
    """(B,H,W,c_pad) buffer -> NCHW-contiguous tensor through the library's transpose kernel."""
    _ready(y)
    B, H, W, cp = y.shape
    C = cp if C is None else C
    dtype = y.dtype if dtype is None else dtype
    out = torch.empty((B, C, H, W), device=y.device, dtype=dtype)
    with torch.cuda.device(y.device):
        _check(lib().cna_nhwc_to_nchw(_p(y), _p(out), B, C, H, W, cp, dtype_code(y.dtype), dtype_code(dtype),
                                      _stream(y)))
    return out


# ------------------------------------------------------------------------------------------- #
# weight packing
# ------------------------------------------------------------------------------------------- #
def _f32c(w):
    return w.detach().to(torch.float32).contiguous()


def pack_taps(w: torch.Tensor, dtype, o_pad=None, i_pad=None) -> torch.Tensor:
    """Dense conv weight (O,I,kh,kw) with kh==1 or kw==1 -> packed [taps][i_pad][o_pad]."""
    _ready(w)
    O, I, kh, kw = w.shape
    assert kh == 1 or kw == 1, "tap convs have a (d,1) or (1,d) filter"
    taps = kh * kw
    o_pad = O if o_pad is None else o_pad
    i_pad = I if i_pad is None else i_pad
    out = torch.empty((taps, i_pad, o_pad), device=w.device, dtype=dtype)
    wf = _f32c(w)
    with torch.cuda.device(w.device):
        _check(lib().cna_pack_taps_weight(_p(wf), _p(out), O, I, taps, o_pad, i_pad, dtype_code(dtype), _stream(w)))
    return out


def pack_dw(w: torch.Tensor, dtype) -> torch.Tensor:
    """Depthwise weight (Cout,1,kh,kw) -> packed [kh*kw][Cout]."""
    _ready(w)
    Co, one, kh, kw = w.shape
    assert one == 1
    out = torch.empty((kh * kw, Co), device=w.device, dtype=dtype)
    wf = _f32c(w)
    with torch.cuda.device(w.device):
        _check(lib().cna_pack_dw_weight(_p(wf), _p(out), Co, kh, kw, dtype_code(dtype), _stream(w)))
    return out


# ------------------------------------------------------------------------------------------- #
# compute
# ------------------------------------------------------------------------------------------- #
def conv_taps(x, wpk, bias, N, taps, axis, stride=(1, 1), pad=0, relu=False, gate=None):
    """x (B,H,W,C) NHWC; wpk [taps][C][N]; bias fp32 (N,) or None -> (B,Ho,Wo,N)."""
    _ready(x)
    B, H, W, C = x.shape
    assert wpk.shape == (taps, C, N) and wpk.dtype == x.dtype, (wpk.shape, (taps, C, N), wpk.dtype, x.dtype)
    sh, sw = stride
    Ho = (H + 2 * pad - taps) // sh + 1 if axis == 0 else (H - 1) // sh + 1
    Wo = (W + 2 * pad - taps) // sw + 1 if axis == 1 else (W - 1) // sw + 1
    y = torch.empty((B, Ho, Wo, N), device=x.device, dtype=x.dtype)
    if gate is not None:
        assert gate.shape == y.shape and gate.dtype == x.dtype and gate.is_contiguous()
    with torch.cuda.device(x.device):
        _check(lib().cna_conv_taps_forward(_p(x), _p(wpk), _p(bias), _p(gate), _p(y), B, H, W, C, N, taps, axis,
                                           sh, sw, pad, int(relu), dtype_code(x.dtype), _stream(x)))
    return y


def conv1x1_tc_supported(C: int, N: int) -> bool:
    return bool(lib().cna_conv1x1_tc_supported(int(C), int(N)))


def conv1x1_pack(w: torch.Tensor) -> torch.Tensor:
    """(N,C,1,1) conv weight -> bf16 [N][C] (the K-major B operand of the tcgen05 1x1 conv)."""
    _ready(w)
    N, C = w.shape[:2]
    out = torch.empty((N, C), device=w.device, dtype=torch.bfloat16)
    wf = _f32c(w)
    with torch.cuda.device(w.device):
        _check(lib().cna_conv1x1_pack(_p(wf), _p(out), N, C, _stream(w)))
    return out


def conv1x1_tc(x, w_nc, bias, relu=False, gate=None):
    """x (B,H,W,C) bf16 NHWC, w_nc bf16 [N][C], bias fp32 (N,) or None -> (B,H,W,N) bf16 [* gate]."""
    _ready(x)
    Bn, H, W, C = x.shape
    N = w_nc.shape[0]
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and w_nc.shape == (N, C) and w_nc.dtype == torch.bfloat16
    y = torch.empty((Bn, H, W, N), device=x.device, dtype=torch.bfloat16)
    if gate is not None:
        assert gate.shape == y.shape and gate.dtype == x.dtype and gate.is_contiguous()
    with torch.cuda.device(x.device):
        _check(lib().cna_conv1x1_tc_forward(_p(x), _p(w_nc), _p(bias), _p(gate), _p(y), Bn * H * W, C, N, int(relu), _stream(x)))
    return y


def conv1x1_tc_ex(x, w_nc, bias, act="none", residual=None, alpha=None, beta=None, delta=None):
    """1x1 conv on tcgen05 with the epilogues of an MSCAN attention block: y = act(x . w^T + bias), act in {none, relu, gelu}, and with
    `residual` (same shape as y): y = alpha[n] * act(...) + beta[n] * residual + delta[n] (per-channel fp32 vectors or None)."""
    _ready(x)
    Bn, H, W, C = x.shape
    N = w_nc.shape[0]
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and w_nc.shape == (N, C) and w_nc.dtype == torch.bfloat16
    y = torch.empty((Bn, H, W, N), device=x.device, dtype=torch.bfloat16)
    if residual is not None:
        assert residual.shape == y.shape and residual.dtype == x.dtype and residual.is_contiguous()
    for v in (alpha, beta, delta):
        assert v is None or (v.dtype == torch.float32 and v.numel() == N and v.is_contiguous())
    code = {"none": 0, "relu": 1, "gelu": 2}[act]
    with torch.cuda.device(x.device):
        _check(lib().cna_conv1x1_tc_forward_ex(_p(x), _p(w_nc), _p(bias), _p(residual), _p(alpha), _p(beta), _p(delta), _p(y),
                                               Bn * H * W, C, N, code, 1 if residual is not None else 0, _stream(x)))
    return y


def conv1x1_tc_slab(a2d, k0, kc, w_nc, n0, nc, bias, y2d, act="none", aux2d=None, alpha=None, beta=None, delta=None):
    """One slab of a 1x1 conv on the tcgen05 kernel: y2d[:, n0:n0+nc] = epilogue(a2d[:, k0:k0+kc] . w_nc[n0:n0+nc, k0:k0+kc]^T + bias) with
    the epilogues of conv1x1_tc_ex; a2d (P, Ca), w_nc (N, Ca), y2d / aux2d (P, Ny) bf16 row-major (aux2d may be y2d: in-place
    accumulation over K slabs).  bias / alpha / beta / delta are fp32 vectors of nc elements (already sliced) or None."""
    _ready(a2d)
    P, Ca = a2d.shape
    N, Cw = w_nc.shape
    Ny = y2d.shape[1]
    assert a2d.dtype == torch.bfloat16 and w_nc.dtype == torch.bfloat16 and y2d.dtype == torch.bfloat16
    assert a2d.is_contiguous() and w_nc.is_contiguous() and y2d.is_contiguous() and y2d.shape[0] == P and Cw == Ca
    assert 0 <= k0 and k0 + kc <= Ca and 0 <= n0 and n0 + nc <= N and n0 + nc <= Ny
    for v in (bias, alpha, beta, delta):
        assert v is None or (v.dtype == torch.float32 and v.numel() == nc and v.is_contiguous())
    if aux2d is not None:
        assert aux2d.shape == y2d.shape and aux2d.dtype == y2d.dtype and aux2d.is_contiguous()
    code = {"none": 0, "relu": 1, "gelu": 2}[act]
    off = lambda t, e: t.data_ptr() + 2 * e
    with torch.cuda.device(a2d.device):
        _check(lib().cna_conv1x1_tc_forward_strided(off(a2d, k0), Ca, off(w_nc, n0 * Cw + k0), Cw, _p(bias), off(aux2d, n0) if aux2d is not None else 0,
                                                    _p(alpha), _p(beta), _p(delta), off(y2d, n0), Ny, P, kc, nc, code, 1 if aux2d is not None else 0,
                                                    _stream(a2d)))
    return y2d


def lowrank2_fused_plan(C: int, K: int, N: int, pool: bool = False, smem_limit: int = 232448) -> dict:
    """Host-only: the fused scheme-2 kernel's shared-memory plan for a (C -> K -> N) layer (no device needed)."""
    kp = 16 if K <= 16 else (32 if K <= 32 else 64)
    out = (ctypes.c_int * 8)()
    _check(lib().cna_lowrank2_fused_plan(round_up(C, 16), kp, int(N), int(bool(pool)), int(smem_limit), ctypes.addressof(out)))
    keys = ("smem_bytes", "wv_resident", "wh_resident", "x_stages", "wh_slots", "store_buffers", "channel_chunk", "n_tile")
    return dict(zip(keys, list(out)))


def lowrank1_fused_supported(B, H, W, C, N, M, d, separable) -> bool:
    return bool(lib().cna_lowrank1_fused_supported(int(B), int(H), int(W), int(C), int(N), int(M), int(d), int(bool(separable))))


def _pad_bases(w: torch.Tensor, C: int, M: int, m_pad: int) -> torch.Tensor:
    """(C*M, ...) per-wide-channel tensor -> (C*m_pad, ...) with zeros for the padding bases m >= M of every input channel."""
    if m_pad == M:
        return w
    out = torch.zeros((C, m_pad) + tuple(w.shape[1:]), device=w.device, dtype=w.dtype)
    out[:, :M] = w.reshape((C, M) + tuple(w.shape[1:]))
    return out.reshape((C * m_pad,) + tuple(w.shape[1:]))


def lowrank1_pack_s(s_conv, M: int, m_pad: int | None = None) -> torch.Tensor:
    """Basis filters of LowRankExpConvV1.s_conv in the per-chunk layout the fused scheme-1 kernel stages with one bulk copy:
    fp32 [C*m_pad/64][rows][64] with rows = d*d taps + bias (grouped d x d conv) or d v-taps + d h-taps (SeparableConv).  Values
    are rounded through bf16 like every other bf16-path weight.  m_pad > M pads every input channel's bases with zero filters, so
    a layer with M in 2..7 bases runs on the kernel instance for 4 or 8 (the padding channels produce zeros and meet zero weights)."""
    m_pad = M if m_pad is None else m_pad
    r = lambda w: w.detach().to(torch.bfloat16).to(torch.float32)
    if hasattr(s_conv, "v_conv"):
        v, h = s_conv.v_conv.weight, s_conv.h_conv.weight
        C, d = v.shape[0] // M, v.shape[2]
        CM = C * m_pad
        v, h = _pad_bases(r(v), C, M, m_pad), _pad_bases(r(h), C, M, m_pad)
        rows = torch.cat([v.reshape(CM // 64, 64, d), h.reshape(CM // 64, 64, d)], 2)
    else:
        w = s_conv.weight
        C, d = w.shape[0] // M, w.shape[2]
        CM = C * m_pad
        b = s_conv.bias.detach().to(torch.float32) if s_conv.bias is not None else torch.zeros(C * M, device=w.device)
        w, b = _pad_bases(r(w), C, M, m_pad), _pad_bases(b, C, M, m_pad)
        rows = torch.cat([w.reshape(CM // 64, 64, d * d), b.reshape(CM // 64, 64, 1)], 2)
    rows = rows.transpose(1, 2)                                             # [chunk][row][64 wide channels]
    # within a row: unit u's channels 8u..8u+3 at floats [4u, 4u+4), channels 8u+4..8u+7 at [32+4u, 32+4u+4): a quarter-warp's
    # 16-byte loads cover 128 contiguous bytes
    return rows.reshape(rows.shape[0], rows.shape[1], 8, 2, 4).transpose(2, 3).reshape(rows.shape[0], rows.shape[1], 64).contiguous()


def lowrank1_pack_d(d_conv_weight: torch.Tensor, C: int, M: int, m_pad: int) -> torch.Tensor:
    """d_conv.weight (N, C*M, 1, 1) -> bf16 [N][C*m_pad] with zero columns for the padding bases."""
    N = d_conv_weight.shape[0]
    w = d_conv_weight.detach().reshape(N, C * M)
    if m_pad != M:
        w = _pad_bases(w.t().contiguous(), C, M, m_pad).t().contiguous()
    return conv1x1_pack(w.reshape(N, C * m_pad, 1, 1))


def lowrank1_fused(x, s_packed, dw_packed, d_bias, M: int, d: int, separable: bool):
    """x (B,H,W,C) bf16 NHWC -> (B,H,W,N) bf16: d_conv(s_conv(x)) in one tcgen05 kernel (csrc/lowrank1_tc.cu)."""
    _ready(x)
    Bn, H, W, C = x.shape
    N = dw_packed.shape[0]
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and dw_packed.dtype == torch.bfloat16
    assert dw_packed.shape == (N, C * M) and dw_packed.is_contiguous() and s_packed.dtype == torch.float32 and s_packed.is_contiguous()
    assert s_packed.shape == (C * M // 64, 2 * d if separable else d * d + 1, 64), tuple(s_packed.shape)
    dev = x.device
    y = torch.empty((Bn, H, W, N), device=dev, dtype=torch.bfloat16)
    fn = lib().cna_lowrank1_fused_forward
    # (a ~20 us kernel at the reference's profiling batches: the device guard is skipped when x's device is already current)
    if torch.cuda.current_device() == dev.index:
        rc = fn(x.data_ptr(), s_packed.data_ptr(), dw_packed.data_ptr(), _p(d_bias), y.data_ptr(), Bn, H, W, C, N, M, d, 1 if separable else 0,
                torch.cuda.current_stream(dev).cuda_stream)
    else:
        with torch.cuda.device(dev):
            rc = fn(x.data_ptr(), s_packed.data_ptr(), dw_packed.data_ptr(), _p(d_bias), y.data_ptr(), Bn, H, W, C, N, M, d,
                    1 if separable else 0, torch.cuda.current_stream(dev).cuda_stream)
    if rc != 0:
        _check(rc)
    return y


def dwconv(x, wpk, bias, mult, kh, kw, stride=(1, 1), pad=(0, 0), res=None, add=None):
    """Depthwise conv, x (B,H,W,Cin) NHWC, wpk [kh*kw][Cin*mult]."""
    _ready(x)
    B, H, W, C = x.shape
    sh, sw = stride
    ph, pw = pad
    assert wpk.shape == (kh * kw, C * mult) and wpk.dtype == x.dtype
    Ho = (H + 2 * ph - kh) // sh + 1
    Wo = (W + 2 * pw - kw) // sw + 1
    y = torch.empty((B, Ho, Wo, C * mult), device=x.device, dtype=x.dtype)
    res_p = 0
    if res is not None:
        assert res.dtype == torch.float32 and res.is_contiguous() and res.shape[:2] == (2, C * mult)
        res_p = res.shape[2]
    if add is not None:
        assert add.shape == y.shape and add.dtype == x.dtype and add.is_contiguous()
    with torch.cuda.device(x.device):
        _check(lib().cna_dwconv_forward(_p(x), _p(wpk), _p(bias), _p(res), res_p, _p(add), _p(y), B, H, W, C, mult,
                                        kh, kw, sh, sw, ph, pw, dtype_code(x.dtype), _stream(x)))
    return y


def dwconv3x3_act(x, wpk, bias, act=None):
    """Depthwise 3x3 ('same', stride 1) + bias + optional exact GELU in one pass: x (B,H,W,C) bf16 NHWC, wpk [9][C] bf16."""
    _ready(x)
    B, H, W, C = x.shape
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and wpk.shape == (9, C) and wpk.dtype == x.dtype
    assert act in (None, "gelu")
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _check(lib().cna_dwconv3x3_act_forward(_p(x), _p(wpk), _p(bias), _p(y), B, H, W, C, 2 if act == "gelu" else 0, _stream(x)))
    return y


def fix_padding_bias_(x_nhwc, res):
    """In place on an NHWC tensor."""
    _ready(x_nhwc)
    B, H, W, C = x_nhwc.shape
    assert res.dtype == torch.float32 and res.is_contiguous() and res.shape[:2] == (2, C)
    with torch.cuda.device(x_nhwc.device):
        _check(lib().cna_fix_padding_bias(_p(x_nhwc), _p(res), B, H, W, C, res.shape[2], dtype_code(x_nhwc.dtype),
                                          _stream(x_nhwc)))
    return x_nhwc


def add(a, b):
    _ready(a)
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _check(lib().cna_add(_p(a), _p(b), _p(out), a.numel(), dtype_code(a.dtype), _stream(a)))
    return out


def maxpool2x2(x):
    _ready(x)
    B, H, W, C = x.shape
    y = torch.empty((B, H // 2, W // 2, C), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        _check(lib().cna_maxpool2x2(_p(x), _p(y), B, H, W, C, dtype_code(x.dtype), _stream(x)))
    return y


def msca_chain_supported(C, k1, ks, dtype) -> bool:
    arr = (c_int * len(ks))(*[int(k) for k in ks])
    return bool(lib().cna_msca_chain_supported(int(C), int(k1), len(ks), ctypes.cast(arr, c_void_p), dtype_code(dtype)))


def msca_chain(x, w0, b0, w1s, b1s, w2s, b2s, ks, identity, res=None):
    """conv0 5x5 -> per cascade (1xk, kx1 with running sum) [-> + conv0 output] [-> border fix] as one host call that enqueues the
    register-tiled depthwise / strip kernels.  w* packed [taps][C] in x.dtype, b* fp32 (C,) or None; lists have one entry per cascade."""
    _ready(x)
    Bn, H, W, C = x.shape
    nb = len(ks)
    assert x.is_contiguous() and len(w1s) == len(w2s) == len(b1s) == len(b2s) == nb
    a0, t, s = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    P = lambda lst: ctypes.cast((c_void_p * nb)(*[_p(v) for v in lst]), c_void_p)
    karr = (c_int * nb)(*[int(k) for k in ks])
    res_p = 0 if res is None else res.shape[2]
    with torch.cuda.device(x.device):
        _check(lib().cna_msca_chain_forward(_p(x), _p(a0), _p(t), _p(s), _p(w0), _p(b0), P(w1s), P(b1s), P(w2s), P(b2s),
                                            ctypes.cast(karr, c_void_p), nb, int(bool(identity)), _p(res), res_p, Bn, H, W, C,
                                            dtype_code(x.dtype), _stream(x)))
    return s


def msca_sd_supported(C, k1, ks, dtype):
    arr = (c_int * len(ks))(*ks)
    return bool(lib().cna_msca_sd_supported(C, k1, len(ks), ctypes.cast(arr, c_void_p), dtype_code(dtype)))


def msca_sd(x, w0, b0, w1cat, b1cat, w2cat, b2cat, ks, identity, res=None):
    """Fused conv0 -> strip cascades (+identity, +FixPaddingBias) on NHWC x; weights fp32 tap-major ([taps][C])."""
    _ready(x)
    B, H, W, C = x.shape
    assert x.is_contiguous()
    y = torch.empty_like(x)
    arr = (c_int * len(ks))(*ks)
    res_p = 0 if res is None else res.shape[2]
    with torch.cuda.device(x.device):
        _check(lib().cna_msca_sd_forward(_p(x), _p(y), _p(w0), _p(b0), _p(w1cat), _p(b1cat), _p(w2cat), _p(b2cat), _p(res),
                                         res_p, ctypes.cast(arr, c_void_p), len(ks), int(identity), B, H, W, C,
                                         dtype_code(x.dtype), _stream(x)))
    return y


def round_up(v, m):
    return (v + m - 1) // m * m


def lowrank2_k_pad(K):
    for kp in (16, 32, 64):
        if K <= kp:
            return kp
    return None


def lowrank2_fused_supported(H, W, C, K, N, d):
    kp = lowrank2_k_pad(K)
    if kp is None:
        return False
    return bool(lib().cna_lowrank2_fused_supported(H, W, round_up(C, 16), kp, N, d))


def lowrank2_pack(v_w, h_w):
    """v_w (K,C,d,1), h_w (N,K,1,d) -> bf16 (wv [d][k_pad][c_pad], wh [d][N][k_pad])."""
    _ready(v_w)
    K, C, d, _ = v_w.shape
    N = h_w.shape[0]
    c_pad, k_pad = round_up(C, 16), lowrank2_k_pad(K)
    wv = torch.empty((d, k_pad, c_pad), device=v_w.device, dtype=torch.bfloat16)
    wh = torch.empty((d, N, k_pad), device=v_w.device, dtype=torch.bfloat16)
    vf, hf = _f32c(v_w), _f32c(h_w)
    with torch.cuda.device(v_w.device):
        _check(lib().cna_lowrank2_pack(_p(vf), _p(hf), _p(wv), _p(wh), C, K, N, d, c_pad, k_pad, _stream(v_w)))
    return wv, wh


def lowrank2_fused_traced(x, wv, wh, bias, tiles=32, relu=False):
    """Diagnostic: like lowrank2_fused, also returns a (tiles + 1, 20) int64 tensor of CTA-0 clock stamps."""
    _ready(x)
    Bn, H, W, cp = x.shape
    d, kp, _ = wv.shape
    N = wh.shape[1]
    y = torch.empty((Bn, H, W, N), device=x.device, dtype=torch.bfloat16)
    trace = torch.zeros((tiles + 1, 20), device=x.device, dtype=torch.int64)   # last row: clock/globaltimer at start, end
    with torch.cuda.device(x.device):
        _check(lib().cna_lowrank2_fused_forward_traced(_p(x), _p(wv), _p(wh), _p(bias), _p(y), Bn, H, W, cp, kp, N, d,
                                                       int(relu), _stream(x), _p(trace), tiles))
    return y, trace


def lowrank2_fused_with_t(x, wv, wh, bias, relu=False, out=None):
    """Like lowrank2_fused, also returning the bf16 intermediate T = v_conv(x), (B,H,W,k_pad), written by the same kernel.
    `out` = (y, t) pre-allocated buffers (CUDA-graph capture: fixed addresses)."""
    _ready(x)
    B, H, W, cp = x.shape
    d, kp, cp2 = wv.shape
    N = wh.shape[1]
    assert cp == cp2 and x.dtype == torch.bfloat16 and x.is_contiguous()
    if out is not None:
        y, t = out
        assert y.shape == (B, H, W, N) and t.shape == (B, H, W, kp) and y.is_contiguous() and t.is_contiguous()
    else:
        y = torch.empty((B, H, W, N), device=x.device, dtype=torch.bfloat16)
        t = torch.empty((B, H, W, kp), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _check(lib().cna_lowrank2_fused_forward_t(_p(x), _p(wv), _p(wh), _p(bias), _p(y), _p(t), B, H, W, cp, kp, N, d, int(relu), _stream(x)))
    return y, t


def tapconv3_tc(x, w_packed, axis, out=None):
    """3-tap 'same' conv on tcgen05: x (B,H,W,c_pad) bf16, w_packed [3][k_pad][c_pad] bf16 -> (B,H,W,k_pad) bf16; taps along H (axis 0) or W (1)."""
    _ready(x)
    B, H, W, cp = x.shape
    d, kp, cp2 = w_packed.shape
    assert d == 3 and cp == cp2 and x.dtype == torch.bfloat16 and x.is_contiguous() and w_packed.dtype == torch.bfloat16
    if out is None:
        out = torch.empty((B, H, W, kp), device=x.device, dtype=torch.bfloat16)
    assert out.shape == (B, H, W, kp) and out.is_contiguous()
    with torch.cuda.device(x.device):
        _check(lib().cna_tapconv3_tc_forward(_p(x), _p(w_packed), _p(out), B, H, W, cp, kp, int(axis), _stream(x)))
    return out


def wgrad3_tc(a, bm, dw, Ca, kb, axis, wide_is_input, accumulate=False):
    """Writes the 3-tap conv weight gradient into fp32 `dw` (reference OIHW layout) on tcgen05.
    a (B,H,W,ca_stride) bf16 = the wide tensor (x or dL/dy), bm (B,H,W,kb_stride<=64) bf16 = the narrow one (dL/dT or T).
    accumulate=True adds into `dw` (the caller zeroed it) instead of overwriting it."""
    _ready(a)
    B, H, W, cs = a.shape
    ks = bm.shape[3]
    assert bm.shape[:3] == a.shape[:3] and a.dtype == bm.dtype == torch.bfloat16 and a.is_contiguous() and bm.is_contiguous()
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.numel() == Ca * kb * 3
    fn = lib().cna_wgrad3_tc_acc if accumulate else lib().cna_wgrad3_tc
    with torch.cuda.device(a.device):
        _check(fn(_p(a), _p(bm), _p(dw), B, H, W, cs, int(Ca), ks, int(kb), int(axis), int(wide_is_input), _stream(a)))
    return dw


def adamw_step_dev_(p, g, m, v, step_dev, lr, betas, eps, wd, grad_scale=1.0, n_decay=None):
    """Graph-capturable fused AdamW: `step_dev` is a 1-element int32 device tensor counting the steps taken.
    Weight decay covers the first `n_decay` elements (default: all)."""
    _ready(p)
    assert step_dev.dtype == torch.int32 and step_dev.numel() == 1
    nd = p.numel() if n_decay is None else int(n_decay)
    with torch.cuda.device(p.device):
        _check(lib().cna_adamw_step_dev(_p(p), _p(g), _p(m), _p(v), p.numel(), nd, float(lr), float(betas[0]), float(betas[1]), float(eps),
                                        float(wd), _p(step_dev), float(grad_scale), _stream(p)))
    return p


def lowrank2_pack_into(v_w, h_w, wv, wh):
    """cna_lowrank2_pack into existing buffers (fixed addresses: what a captured CUDA graph needs)."""
    _ready(v_w)
    K, C, d = v_w.shape[0], v_w.shape[1], v_w.shape[2]
    N = h_w.shape[0]
    assert v_w.dtype == h_w.dtype == torch.float32 and v_w.is_contiguous() and h_w.is_contiguous()
    with torch.cuda.device(v_w.device):
        _check(lib().cna_lowrank2_pack(_p(v_w), _p(h_w), _p(wv), _p(wh), C, K, N, d, wv.shape[2], wv.shape[1], _stream(v_w)))
    return wv, wh


def dwconv_wgrad(x, g, kh, kw):
    """Depthwise (stride 1, 'same', any channel multiplier) weight gradient: x (B,H,W,Cin), g (B,H,W,Cin*mult) -> dw fp32 (Cout,1,kh,kw)."""
    _ready(x)
    Bn, H, W, Cin = x.shape
    Cout = g.shape[3]
    assert g.shape[:3] == x.shape[:3] and Cout % Cin == 0 and g.dtype == x.dtype and x.is_contiguous() and g.is_contiguous()
    dw = torch.empty((Cout, 1, kh, kw), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _check(lib().cna_dwconv_wgrad(_p(x), _p(g), _p(dw), Bn, H, W, Cin, Cout // Cin, kh, kw, kh // 2, kw // 2, dtype_code(x.dtype),
                                      _stream(x)))
    return dw


def dwconv_dgrad(g, wpk, mult, kh, kw):
    """Data gradient of a stride-1 'same' depthwise conv with channel multiplier: g (B,H,W,Cin*mult), wpk [kh*kw][Cin*mult] -> (B,H,W,Cin)."""
    _ready(g)
    Bn, H, W, Cout = g.shape
    assert Cout % mult == 0 and wpk.shape == (kh * kw, Cout) and wpk.dtype == g.dtype and g.is_contiguous()
    gx = torch.empty((Bn, H, W, Cout // mult), device=g.device, dtype=g.dtype)
    with torch.cuda.device(g.device):
        _check(lib().cna_dwconv_dgrad(_p(g), _p(wpk), _p(gx), Bn, H, W, Cout // mult, int(mult), kh, kw, kh // 2, kw // 2,
                                      dtype_code(g.dtype), _stream(g)))
    return gx


def fix_padding_bias_grad(g, p):
    """Gradient of FixPaddingBias.res: g (B,H,W,C) NHWC -> fp32 (2,C,p)."""
    _ready(g)
    Bn, H, W, C = g.shape
    assert g.is_contiguous()
    dres = torch.empty((2, C, p), device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        _check(lib().cna_fix_padding_bias_grad(_p(g), _p(dres), Bn, H, W, C, int(p), dtype_code(g.dtype), _stream(g)))
    return dres


def mul(a, b):
    """Elementwise product of two same-shape contiguous tensors."""
    _ready(a)
    assert a.shape == b.shape and a.dtype == b.dtype and a.is_contiguous() and b.is_contiguous()
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _check(lib().cna_mul(_p(a), _p(b), _p(out), a.numel(), dtype_code(a.dtype), _stream(a)))
    return out


def bias_grad(g, db):
    """db[n] = sum over pixels of g (B,H,W,N) NHWC; db fp32 (N,), overwritten."""
    _ready(g)
    N = g.shape[3]
    assert g.is_contiguous() and db.dtype == torch.float32 and db.numel() == N
    with torch.cuda.device(g.device):
        _check(lib().cna_bias_grad(_p(g), _p(db), g.numel() // N, N, dtype_code(g.dtype), _stream(g)))
    return db


def lowrank2_fused(x, wv, wh, bias, relu=False, pool=False):
    """x (B,H,W,c_pad) bf16 NHWC -> (B,H,W,N) bf16 (or (B,H/2,W/2,N) with the fused 2x2 max-pool); stride 1, 'same'."""
    _ready(x)
    B, H, W, cp = x.shape
    d, kp, cp2 = wv.shape
    N = wh.shape[1]
    assert cp == cp2 and x.dtype == torch.bfloat16 and x.is_contiguous()
    y = torch.empty((B, H // 2, W // 2, N) if pool else (B, H, W, N), device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _check(lib().cna_lowrank2_fused_forward(_p(x), _p(wv), _p(wh), _p(bias), _p(y), B, H, W, cp, kp, N, d,
                                                int(relu), int(pool), _stream(x)))
    return y


# ------------------------------------------------------------------------------------------- #
# fitting step
# ------------------------------------------------------------------------------------------- #
def svd_small(A: torch.Tensor):
    """Batched SVD of small fp32 matrices A (..., m, n), min(m,n) <= max(m,n) <= 64 -> (U, S, Vh), thin."""
    _ready(A)
    *lead, m, n = A.shape
    Af = A.detach().to(torch.float32).reshape(-1, m, n)
    if m < n:   # A = U S Vh  <=>  A^T = Vh^T S U^T
        U2, S, Vh2 = svd_small(Af.transpose(1, 2).contiguous())
        return (Vh2.transpose(1, 2).reshape(*lead, m, m), S.reshape(*lead, m), U2.transpose(1, 2).reshape(*lead, m, n))
    Af = Af.contiguous()
    nb = Af.shape[0]
    U = torch.empty((nb, m, n), device=A.device, dtype=torch.float32)
    S = torch.empty((nb, n), device=A.device, dtype=torch.float32)
    Vh = torch.empty((nb, n, n), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        _check(lib().cna_svd_small_batched(_p(Af), _p(U), _p(S), _p(Vh), nb, m, n, _stream(A)))
    return U.reshape(*lead, m, n), S.reshape(*lead, n), Vh.reshape(*lead, n, n)


def lowrank1_svd_fit(W2d: torch.Tensor, M: int):
    """W2d (R, q) fp32 -> weights (R, M), bases (M, q)."""
    _ready(W2d)
    R, q = W2d.shape
    Wf = _f32c(W2d)
    weights = torch.empty((R, M), device=W2d.device, dtype=torch.float32)
    bases = torch.empty((M, q), device=W2d.device, dtype=torch.float32)
    with torch.cuda.device(W2d.device):
        _check(lib().cna_lowrank1_svd_fit(_p(Wf), R, q, M, _p(weights), _p(bases), _stream(W2d)))
    return weights, bases


def lowrank1_alt_fit(W2d, d, weights, bases, lmda_list, max_iter, epsilon=1e-3, mm_rounds=6, fista_iters=40):
    """Alternating nuclear-norm fit on the device (cna_lowrank1_alt_fit).  W2d (R, d*d); weights (R, M) and bases (M, d*d) are the
    starting point and are updated IN PLACE.  Returns the objective trace (after every B-step and A-step)."""
    _ready(W2d)
    R, q = W2d.shape
    M = bases.shape[0]
    assert q == d * d and weights.shape == (R, M) and bases.shape == (M, q)
    for t in (W2d, weights, bases):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.device == W2d.device
    lm = (ctypes.c_double * len(lmda_list))(*[float(v) for v in lmda_list])
    cap = 2 * len(lmda_list) * max_iter
    trace = (ctypes.c_double * cap)()
    n = c_int(0)
    with torch.cuda.device(W2d.device):
        _check(lib().cna_lowrank1_alt_fit(_p(W2d), R, int(d), M, _p(weights), _p(bases), ctypes.cast(lm, c_void_p), len(lmda_list),
                                          int(max_iter), float(epsilon), int(mm_rounds), int(fista_iters), ctypes.cast(trace, c_void_p),
                                          cap, ctypes.cast(ctypes.byref(n), c_void_p), _stream(W2d)))
    return [trace[i] for i in range(n.value)]


def msca_equiv_kernel(w1s, b1s, w2s, b2s):
    """Lists over branches of conv1.weight (C,1,1,k), conv1.bias (C,)|None, conv2.weight (C,1,k,1), conv2.bias."""
    _ready(w1s[0])
    dev = w1s[0].device
    C = w1s[0].shape[0]
    ks = [int(w.shape[-1]) for w in w1s]
    K = max(ks)
    w1cat = torch.cat([_f32c(w).reshape(-1) for w in w1s])
    w2cat = torch.cat([_f32c(w).reshape(-1) for w in w2s])
    zero = torch.zeros(C, device=dev, dtype=torch.float32)
    b1 = torch.stack([zero if b is None else _f32c(b) for b in b1s]).contiguous()
    b2 = torch.stack([zero if b is None else _f32c(b) for b in b2s]).contiguous()
    weight = torch.empty((C, 1, K, K), device=dev, dtype=torch.float32)
    bias = torch.empty((C,), device=dev, dtype=torch.float32)
    res = torch.empty((2, C, K // 2), device=dev, dtype=torch.float32)
    ks_arr = (c_int * len(ks))(*ks)
    with torch.cuda.device(dev):
        _check(lib().cna_msca_equiv_kernel(_p(w1cat), _p(b1), _p(w2cat), _p(b2), ctypes.cast(ks_arr, c_void_p),
                                           len(ks), C, _p(weight), _p(bias), _p(res), _stream(weight)))
    return weight, bias, res


def svd_topk(A: torch.Tensor, r: int, max_sweeps: int = 30):
    """r leading singular triplets of a large fp32 matrix (either orientation) -> U (m,r), S (r), Vh (r,n)."""
    _ready(A)
    m, n = A.shape
    if n > m:
        U2, S, Vh2 = svd_topk(A.t().contiguous(), r, max_sweeps)
        return Vh2.t().contiguous(), S, U2.t().contiguous()
    Af = _f32c(A)
    U = torch.empty((m, r), device=A.device, dtype=torch.float32)
    S = torch.empty((r,), device=A.device, dtype=torch.float32)
    Vh = torch.empty((r, n), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        _check(lib().cna_svd_jacobi_large(_p(Af), m, n, r, _p(U), _p(S), _p(Vh), max_sweeps, _stream(A)))
    return U, S, Vh


# ------------------------------------------------------------------------------------------- #
# reconstruction step
# ------------------------------------------------------------------------------------------- #
def l2_recon(y_new, y_old, scale, want_grad=True, db=None):
    """-> (norms (B,) fp32, grad like y_new or None): per-sample ||y_new - y_old||_2 and d(scale*sum_b norm_b)/dy_new.
    `db` (fp32, (N,)): also filled with the bias gradient sum_p grad[p][n] (N = last dim of the NHWC y)."""
    _ready(y_new)
    assert y_new.shape == y_old.shape and y_new.dtype == y_old.dtype and y_new.is_contiguous() and y_old.is_contiguous()
    Bn = y_new.shape[0]
    per = y_new.numel() // Bn
    scratch = torch.empty((Bn,), device=y_new.device, dtype=torch.float32)
    norms = torch.empty((Bn,), device=y_new.device, dtype=torch.float32)
    g = torch.empty_like(y_new) if want_grad else None
    with torch.cuda.device(y_new.device):
        if db is not None:
            assert want_grad and db.dtype == torch.float32 and db.numel() == y_new.shape[-1]
            _check(lib().cna_l2_recon_loss_grad_bias(_p(y_new), _p(y_old), _p(scratch), _p(norms), _p(g), _p(db), y_new.shape[-1], Bn,
                                                     per, float(scale), dtype_code(y_new.dtype), _stream(y_new)))
        else:
            _check(lib().cna_l2_recon_loss_grad(_p(y_new), _p(y_old), _p(scratch), _p(norms), _p(g), Bn, per, float(scale),
                                                dtype_code(y_new.dtype), _stream(y_new)))
    return norms, g


def l2_norm_backward(y_new, y_old, norms, upstream):
    """g[b] = (y_new[b] - y_old[b]) * upstream[b] / norms[b] (0 where the norm is 0); same memory layout as y_new.
    y_old=None reads as zero, norms=None as one: g[b] = y_new[b] * upstream[b]."""
    _ready(y_new)
    assert y_new.is_contiguous() and (y_old is None or (y_new.shape == y_old.shape and y_new.dtype == y_old.dtype and y_old.is_contiguous()))
    Bn = y_new.shape[0]
    up = upstream.detach().to(torch.float32).contiguous()
    assert (norms is None or (norms.dtype == torch.float32 and norms.numel() == Bn)) and up.numel() == Bn
    g = torch.empty_like(y_new)
    with torch.cuda.device(y_new.device):
        _check(lib().cna_l2_norm_backward(_p(y_new), _p(y_old), _p(norms), _p(up), _p(g), Bn, y_new.numel() // Bn,
                                          dtype_code(y_new.dtype), _stream(y_new)))
    return g


def l2_recon_into(y_new, y_old, scale, sumsq, norms, g, db):
    """Loss / gradient / bias-gradient kernels of `l2_recon` on caller-owned buffers, WITHOUT zeroing `sumsq` (B,) and `db` (N,):
    a multi-layer step zeroes them with one memset for all layers.  Capturable in a CUDA graph (no allocation)."""
    _ready(y_new)
    Bn = y_new.shape[0]
    per = y_new.numel() // Bn
    assert y_new.shape == y_old.shape == g.shape and y_new.is_contiguous() and y_old.is_contiguous() and g.is_contiguous()
    assert sumsq.numel() == Bn and norms.numel() == Bn and db.dtype == torch.float32 and db.numel() == y_new.shape[-1]
    with torch.cuda.device(y_new.device):
        _check(lib().cna_l2_recon_loss_grad_bias_acc(_p(y_new), _p(y_old), _p(sumsq), _p(norms), _p(g), _p(db), y_new.shape[-1], Bn,
                                                     per, float(scale), dtype_code(y_new.dtype), _stream(y_new)))
    return norms


def conv_taps_wgrad(x, g, taps, axis, pad, dw_out=None, db_out=None, want_bias=False):
    """x (B,H,W,C), g (B,H,W,N) NHWC -> dw (N,C,taps) fp32 [, db (N,) fp32] for a stride-1 'same' tap conv."""
    _ready(x)
    Bn, H, W, C = x.shape
    N = g.shape[3]
    assert g.shape[:3] == x.shape[:3] and g.dtype == x.dtype and x.is_contiguous() and g.is_contiguous()
    dw = torch.empty((N, C, taps), device=x.device, dtype=torch.float32) if dw_out is None else dw_out
    db = None
    if want_bias or db_out is not None:
        db = torch.empty((N,), device=x.device, dtype=torch.float32) if db_out is None else db_out
    with torch.cuda.device(x.device):
        _check(lib().cna_conv_taps_wgrad(_p(x), _p(g), _p(dw), _p(db), Bn, H, W, C, N, taps, axis, pad,
                                         dtype_code(x.dtype), _stream(x)))
    return dw, db


def pack_taps_dgrad(w, dtype):
    """(O,I,kh,kw) tap-conv weight -> packed weight of the transposed conv: [taps][O][I] with flipped taps."""
    _ready(w)
    O, I, kh, kw = w.shape
    taps = kh * kw
    out = torch.empty((taps, O, I), device=w.device, dtype=dtype)
    wf = _f32c(w)
    with torch.cuda.device(w.device):
        _check(lib().cna_pack_taps_weight_dgrad(_p(wf), _p(out), O, I, taps, dtype_code(dtype), _stream(w)))
    return out


def adamw_step_(param, grad, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                grad_scale=1.0, n_decay=None):
    """Weight decay covers the first `n_decay` elements (default: all) — biases go last and get none, as in the reference's
    timm-built optimizer."""
    _ready(param)
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.dtype == torch.float32 and t.is_contiguous() and t.numel() == param.numel()
    nd = param.numel() if n_decay is None else int(n_decay)
    with torch.cuda.device(param.device):
        _check(lib().cna_adamw_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), nd, float(lr),
                                    float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step),
                                    float(grad_scale), _stream(param)))
    return param
