"""CPU-side checks: the C-ABI library loads and exports every symbol include/cna_b200.h declares, the
host-side mirror keeps the reference's registry names / constructor signatures / state_dict keys, and
the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from approx import backend
    hdr = open(os.path.join(ROOT, "include", "cna_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(cna_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 15
    lib = ctypes.CDLL(backend.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cna_b200.h but not exported"
    assert set(backend.exported_symbols()) <= set(declared)
    assert backend.lib().cna_version() >= 100


def test_registries_hold_the_reference_names():
    from approx.layers import LAYER
    from approx.core import APP
    for n in ("LowRankExpConvV1", "LowRankExpConvV2", "MSCA", "MSCAProfile", "SimpleConv", "DummyLayer"):
        assert n in LAYER
    for n in ("LowRankExpV1", "MscaRep", "MscaProfile", "MscaRepProfile", "Dummy", "LowRankExpV2"):
        assert n in APP
    with pytest.raises(AssertionError):
        LAYER.get("nope")
    with pytest.raises(AssertionError):           # double registration asserts, like the reference
        LAYER.register_module(name="MSCA", module=nn.Identity)


def test_build_from_cfg_behaviour():
    from approx.layers import build_layer
    from approx.core import build_app
    m = build_layer(dict(type="MSCA", num_channel=8, k1_size=5, k_sizes=(7, 11, 21)))
    assert m.num_channel == 8 and m.switchable_layer() is m.sd_convs
    app = build_app(dict(type="MscaRep", decomp=1, fix=True), deploy=False)
    assert app.decomp == 1 and app.fix and app.src_type is type(m)
    with pytest.raises(TypeError):
        build_layer(dict(type="MSCA", bogus=1))
    assert build_layer(None) is None


def test_state_dict_keys_match_reference(golden):
    from approx.layers import LowRankExpConvV1, LowRankExpConvV2, MSCA
    from approx.core import MscaRep
    g = golden("lowrank1")
    keys = lambda prefix, gg: sorted(k[len(prefix):] for k in gg if k.startswith(prefix) and "." in k[len(prefix):]
                                      and not k[len(prefix):].startswith("dec_"))
    assert sorted(LowRankExpConvV1(8, 12, 3, 1, 1, num_base=4).state_dict()) == keys("a_", g)
    assert sorted(LowRankExpConvV1(8, 12, 3, 1, 1, num_base=4, decomp=True).state_dict()) == \
        sorted(k[len("a_dec_"):] for k in g if k.startswith("a_dec_"))
    g2 = golden("lowrank2_shipped")
    m2 = LowRankExpConvV2(16, 8, (3, 3), (1, 1), (1, 1))
    assert sorted(m2.state_dict()) == keys("a_", g2)
    assert tuple(m2.h_conv.weight.shape) == tuple(g2["a_h_conv.weight"].shape)        # depthwise, as shipped
    assert tuple(LowRankExpConvV2(16, 8, (3, 3), (1, 1), (1, 1), out_channels=32).h_conv.weight.shape) == (32, 8, 1, 3)
    gm = golden("msca")
    src = MSCA(8, 5, (7, 11, 21))
    assert sorted(src.state_dict()) == keys("src_", gm)
    for decomp, fix in [(0, True), (0, False), (1, True), (2, True), (4, True), (3, False)]:
        tag = f"d{decomp}{'f' if fix else 'n'}_"
        sub = MscaRep(decomp=decomp, fix=fix).initialize(src)      # structure only: no device needed
        assert sorted(sub.new.state_dict()) == keys(tag, gm), tag
        assert sorted(sub.state_dict()) == sorted([f"old.{k}" for k in src.state_dict()] +
                                                  [f"new.{k}" for k in sub.new.state_dict()])
    assert isinstance(MscaRep(decomp=1, fix=True, deploy=True).initialize(src), MSCA)


def test_same_seed_same_initialisation_as_reference(golden):
    """Parameter containers are stock nn.Conv2d built in the reference's order, so a seed reproduces its init."""
    from approx.layers import LowRankExpConvV2
    torch.manual_seed(7)
    m = LowRankExpConvV2(16, 8, (3, 3), (1, 1), (1, 1)).to(torch.bfloat16)
    g = golden("bf16")
    assert np.array_equal(m.v_conv.weight.detach().float().numpy(), g["v2_v_conv.weight"])
    assert np.array_equal(m.h_conv.bias.detach().float().numpy(), g["v2_h_conv.bias"])


def test_num_bases_cursor_and_rewind():
    from approx.core import LowRankExpV1
    app = LowRankExpV1(num_bases=(4, 6), max_iter=0, lmda_length=1, min_lmda=0.001, max_lmda=0.001)
    a = app.initialize(nn.Conv2d(4, 8, 3, 1, 1)); b = app.initialize(nn.Conv2d(8, 8, 5, 1, 2))
    assert a.new.num_base == 4 and b.new.num_base == 6 and b.new.kernel_size == (5, 5)
    app.rewind()
    assert app.initialize(nn.Conv2d(4, 8, 3, 1, 1)).new.num_base == 4
    assert a.new.bias is a.new.d_conv.bias and torch.equal(a.new.bias.data, a.old.bias.data)
    with pytest.raises(AssertionError):
        app.initialize(nn.Linear(2, 2))


def test_no_cpu_fallback():
    from approx.layers import LowRankExpConvV2, MSCA, CascadeConv, FixPaddingBias
    x = torch.randn(1, 8, 8, 8)
    for m in (MSCA(8, 5, (7, 11, 21)), LowRankExpConvV2(8, 4, (3, 3), (1, 1), (1, 1), out_channels=8),
              CascadeConv(8, 7, 3, True, True), FixPaddingBias(8, 3)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(x)
    if not torch.cuda.is_available():
        from approx.core import MscaRep
        app = MscaRep(decomp=0, fix=True)
        sub = app.initialize(MSCA(8, 5, (7, 11, 21)))
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            app.optimize(sub)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "convnet-approximater_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|approx_oracle", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert not pat.search(open(os.path.join(dp, f)).read()), (dp, f)


def test_lowrank1_pack_layout_cpu():
    """Host-side packing of the scheme-1 basis filters (pure tensor re-layout, no kernel): per 64 wide channels
    [rows][64] with unit u's channels 8u..8u+3 at floats [4u,4u+4) and 8u+4..8u+7 at [32+4u, 32+4u+4)."""
    import torch
    from approx import backend as B
    from approx.layers import LowRankExpConvV1
    torch.manual_seed(0)
    C, M, d = 16, 8, 3
    layer = LowRankExpConvV1(C, 32, d, 1, 1, num_base=M)
    sp = B.lowrank1_pack_s(layer.s_conv, M)
    assert sp.shape == (C * M // 64, d * d + 1, 64) and sp.dtype == torch.float32
    w = layer.s_conv.weight.detach().to(torch.bfloat16).float().reshape(C * M, d * d)
    for cm in (0, 5, 63, 64, 77, 127):
        chunk, local = divmod(cm, 64)
        u, e = divmod(local, 8)
        pos = (e // 4) * 32 + u * 4 + (e % 4)
        assert torch.equal(sp[chunk, :d * d, pos], w[cm])
        assert sp[chunk, d * d, pos] == layer.s_conv.bias.detach()[cm]
    layer.s_conv = __import__("approx.layers.low_rank_conv", fromlist=["SeparableConv"]).SeparableConv(C, M, (d, d), (1, 1), (1, 1))
    sp = B.lowrank1_pack_s(layer.s_conv, M)
    assert sp.shape == (C * M // 64, 2 * d, 64)
    v = layer.s_conv.v_conv.weight.detach().to(torch.bfloat16).float().reshape(C * M, d)
    h = layer.s_conv.h_conv.weight.detach().to(torch.bfloat16).float().reshape(C * M, d)
    cm = 70; u, e = divmod(cm - 64, 8); pos = (e // 4) * 32 + u * 4 + (e % 4)
    assert torch.equal(sp[1, :d, pos], v[cm]) and torch.equal(sp[1, d:, pos], h[cm])


def test_lowrank1_fused_plan_cpu():
    """The fused scheme-1 kernel's host-side plan (tile shape + shared-memory budget) is pure host code: the AlexNet layers of the
    reference's config are covered, everything outside the documented envelope is refused (callers then take the unfused kernels)."""
    from approx import backend as B
    for (C, N, d, H) in [(64, 192, 5, 27), (192, 384, 3, 13), (384, 256, 3, 13), (256, 256, 3, 13)]:
        for sep in (False, True):
            assert B.lowrank1_fused_supported(64, H, H, C, N, 8, d, sep), (C, N, d, H, sep)
    assert B.lowrank1_fused_supported(1, 7, 128, 8, 16, 8, 3, False)          # W = 128: one row per tile
    assert not B.lowrank1_fused_supported(1, 7, 129, 8, 16, 8, 3, False)      # wider than a tile
    assert not B.lowrank1_fused_supported(4, 13, 13, 64, 192, 8, 7, False)    # 7x7 bases: no instance
    assert B.lowrank1_fused_supported(4, 13, 13, 64, 192, 16, 3, False)       # M = 16: 4 channels per chunk, staged as aligned octets
    assert not B.lowrank1_fused_supported(4, 13, 13, 64, 192, 32, 3, False)   # M = 32: no instance
    assert B.lowrank1_fused_supported(4, 13, 13, 64, 192, 2, 5, True)         # M = 2: 32 channels per chunk
    assert not B.lowrank1_fused_supported(4, 13, 13, 48, 192, 2, 3, False)    # ... which must divide C
    assert not B.lowrank1_fused_supported(4, 13, 13, 12, 192, 16, 3, False)   # C % 8 != 0
    assert not B.lowrank1_fused_supported(4, 13, 13, 12, 192, 8, 3, False)    # C % 8 != 0
    assert not B.lowrank1_fused_supported(4, 13, 13, 64, 200, 8, 3, False)    # N % 16 != 0
    assert not B.lowrank1_fused_supported(4, 13, 13, 64, 528, 8, 3, False)    # N > 512
    assert not B.lowrank1_fused_supported(4, 13, 13, 64, 272, 8, 3, False)    # 256 < N, N % 32 != 0


def test_lowrank2_fused_plan_cpu():
    """The fused scheme-2 kernel's shared-memory plan is pure host code: every VGG-16 layer at every rank bucket fits the 227 KB opt-in
    limit with rings deep enough to pipeline, everything stays resident where it can, and partial residency (one of Wv / Wh resident,
    the other streamed) appears exactly where it was measured to pay (<= 2 channel chunks or >= 4 N-tiles)."""
    from approx import backend as B
    vgg = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 512), (512, 512), (512, 512),
           (512, 512), (512, 512), (512, 512)]
    pooled = {1, 3, 6, 9, 12}
    for K in (4, 16, 32, 64):
        for i, (C, N) in enumerate(vgg):
            pl = B.lowrank2_fused_plan(C, K, N, pool=i in pooled)
            assert pl["smem_bytes"] <= 232448, (K, i, pl)
            assert pl["x_stages"] >= 2 and 1 <= pl["store_buffers"] <= 3, (K, i, pl)
            assert pl["wh_resident"] or pl["wh_slots"] >= 3, (K, i, pl)
            assert pl["channel_chunk"] == (16 if C == 3 else 64) and pl["n_tile"] == (64 if N == 64 else 128)
            if pl["wv_resident"] and pl["wh_resident"]:
                assert pl["x_stages"] >= 3, (K, i, pl)
    full = lambda C, N, pool=False: B.lowrank2_fused_plan(C, 64, N, pool)
    assert all(full(C, N)["wv_resident"] and full(C, N)["wh_resident"] for (C, N) in vgg[:4])       # 64- / 128-channel layers
    p5, p6, p8, p9 = full(128, 256), full(256, 256), full(256, 512), full(512, 512)
    assert (p5["wv_resident"], p5["wh_resident"]) == (0, 1) and (p8["wv_resident"], p8["wh_resident"]) == (1, 0)
    assert (p6["wv_resident"], p6["wh_resident"]) == (0, 0) and (p9["wv_resident"], p9["wh_resident"]) == (0, 0)
    assert full(3, 64)["store_buffers"] == 3                                                        # layer 1: smem to spare
    with pytest.raises(RuntimeError):
        B.lowrank2_fused_plan(64, 64, 72)                                                           # N % 64 != 0
    with pytest.raises(RuntimeError):
        B.lowrank2_fused_plan(64, 64, 64, smem_limit=48 * 1024)                                     # no opt-in shared memory


def test_hook_registry_holds_the_reference_names_and_signatures():
    """approx.hooks: the names the shipped configs use, with the reference's constructor arguments (src/approx/hooks/*.py)."""
    import inspect
    from approx.hooks import HOOK, Hook, build_hook, get_priority, Priority
    want = {
        "InferenceTimeHook": ["runner", "priority", "infer_cfg"],
        "Fps": ["runner", "priority", "repeat_times", "log_interval", "num_workers", "use_torch_benchmark", "dataset_args", "data_config"],
        "L2Reconstruct": ["runner", "priority", "asym", "l2_weight", "cls_weight", "epoch_behavior", "no_norm", "dataset_args",
                          "optim_args", "sche_args", "data_config", "other_args"],
        "LowRankExpV1Decomp": ["runner", "priority"],
        "CkptHook": ["runner", "priority", "ckpt_cfg"],
        "ModelAnalysis": ["runner", "priority", "input_shape"],
    }
    for name, args in want.items():
        cls = HOOK.get(name)
        assert issubclass(cls, Hook)
        got = list(inspect.signature(cls.__init__).parameters)[1:]
        assert got[:len(args)] == args, (name, got)
    assert Hook.stages == ('before_run', 'after_register', 'after_initialize', 'after_optimize', 'after_run')
    assert get_priority("normal") == 50 and get_priority(Priority.HIGH) == 30 and get_priority(7) == 7
    import pytest
    with pytest.raises(ValueError):
        get_priority(101)

    class R:        # minimal runner stand-in
        model = None
    h = build_hook(dict(type="LowRankExpV1Decomp", priority="LOW"), runner=R())
    assert h.priority == 70 and h.name == "LowRankExpV1Decomp"
    from approx.utils.general import is_method_overridden
    assert is_method_overridden("after_run", Hook, h) and not is_method_overridden("before_run", Hook, h)


def test_synthetic_loader_and_lr_schedule():
    import torch
    from approx.hooks._common import SyntheticLoader
    from approx.hooks.finetune import _schedule
    from approx.hooks._common import ArgTable, DEFAULT_SCHED_ARGS
    ld = SyntheticLoader(3, (3, 8, 8), num_batches=2, num_classes=5, device="cpu")
    a = [(x.clone(), t.clone()) for x, t in ld]
    b = [(x.clone(), t.clone()) for x, t in ld]
    assert len(a) == len(ld) == 2 and a[0][0].shape == (3, 3, 8, 8) and int(a[0][1].max()) < 5
    assert all(torch.equal(p[0], q[0]) and torch.equal(p[1], q[1]) for p, q in zip(a, b))      # every epoch sees the same batches
    assert _schedule(ArgTable(DEFAULT_SCHED_ARGS), 1e-3)(7) == 1e-3
    cos = _schedule(ArgTable(DEFAULT_SCHED_ARGS, dict(sched='cosine', epochs=10, min_lr=1e-5)), 1e-3)
    assert abs(cos(0) - 1e-3) < 1e-12 and abs(cos(10) - 1e-5) < 1e-12 and cos(5) < cos(1)


def test_lowrank1_plan_record_host_logic(monkeypatch):
    """Host logic of `LowRankExpConvV1._fused` without a GPU: the kernel entry points are replaced by counters, the layer and its
    input stay on the CPU.  The per-layer record must reuse the packed operands call after call, pick the smallest basis instance
    whose chunks divide C*M, and re-plan after `.weight.data = ...`, `decomp()`, `_ops.invalidate` and a new input shape — and only then."""
    import torch
    from approx import backend as B
    from approx.layers import LowRankExpConvV1, _ops
    calls = {"pack_s": 0, "pack_d": 0, "supported": [], "fused": []}

    def pack_s(s_conv, M, m_pad=None):
        calls["pack_s"] += 1
        return torch.zeros(1)

    def pack_d(w, C, M, m_pad):
        calls["pack_d"] += 1
        return torch.zeros(1)

    def supported(Bn, H, W, C, N, M, d, sep):
        calls["supported"].append(M)
        return True

    def fused(x, sp, dp, bias, M, d, sep):
        calls["fused"].append((M, d, sep))
        return x
    monkeypatch.setattr(B, "lowrank1_pack_s", pack_s)
    monkeypatch.setattr(B, "lowrank1_pack_d", pack_d)
    monkeypatch.setattr(B, "lowrank1_fused_supported", supported)
    monkeypatch.setattr(B, "lowrank1_fused", fused)
    torch.manual_seed(0)
    layer = LowRankExpConvV1(64, 32, 3, 1, 1, num_base=3).bfloat16().eval()
    x = torch.zeros(2, 5, 5, 64, dtype=torch.bfloat16)                     # NHWC, as forward_nhwc receives it

    def run(n=3, inp=x):
        for _ in range(n):
            assert layer._fused(inp) is inp
    run()
    assert (calls["pack_s"], calls["pack_d"]) == (1, 1) and calls["supported"] == [4]      # 64*2 % 64 == 0 but 2 < 3 bases: instance 4
    assert calls["fused"] == [(4, 3, False)] * 3
    layer.d_conv.weight.data = layer.d_conv.weight.data.clone()                             # the reference's write-back style
    run()
    assert (calls["pack_s"], calls["pack_d"]) == (1, 2)                                     # only d_conv's pack is rebuilt
    layer.s_conv.bias.data.zero_()                                                          # in place through .data: invisible ...
    run()
    assert (calls["pack_s"], calls["pack_d"]) == (1, 2)
    _ops.invalidate(layer)                                                                  # ... until invalidate()
    run()
    assert (calls["pack_s"], calls["pack_d"]) == (2, 3)
    run(inp=torch.zeros(3, 5, 5, 64, dtype=torch.bfloat16))                                 # new shape: re-plan, packs reused
    assert (calls["pack_s"], calls["pack_d"]) == (2, 3) and calls["supported"] == [4, 4, 4, 4]      # first call, write-back, invalidate, new shape
    assert layer._fused(x.float()) is None and layer._fused(x.float()) is None              # fp32 input: not this kernel, decided once
    assert layer._fused(x[:, :, ::2]) is None                                               # strided view: not this kernel
    # decomp() swaps s_conv for a SeparableConv (needs the device SVD: stand in for it here)
    from approx.layers.low_rank_conv import SeparableConv
    layer.s_conv = SeparableConv(64, 3, (3, 3), (1, 1), (1, 1)).bfloat16()
    n_fused = len(calls["fused"])
    run()
    assert calls["pack_s"] == 3 and calls["fused"][n_fused:] == [(4, 3, True)] * 3
    # smallest instance: 2 bases on a 32-channel-multiple layer, 16 for 9..16, none above
    for M, want in ((2, 2), (5, 8), (9, 16), (16, 16), (17, None)):
        lay = LowRankExpConvV1(64, 32, 3, 1, 1, num_base=M).bfloat16().eval()
        calls["fused"].clear()
        out = lay._fused(x)
        assert (out is None and want is None) or calls["fused"] == [(want, 3, False)], (M, want, calls["fused"])
    lay = LowRankExpConvV1(16, 32, 3, 1, 1, num_base=2).bfloat16().eval()                     # 16*2 = 32 wide channels: instance 4 (64)
    calls["fused"].clear()
    lay._fused(torch.zeros(2, 5, 5, 16, dtype=torch.bfloat16))
    assert calls["fused"] == [(4, 3, False)]


def test_empty_batch_wrapper_host_logic():
    """`_ops.empty_batch_safe` without a GPU: a batch of zero images is answered from one zero image pushed through the wrapped
    forward (shape and dtype of the result, no gradient graph), everything else — positional or keyword `x`, a zero-sized H —
    reaches the wrapped forward untouched."""
    from approx.layers import _ops
    seen = []

    class Layer(nn.Module):
        @_ops.empty_batch_safe
        def forward(self, x, scale=1):
            seen.append(tuple(x.shape))
            return x[:, :1, ::2, ::2].repeat(1, 5, 1, 1) * scale

    layer = Layer()
    import approx.backend as B
    orig = B.require_cuda
    B.require_cuda = lambda *a: None                      # the wrapper validates the empty input like any other
    try:
        y = layer(torch.zeros(0, 3, 8, 6))
        assert y.shape == (0, 5, 4, 3) and seen == [(1, 3, 8, 6)] and not y.requires_grad
        y = layer(torch.ones(2, 3, 8, 6), scale=2)
        assert y.shape == (2, 5, 4, 3) and float(y.max()) == 2.0 and seen[-1] == (2, 3, 8, 6)
        assert layer(x=torch.ones(1, 3, 4, 4)).shape == (1, 5, 2, 2)
        layer(torch.zeros(0, 3, 0, 6))                    # no image to imitate: handed through as it is
        assert seen[-1] == (0, 3, 0, 6)
        with pytest.raises(ValueError):
            _ops.check_input(torch.zeros(0, 3, 8))
    finally:
        B.require_cuda = orig
    assert layer.forward.__name__ == "forward"
