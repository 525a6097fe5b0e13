"""Overlay (drop-in) mode: with CNA_REFERENCE_SRC set, the reference's own model / filter / config code runs on top of
this backend's `approx.layers` and `approx.core`.  Needs the reference tree, which exists only in the build container
(skipped elsewhere); CPU only, so it checks structure — registration, substitution, state_dict keys — and that the
layers in place are this backend's (their forward refuses CPU tensors)."""
import os
import subprocess
import sys

import pytest

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, types
for name in ("timm", "timm.data", "timm.models", "timm.utils", "timm.loss", "timm.optim", "timm.scheduler", "ptflops"):
    sys.modules.setdefault(name, types.ModuleType(name))
import torch
import approx.layers, approx.core
from approx.models import build_model, MODEL
from approx.core import build_app
from approx.filters import build_filter
from approx.utils import init_cfg, get_cfg
assert approx.layers.__file__.startswith(sys.argv[1]), approx.layers.__file__
import approx.models.switchable as sw
assert sw.__file__.startswith("/root/reference"), sw.__file__
init_cfg("/root/reference/configs/msca-rep/msca-rep_d1_fix_mscan-t.py")
cfg = get_cfg()
model = build_model(dict(cfg.model, init_cfg=None))
app = build_app(cfg.app, deploy=False)
filters = [build_filter(f) for f in (cfg.filters or [])]
model.register_switchable(app.src_type, filters)
assert model.length_switchable == 13
for idx in range(model.length_switchable):
    src = model.get_switchable_module(idx)
    model.set_switchable_module(idx, app.initialize, src=src)
sub = model.get_switchable_module(0)
assert type(sub).__module__ == "approx.layers.substituton"
keys = [k for k in model.state_dict() if ".new.sd_convs" in k and "block.0" in k or ".new.sd_convs" in k]
assert any(k.endswith("new.sd_convs.0.conv1.weight") for k in keys) and any(k.endswith("new.sd_convs.1.res") for k in keys)
try:
    sub.new(torch.randn(1, 32, 8, 8))
    raise SystemExit("expected the backend to refuse a CPU tensor")
except RuntimeError as e:
    assert "no CPU fallback" in str(e)
# low-rank config through the reference's AlexNet + SimpleConvFilter + IndicesFilter
init_cfg("/root/reference/configs/low-rank-exp/low-rank-exp-v1_l2345_svd_dodecomp_alexnet.py")
cfg = get_cfg()
model = build_model(dict(cfg.model, init_cfg=None))
app = build_app(cfg.app, deploy=False)
filters = [build_filter(f) for f in cfg.filters]
model.register_switchable(app.src_type, filters)
assert model.length_switchable == 4
for idx in range(4):
    model.set_switchable_module(idx, app.initialize, src=model.get_switchable_module(idx))
names = [type(model.get_switchable_module(i).new).__name__ for i in range(4)]
assert names == ["LowRankExpConvV1"] * 4
print("overlay ok")
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_reference_models_and_configs_drive_this_backend():
    env = dict(os.environ, CNA_REFERENCE_SRC=REF, PYTHONPATH=os.path.join(ROOT, "convnet-approximater_b200"))
    pkg = os.path.join(ROOT, "convnet-approximater_b200")
    out = subprocess.run([sys.executable, "-c", SCRIPT, pkg], env=env, capture_output=True, text=True, cwd="/tmp", timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "overlay ok" in out.stdout


REF_SCRIPT = r"""
import sys, types
for name in ("timm", "timm.data", "timm.models", "timm.utils", "timm.loss", "timm.optim", "timm.scheduler", "ptflops", "cvxpy"):
    sys.modules.setdefault(name, types.ModuleType(name))
import torch
import approx.layers, approx.core
from approx.models import build_model
from approx.core import build_app
from approx.filters import build_filter
from approx.utils import init_cfg, get_cfg
mode, cfg_file, ckpt_in, ckpt_out = sys.argv[1:5]
init_cfg(cfg_file)
cfg = get_cfg()
torch.manual_seed(0)
model = build_model(dict(cfg.model, init_cfg=None))
app = build_app(cfg.app, deploy=False)
filters = [build_filter(f) for f in (cfg.filters or [])]
model.register_switchable(app.src_type, filters)
for idx in range(model.length_switchable):
    model.set_switchable_module(idx, app.initialize, src=model.get_switchable_module(idx))
post = type(app).__name__ != "LowRankExpV1"     # LowRankExpV1.optimize needs cvxpy even for its SVD branch; decomp() here needs the GPU
if mode == "reference":
    assert approx.layers.__file__.startswith("/root/reference"), approx.layers.__file__
    for sub in model.switchable_modules():
        if post:
            app.optimize(sub)                   # the reference's own CPU fit (torch.linalg.svd)
else:
    assert not approx.layers.__file__.startswith("/root/reference"), approx.layers.__file__
if ckpt_in != "-":
    sd = torch.load(ckpt_in, map_location="cpu")["state_dict"]
    missing, unexpected = model.load_state_dict(sd, strict=True)     # same keys, same shapes, both directions
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
torch.save(dict(state_dict=model.state_dict()), ckpt_out)       # before post-processing: what the other side loads
for idx in range(model.length_switchable if post else 0):       # Runner.run's post-process step (Substitution -> replacement layer)
    model.set_switchable_module(idx, app.postprocess, sub=model.get_switchable_module(idx))
torch.save(dict(state_dict=model.state_dict()), ckpt_out + ".post")
print("ckpt ok", len(model.state_dict()))
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
@pytest.mark.parametrize("cfg", ["configs/msca-rep/msca-rep_d1_fix_mscan-t.py",
                                 "configs/low-rank-exp/low-rank-exp-v1_l2345_svd_dodecomp_alexnet.py"])
def test_checkpoints_interchange_with_the_reference(cfg, tmp_path):
    """A checkpoint written by the unmodified reference (its layers, its CPU fit) loads strictly into the same model built on
    this backend's layers, and the post-processed checkpoints of both carry identical keys, shapes and values."""
    import torch
    pkg = os.path.join(ROOT, "convnet-approximater_b200")
    cfg_file = os.path.join("/root/reference", cfg)
    ref_ckpt, our_ckpt = str(tmp_path / "ref.pth"), str(tmp_path / "ours.pth")

    def run(mode, ckpt_in, ckpt_out):
        if mode == "reference":
            env = dict(os.environ, PYTHONPATH=REF)
            env.pop("CNA_REFERENCE_SRC", None)
        else:
            env = dict(os.environ, CNA_REFERENCE_SRC=REF, PYTHONPATH=pkg)
        script = REF_SCRIPT
        out = subprocess.run([sys.executable, "-c", script, mode, cfg_file, ckpt_in, ckpt_out], env=env, capture_output=True,
                             text=True, cwd="/tmp", timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
        assert "ckpt ok" in out.stdout

    run("reference", "-", ref_ckpt)
    run("overlay", ref_ckpt, our_ckpt)
    a = torch.load(ref_ckpt + ".post", map_location="cpu")["state_dict"]
    b = torch.load(our_ckpt + ".post", map_location="cpu")["state_dict"]
    assert list(a) == list(b)
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


CONVERTER_SCRIPT = r"""
import runpy, sys, types
for name in ("timm", "timm.data", "timm.models", "timm.utils", "timm.loss", "timm.optim", "timm.scheduler", "ptflops", "cvxpy"):
    sys.modules.setdefault(name, types.ModuleType(name))
import torch
import approx.layers
mode, script, cfg_file, ckpt = sys.argv[1:5]
assert approx.layers.__file__.startswith("/root/reference") == (mode == "reference"), approx.layers.__file__
if script == "make":
    from approx.models import build_model
    from approx.core import build_app
    from approx.filters import build_filter
    from approx.utils.config import Config
    cfg = Config(); cfg.load_from_file(cfg_file)
    torch.manual_seed(0)
    model = build_model(cfg.model)
    app = build_app(cfg.app, deploy=False)
    filters = [build_filter(f) for f in cfg.filters] if 'filters' in cfg else []
    model.register_switchable(app.src_type, filters)
    for idx in range(model.length_switchable):
        model.set_switchable_module(idx, app.initialize, src=model.get_switchable_module(idx))
    torch.save(dict(state_dict=model.state_dict()), ckpt)
else:
    sys.argv = [script, "--config", cfg_file, "--ckpt", ckpt]
    runpy.run_path(script, run_name="__main__")        # the reference's converter, unmodified
print("converter ok")
"""


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
@pytest.mark.parametrize("cfg", ["configs/msca-rep/msca-rep_d1_fix_mscan-t.py",
                                 "configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py"])   # (do_decomp needs the GPU)
def test_reference_checkpoint_converters_run_on_this_backend(cfg, tmp_path):
    """The reference's `scripts/ckpt_converter/remove_substitution.py` and `add_substitution.py`, run
    unmodified on top of this backend's layers / apps (`deploy=True` and `postprocess` included): Substitution checkpoint →
    deployed checkpoint → Substitution checkpoint.  The key lists equal the ones the unmodified reference produces, and the
    replacement layers' tensors survive the round trip bit for bit."""
    import torch
    pkg = os.path.join(ROOT, "convnet-approximater_b200")
    cfg_file = os.path.join("/root/reference", cfg)
    conv = "/root/reference/scripts/ckpt_converter/"

    def run(mode, script, ckpt):
        if mode == "reference":
            env = dict(os.environ, PYTHONPATH=REF)
            env.pop("CNA_REFERENCE_SRC", None)
        else:
            env = dict(os.environ, CNA_REFERENCE_SRC=REF, PYTHONPATH=pkg)
        out = subprocess.run([sys.executable, "-c", CONVERTER_SCRIPT, mode, script, cfg_file, ckpt], env=env,
                             capture_output=True, text=True, cwd=str(tmp_path), timeout=600)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
        assert "converter ok" in out.stdout

    got = {}
    for mode in ("reference", "overlay"):
        ckpt = str(tmp_path / f"{mode}.pth")
        run(mode, "make", ckpt)
        run(mode, conv + "remove_substitution.py", ckpt)
        run(mode, conv + "add_substitution.py", str(tmp_path / f"{mode}_rm-sub.pth"))
        got[mode] = [torch.load(str(tmp_path / (mode + sfx)), map_location="cpu")["state_dict"]
                     for sfx in (".pth", "_rm-sub.pth", "_rm-sub_add-sub.pth")]
    for ref_sd, our_sd in zip(got["reference"], got["overlay"]):
        assert list(ref_sd) == list(our_sd)
        assert all(ref_sd[k].shape == our_sd[k].shape for k in ref_sd)
    sub, deployed, back = got["overlay"]
    assert not any(".new." in k or ".old." in k for k in deployed)
    assert list(sub) == list(back)
    new_keys = [k for k in sub if ".new." in k]
    assert new_keys
    for k in new_keys:
        assert torch.equal(sub[k], deployed[k.replace(".new.", ".")]) and torch.equal(sub[k], back[k]), k
