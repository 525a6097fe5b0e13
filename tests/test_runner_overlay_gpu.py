"""§8(f) rank 1 and 2 on the GPU: the reference's OWN driver code — `Runner.run` (runner/runner.py:40-81), its config loader, its
models (AlexNet, MSCAN_Classifier), filters and (de)serialiser — runs on top of this backend's `approx.layers`, `approx.core` and
`approx.hooks` (overlay mode, CNA_REFERENCE_SRC), with the shipped configs:

  1. register -> initialize -> optimize (device fit) -> postprocess -> save -> hooks, on CUDA, through the reference's Runner;
  2. the checkpoint it wrote is loaded by the UNMODIFIED reference (its own layers, `main.py --checkpoint` deploy flow) in a second
     process on the CPU, and the two models' outputs on the same input agree;
  3. `L2Reconstruct` (this backend's hook, reference semantics) trains the replaced AlexNet layers for a few steps and its loss
     curve / parameters follow a plain-torch restatement of the reference hook's arithmetic (hooks/finetune.py:293-331).

The reference tree is /root/reference in the build container and baseline/_ref (copied there by baseline/make_ref.py, git-ignored,
shipped with the gpurun snapshot) on the GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "convnet-approximater_b200")
REF = next((p for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")) if os.path.isdir(os.path.join(p, "src", "approx"))), None)
needs_ref = pytest.mark.skipif(REF is None, reason="no reference tree (run baseline/make_ref.py where /root/reference exists)")

STUBS = r'''
import sys, types
class _Any(types.ModuleType):
    """permissive stand-in for an uninstalled optional dependency of modules this flow never calls"""
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name.isupper():
            return (0.5, 0.5, 0.5)
        return lambda *a, **k: (0, 0)
for name in STUB_NAMES:
    sys.modules.setdefault(name, _Any(name))
'''

OVERLAY_RUN = STUBS.replace("STUB_NAMES", '("ptflops",)') + r'''
import os, torch
# the layers the reference model keeps as stock torch modules (stem, MLPs, classifier) must run exact fp32 for the comparison with
# the CPU reference: torch's GPU convs / matmuls default to TF32
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
ref, cfg_rel, work, mode = sys.argv[1:5]
ref_src = os.path.join(ref, "src")
import approx.layers, approx.core, approx.hooks
for m in (approx.layers, approx.core, approx.hooks):
    assert not m.__file__.startswith(ref_src), m.__file__                 # this backend's packages ...
from approx.runner import Runner
import approx.runner.runner, approx.models.switchable, approx.utils.config
for m in (approx.runner.runner, approx.models.switchable, approx.utils.config):
    assert m.__file__.startswith(ref_src), m.__file__                     # ... under the reference's own driver code
from approx.utils.config import init_cfg, update_cfg, get_cfg
from approx.utils import random_seed
from approx import backend as B
init_cfg(os.path.join(ref, cfg_rel))
cfg = get_cfg()
cfg.model['init_cfg'] = None                                              # no pretrained file on this machine: seeded random init
# dataset-bound hooks: no ImageNet / CIFAR on this machine -> ClassEvalHook and Fps read the synthetic loader at a small size
for h in (cfg.hooks or []):
    if h['type'] == 'ClassEvalHook':
        h['eval_cfg'].update(dataset='synthetic', batch_size=4, num_batches=2, input_size=(3, 96, 96), num_classes=10)
    if h['type'] == 'Fps':
        h.update(dataset_args=dict(name='synthetic', batch_size=4, num_batches=8), data_config=dict(input_size=(3, 96, 96)), total_iters=8)
    if h['type'] == 'InferenceTimeHook':
        h['infer_cfg']['input_size'] = (8, 3, 96, 96)
    if h['type'] == 'ModelAnalysis':
        h['input_shape'] = (3, 96, 96)
if mode == "l2":
    cfg.hooks = [dict(type='L2Reconstruct', priority=50, asym=(sys.argv[5] == "asym"), l2_weight=1.0, cls_weight=0.3,
                      epoch_behavior=[0, -2, -1],
                      dataset_args=dict(name='synthetic', batch_size=4, num_batches=2, num_classes=10),
                      data_config=dict(input_size=(3, 96, 96)),
                      optim_args=dict(opt='adamw', lr=1e-3, weight_decay=0.01),
                      sche_args=dict(epochs=3), other_args=dict(log_interval=1))]
update_cfg(work_dir=work, device='cuda:0', log_file=None, checkpoint=None, deploy=False, config_name="run", distributed=False,
           world_size=1, rank=0, local_rank=0, seed=42, skip_optim=None, skip_post=(True if mode == "l2" else None))
random_seed(42, 0)
n0 = B.launch_count()
runner = Runner(deploy=False, skip_optim=None, skip_post=cfg.skip_post)
if mode == "l2":
    # snapshot what the hook starts from: the fitted student (before any step) for the plain-torch restatement
    import copy
    orig_after_opt = runner.hooks[0].after_optimize
    def wrapped():
        torch.save(dict(student=copy.deepcopy(runner.model).cpu().state_dict()), os.path.join(work, "before.pth"))
        orig_after_opt()
    runner.hooks[0].after_optimize = wrapped
runner.run()
assert B.launch_count() > n0, "no kernel of libcna_b200.so ran"
model = runner.model.cuda().eval()
torch.manual_seed(7)
x = torch.randn(2, 3, 96, 96)
with torch.no_grad():
    y = model(x.cuda()).float().cpu()
assert torch.isfinite(y).all()
out = dict(x=x, y=y, keys=sorted(model.state_dict()))
if mode == "l2":
    out["history"] = runner.hooks[0].history
    out["after"] = {k: v.cpu() for k, v in model.state_dict().items()}
else:
    res = [getattr(h, "result", None) for h in runner.hooks]
    out["hook_results"] = res
torch.save(out, os.path.join(work, "out.pth"))
assert os.path.isfile(os.path.join(work, "run.pth"))                      # Runner.run saved the post-processed model
print("OVERLAY_RUN_OK", len(out["keys"]))
'''

REFERENCE_DEPLOY = STUBS.replace("STUB_NAMES", '("ptflops", "cvxpy", "timm", "timm.data", "timm.models", "timm.utils", "timm.loss", '
                                 '"timm.optim", "timm.scheduler")') + r'''
import os, torch
ref, cfg_rel, work = sys.argv[1:4]
import approx.layers, approx.core
assert approx.layers.__file__.startswith(os.path.join(ref, "src")) and approx.core.__file__.startswith(os.path.join(ref, "src"))   # the UNMODIFIED reference
from approx.runner import Runner
from approx.utils.config import init_cfg, update_cfg, get_cfg
init_cfg(os.path.join(ref, cfg_rel))
cfg = get_cfg()
cfg.model['init_cfg'] = None
cfg.hooks = []
update_cfg(work_dir=work, device='cpu', log_file=None, checkpoint=os.path.join(work, "run.pth"), deploy=True, config_name="ref",
           distributed=False, world_size=1, rank=0, local_rank=0, seed=42, skip_optim=True, skip_post=True)
torch.manual_seed(0)
runner = Runner(deploy=True, skip_optim=True, skip_post=True)
runner.run()                                                              # initialize (deploy) -> load the checkpoint -> save
model = runner.model.eval()
got = torch.load(os.path.join(work, "out.pth"))
assert sorted(model.state_dict()) == got["keys"], "state_dict keys differ between the backends"
with torch.no_grad():
    y = model(got["x"]).float()
err = ((y - got["y"]).norm() / got["y"].norm()).item()
print("REFERENCE_DEPLOY_ERR", err)
assert err < 2e-4, err
'''


def _run(script, args, overlay, timeout=900):
    if overlay:
        env = dict(os.environ, CNA_REFERENCE_SRC=os.path.join(REF, "src"), PYTHONPATH=PKG)
    else:
        env = dict(os.environ, PYTHONPATH=os.path.join(REF, "src"))
        env.pop("CNA_REFERENCE_SRC", None)
    out = subprocess.run([sys.executable, "-c", script, REF, *args], env=env, capture_output=True, text=True, cwd="/tmp",
                         timeout=timeout)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-6000:]
    return out.stdout


@needs_ref
@pytest.mark.parametrize("cfg", ["configs/msca-rep/msca-rep_d1_fix_mscan-t.py",
                                 "configs/msca-rep/profiler/msca-rep-profile_d0_fix_mscan-t.py",     # (msca-rep_d0_fix_mscan-t.py names a _base_ path that does not exist in the reference)
                                 "configs/msca-rep/fps/msca-rep_d1_mscan-t_fps.py",
                                 "configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py"])
def test_reference_runner_drives_this_backend_and_checkpoint_loads_in_the_reference(cfg, tmp_path):
    work = str(tmp_path)
    out = _run(OVERLAY_RUN, [cfg, work, "plain"], overlay=True)
    assert "OVERLAY_RUN_OK" in out
    out = _run(REFERENCE_DEPLOY, [cfg, work], overlay=False)
    assert "REFERENCE_DEPLOY_ERR" in out


@needs_ref
def test_reference_runner_dodecomp_config_runs_with_measurement_hooks(tmp_path):
    """`do_decomp=True` (rank-1 split of the bases in postprocess) + the shipped config's hooks on this backend."""
    import torch
    work = str(tmp_path)
    _run(OVERLAY_RUN, ["configs/low-rank-exp/low-rank-exp-v1_l2345_svd_dodecomp_alexnet.py", work, "plain"], overlay=True)
    got = torch.load(os.path.join(work, "out.pth"))
    assert any(k.endswith("s_conv.v_conv.weight") for k in got["keys"]) and any(k.endswith("s_conv.h_conv.weight") for k in got["keys"])


@needs_ref
@pytest.mark.parametrize("mode", ["sym"])
def test_l2_reconstruct_hook_follows_the_reference_arithmetic(mode, tmp_path):
    """3 epochs x 2 synthetic batches of `L2Reconstruct` (symmetric teacher; epoch_behavior = [0, -2, -1]: replaced layer 0 only,
    everything unfrozen, all replaced layers) through the reference's Runner on CUDA.  The same schedule is restated in plain torch on the CPU from the hook's
    starting state: teacher pass with the old layers (eval, no_grad), student pass with the new ones, per-sample norms,
    total = l2_w * mean_b(sum_l norm / L) + cls_w * CE, AdamW without decay on biases — per-step losses and the final parameters
    must agree (dropout is inactive: AlexNet's classifier dropout p is set to 0 through the config)."""
    import torch
    import torch.nn.functional as F
    sys.path.insert(0, ROOT)
    from oracle import approx_oracle as O
    work = str(tmp_path)
    cfg = "configs/low-rank-exp/low-rank-exp-v1_l2345_svd_alexnet.py"
    script = OVERLAY_RUN.replace("cfg.model['init_cfg'] = None", "cfg.model['init_cfg'] = None; cfg.model['dropout'] = 0.0")
    _run(script, [cfg, work, "l2", mode], overlay=True)
    got = torch.load(os.path.join(work, "out.pth"))
    P = {k: v.clone().float().requires_grad_(True) for k, v in torch.load(os.path.join(work, "before.pth"))["student"].items()}
    P0 = {k: v.detach().clone() for k, v in P.items()}
    SUBS = [(3, 64, 5), (6, 192, 3), (8, 384, 3), (10, 256, 3)]

    def forward(x, use_new, cache, norms):
        def sub(i, x, C, d):
            pre = f"features.{i}"
            if use_new:
                t = F.conv2d(x, P[pre + ".new.s_conv.weight"], P[pre + ".new.s_conv.bias"], 1, d // 2, 1, C)
                y = F.conv2d(t, P[pre + ".new.d_conv.weight"], P[pre + ".new.d_conv.bias"])
                norms.append(O.l2_recon_norm(y, cache[i]))
            else:
                y = F.conv2d(x, P[pre + ".old.weight"], P[pre + ".old.bias"], 1, d // 2)
                # the reference caches the teacher layer's output tensor ITSELF (hooks/finetune.py:76-77, no clone) and AlexNet's next
                # module is nn.ReLU(inplace=True): by the time the student's norm is taken the cached target holds relu(y_old)
                cache[i] = torch.relu(y).detach()
            return y
        x = F.max_pool2d(F.relu(F.conv2d(x, P["features.0.weight"], P["features.0.bias"], 4, 2)), 3, 2)
        x = F.max_pool2d(F.relu(sub(3, x, 64, 5)), 3, 2)
        for (i, C, d) in SUBS[1:]:
            x = F.relu(sub(i, x, C, d))
        x = torch.flatten(F.adaptive_avg_pool2d(F.max_pool2d(x, 3, 2), (6, 6)), 1)
        x = F.relu(F.linear(x, P["classifier.1.weight"], P["classifier.1.bias"]))
        x = F.relu(F.linear(x, P["classifier.4.weight"], P["classifier.4.bias"]))
        return F.linear(x, P["classifier.6.weight"], P["classifier.6.bias"])

    def trainable(behavior):
        if behavior == -2:                  # (reached with the "frozen" flag set by epoch 0, so the reference does unfreeze here)
            return [k for k in P]
        idx = [SUBS[behavior][0]] if behavior >= 0 else [s[0] for s in SUBS]
        return [k for k in P if any(k.startswith(f"features.{i}.") for i in idx)]     # freeze_except: the whole Substitution (old + new)

    state = {k: dict(m=torch.zeros_like(v), v=torch.zeros_like(v), t=0) for k, v in P.items()}
    lr, wd, b1, b2, eps = 1e-3, 0.01, 0.9, 0.999, 1e-8
    losses = []
    for epoch, behavior in enumerate([0, -2, -1]):
        names = set(trainable(behavior))
        gen = torch.Generator(device="cuda").manual_seed(42)       # SyntheticLoader: same seed every epoch, generated on the device
        for _ in range(2):
            x = torch.randn(4, 3, 96, 96, device="cuda", generator=gen).cpu()
            tgt = torch.randint(0, 10, (4,), device="cuda", generator=gen).cpu()
            cache, norms = {}, []
            with torch.no_grad():
                forward(x, False, cache, norms)
            out = forward(x, True, cache, norms)
            total_norm = (sum(norms) / len(SUBS)).mean()
            total = 1.0 * total_norm + 0.3 * F.cross_entropy(out, tgt)
            for v in P.values():
                v.grad = None
            total.backward()
            losses.append(total.item())
            with torch.no_grad():
                for k, v in P.items():
                    if k not in names or v.grad is None:
                        continue
                    st = state[k]
                    st["t"] += 1
                    decay = 0.0 if (v.ndim <= 1 or k.endswith(".bias")) else wd
                    v.mul_(1 - lr * decay)
                    st["m"].mul_(b1).add_(v.grad, alpha=1 - b1)
                    st["v"].mul_(b2).addcmul_(v.grad, v.grad, value=1 - b2)
                    denom = st["v"].sqrt() / (1 - b2 ** st["t"]) ** 0.5 + eps
                    v.addcdiv_(st["m"], denom, value=-lr / (1 - b1 ** st["t"]))
    hist = [h["total"] for h in got["history"]]
    assert len(hist) == len(losses) == 6
    for a, b in zip(hist, losses):
        assert abs(a - b) / abs(b) < 2e-3, (hist, losses)
    # Adam moves every trained element by ~lr per step whatever the gradient's size, so fp32 noise in near-zero gradients shows up
    # at the scale of the UPDATE, not of the parameter: the two trajectories must agree to a small fraction of the distance moved
    moved = 0
    for k, v in P.items():
        ref, mine = v.detach(), got["after"][k].float()
        step = (ref - P0[k]).norm()
        moved += int(step > 0)
        assert (ref - mine).norm() <= 0.1 * step + 1e-5 * (1 + ref.norm()), (k, (ref - mine).norm().item(), step.item())
    assert moved >= 20
