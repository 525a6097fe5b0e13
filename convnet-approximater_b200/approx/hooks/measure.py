"""Measurement hooks on the B200 backend: InferenceTimeHook, Fps, ModelAnalysis.

Constructor signatures and config keys are the reference's (src/approx/hooks/inference_time_hook.py:9-24, fps.py:13-27,
model_analysis.py:9-13), so the shipped configs build these unchanged.  What they report:

  InferenceTimeHook  torch.profiler table like the reference PLUS device-timed numbers the profiler cannot give reliably:
                     CUDA-event ms per forward (mean of `repeat` runs after a warm-up), images/s, and how many kernels of
                     libcna_b200.so one forward launches (so a silent fallback to eager PyTorch would show up as 0).
  Fps                the reference's loop (5 warm-up batches, then timed batches of `model(input)` bracketed by synchronize),
                     with the batches from `make_loader` (timm if a real dataset is named and installed, else synthetic) and the
                     timed region measured with CUDA events as well as the host clock.
  ModelAnalysis      MACs / parameters by forward hooks on Conv2d / Linear (ptflops is not a dependency of this backend).
"""
import time

import numpy as np
import torch
from torch import nn

from ._common import ArgTable, DEFAULT_DATA_CONFIG, DEFAULT_DATASET_ARGS, log, make_loader
from .hook import HOOK, Hook


def _device_time(fn, repeat, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(repeat):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / repeat


@HOOK.register_module()
class InferenceTimeHook(Hook):
    def __init__(self, runner, priority, infer_cfg):
        super().__init__(runner, priority)
        infer_cfg = dict(infer_cfg)
        self.input_size = tuple(infer_cfg.pop("input_size", (256, 3, 224, 224)))
        self.profile_args = infer_cfg.pop("profile_args", None)
        self.key_args = infer_cfg.pop("key_args", {})
        self.table_args = infer_cfg.pop("table_args", dict(sort_by='cuda_time_total', row_limit=10))
        self.repeat = int(infer_cfg.pop("repeat", 10))
        self.dtype = infer_cfg.pop("dtype", "float32")
        self.result = None
        log().info(f"InferenceTimeHook Config:\ninput_size: {self.input_size}\nprofile_args: {self.profile_args}\n"
                   f"key_args: {self.key_args}\ntable_args: {self.table_args}")

    def after_run(self):
        from approx import backend as B
        from torch.profiler import ProfilerActivity, profile
        dt = {"float32": torch.float32, "bfloat16": torch.bfloat16}[self.dtype]
        model = self.runner.model.cuda().to(dt).eval()
        x = torch.randn(*self.input_size, device="cuda").to(dt)
        with torch.no_grad():
            model(x)                                              # first call: weight packing, kernel attributes
            n0 = B.launch_count()
            model(x)
            launches = B.launch_count() - n0
            ms = _device_time(lambda: model(x), self.repeat, 2)
            args = self.profile_args or dict(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA])
            with profile(**args) as prof:
                model(x)
        table = prof.key_averages(**self.key_args).table(**self.table_args)
        self.result = dict(ms_per_forward=ms, images_per_s=self.input_size[0] / ms * 1e3, cna_kernel_launches=launches)
        log().info(f"Profile Result:\n{table}")
        log().info(f"Device time (CUDA events, mean of {self.repeat}): {ms:.3f} ms / forward of {self.input_size}, "
                   f"{self.result['images_per_s']:.1f} img/s; kernels launched from libcna_b200.so per forward: {launches}")


@HOOK.register_module()
class Fps(Hook):
    def __init__(self, runner, priority, repeat_times=1, log_interval=50, num_workers=8, use_torch_benchmark=False,
                 dataset_args={}, data_config={}, total_iters=200):
        super().__init__(runner, priority)
        self.repeat_times, self.log_interval, self.num_workers = repeat_times, log_interval, num_workers
        self.use_torch_benchmark = use_torch_benchmark
        self.dataset_args = ArgTable(DEFAULT_DATASET_ARGS, dataset_args)
        self.data_config = ArgTable(DEFAULT_DATA_CONFIG, data_config)
        self.total_iters = total_iters
        self.result = None

    def after_run(self):
        torch.backends.cudnn.benchmark = self.use_torch_benchmark      # kept for the torch layers left in the model
        model = self.runner.model.cuda().eval()
        num_warmup = 5
        fps_host, fps_dev = [], []
        for rep in range(self.repeat_times):
            print(f'Run {rep + 1}:')
            loader = make_loader('validation', self.dataset_args, self.data_config, dict(num_workers=self.num_workers))
            host_s, dev_ms, imgs = 0.0, 0.0, 0
            for i, (inp, _) in enumerate(loader):
                inp = inp.cuda()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                a.record()
                with torch.no_grad():
                    model(inp)
                b.record()
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                if i >= num_warmup:
                    host_s += dt
                    dev_ms += a.elapsed_time(b)
                    imgs += inp.size(0)
                    if (i + 1) % self.log_interval == 0:
                        print(f'Done iter [{i + 1:<3}/ {self.total_iters}], fps: {imgs / host_s:.2f} img / s')
                if (i + 1) == self.total_iters:
                    break
            if imgs == 0:
                raise RuntimeError("Fps: the loader produced no batch after the warm-up iterations")
            print(f'Overall fps: {imgs / host_s:.2f} img / s (host clock), {imgs / dev_ms * 1e3:.2f} img / s (CUDA events)\n')
            fps_host.append(imgs / host_s)
            fps_dev.append(imgs / dev_ms * 1e3)
        self.result = dict(average_fps=round(float(np.mean(fps_host)), 2), fps_variance=round(float(np.var(fps_host)), 4),
                           average_fps_device=round(float(np.mean(fps_dev)), 2))
        print(f'Average fps of {self.repeat_times} evaluations: {self.result["average_fps"]}')
        print(f'The variance of {self.repeat_times} evaluations: {self.result["fps_variance"]}')


def count_macs_params(model: nn.Module, input_shape, device=None):
    """(MACs of one forward of a (1, *input_shape) input, parameter count): Conv2d and Linear layers, counted by forward hooks.
    Runs on `device` (default: where the model is); the replaced layers keep stock nn.Conv2d containers, so their factor
    convolutions are counted like any other conv."""
    macs = [0]
    handles = []

    def conv_hook(m, inp, out):
        kh, kw = m.kernel_size
        macs[0] += out.numel() // out.shape[0] * (m.in_channels // m.groups) * kh * kw

    def lin_hook(m, inp, out):
        macs[0] += out.numel() // out.shape[0] * m.in_features

    # the replaced layers never call their nn.Conv2d containers' forward, so count those from the containers' shapes
    def container_hook(m, inp, out):
        x = inp[0]
        H, W = x.shape[2], x.shape[3]
        for sub in m.modules():
            if isinstance(sub, nn.Conv2d) and sub is not m:
                kh, kw = sub.kernel_size
                ho = (H + 2 * sub.padding[0] - kh) // sub.stride[0] + 1
                wo = (W + 2 * sub.padding[1] - kw) // sub.stride[1] + 1
                macs[0] += ho * wo * sub.out_channels * (sub.in_channels // sub.groups) * kh * kw
                H, W = ho, wo          # sub-convs run in definition order (MSCA's are all 'same'-sized)

    from approx.layers import LowRankExpConvV1, LowRankExpConvV2, MSCA
    own = (LowRankExpConvV1, LowRankExpConvV2, MSCA)
    own_ids = set()
    for m in model.modules():
        if isinstance(m, own):
            handles.append(m.register_forward_hook(container_hook))
            own_ids.update(id(s) for s in m.modules())
    for m in model.modules():
        if id(m) in own_ids:
            continue
        if isinstance(m, nn.Conv2d):
            handles.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.Linear):
            handles.append(m.register_forward_hook(lin_hook))
    p = next(model.parameters())
    x = torch.zeros(1, *input_shape, device=device or p.device, dtype=p.dtype)
    was = model.training
    model.eval()
    with torch.no_grad():
        model(x)
    model.train(was)
    for h in handles:
        h.remove()
    return macs[0], sum(q.numel() for q in model.parameters())


@HOOK.register_module()
class ModelAnalysis(Hook):
    def __init__(self, runner, priority, input_shape=(3, 224, 224)):
        super().__init__(runner, priority)
        self.input_shape = tuple(input_shape)
        self.result = None

    def after_run(self):
        model = self.runner.model.cuda()
        macs, params = count_macs_params(model, self.input_shape)
        self.result = dict(macs=macs, params=params)
        log().info(f"Model Macs: {macs / 1e9:.3f} GMac, Params: {params / 1e6:.3f} M")


@HOOK.register_module()
class ClassEvalHook(Hook):
    """after_run: top-1 / top-5 / loss of the finished model on the validation split (reference hooks/class_eval_hook.py:58-68,
    classification/validate.py).  eval_cfg keys of the reference that matter here: dataset, data, batch_size, input_size, mean, std,
    interpolation, crop_pct.  Without a dataset library (timm is not a dependency of this backend) the hook evaluates on the synthetic
    loader and SAYS SO: the accuracies are then meaningless, the images/s figure is not."""

    def __init__(self, runner, priority, eval_cfg):
        super().__init__(runner, priority)
        cfg = dict(batch_size=128, input_size=(3, 224, 224), dataset='', data='path/to/dataset', mean=(0.485, 0.456, 0.406),
                   std=(0.229, 0.224, 0.225), interpolation='bicubic', crop_pct=0.875, num_batches=8, num_classes=1000)
        cfg.update(eval_cfg)
        self.eval_cfg = cfg
        self.result = None

    def after_run(self):
        cfg = self.eval_cfg
        ds = ArgTable(dict(name=cfg['dataset'], root=cfg['data'], batch_size=cfg['batch_size'], num_batches=cfg['num_batches'],
                           num_classes=cfg['num_classes']))
        dc = ArgTable(DEFAULT_DATA_CONFIG, dict(input_size=cfg['input_size'], mean=cfg['mean'], std=cfg['std'],
                                                interpolation=cfg['interpolation'], crop_pct=cfg['crop_pct']))
        synthetic = False
        try:
            loader = make_loader('validation', ds, dc, dict(num_workers=cfg.get('workers', 4)))
        except RuntimeError as e:
            log().warning(f"ClassEvalHook: {e}; evaluating on SYNTHETIC data — accuracy numbers below are meaningless")
            ds['name'] = 'synthetic'
            loader = make_loader('validation', ds, dc, None)
            synthetic = True
        model = self.runner.model.cuda().eval()
        loss_fn = nn.CrossEntropyLoss()
        n, top1, top5, loss_sum = 0, 0, 0, 0.0
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        with torch.no_grad():
            for inp, target in loader:
                inp, target = inp.cuda(), target.cuda()
                out = model(inp).float()
                target = target.clamp_max(out.shape[1] - 1)
                loss_sum += float(loss_fn(out, target)) * inp.size(0)
                hit = out.topk(min(5, out.shape[1]), dim=1).indices.eq(target.view(-1, 1))
                top1 += int(hit[:, :1].sum()); top5 += int(hit.sum()); n += inp.size(0)
        b.record()
        torch.cuda.synchronize()
        n = max(n, 1)
        self.result = dict(top1=100.0 * top1 / n, top5=100.0 * top5 / n, loss=loss_sum / n, images=n,
                           images_per_s=n / a.elapsed_time(b) * 1e3, synthetic=synthetic or ds['name'] in ('', 'synthetic'))
        log().info(f"ClassEval: {self.result}")
