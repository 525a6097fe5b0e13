"""`L2Reconstruct`: output-L2 reconstruction fine-tuning of the replaced layers, on this backend's kernels.

Semantics of the reference hook (src/approx/hooks/finetune.py:95-420), restated:

  teacher   asym=True : a second copy of the model (`ori_model`) keeps the ORIGINAL layers and sees the same batch; each replaced
                        layer is fitted to the teacher layer's output while its own input comes from the student's earlier,
                        already-replaced layers (:126-149, 294-299).
            asym=False: the same model runs the batch twice — `Substitution`s switched to `old` (eval, no_grad, outputs cached),
                        then switched to `new` for the training pass (:300-312).
  loss      norm_l[b] = || y_new_l[b] - y_old_l[b] ||_2  (not squared, :80-82);
            total = l2_weight * mean_b( sum_l norm_l[b] / L ) + cls_weight * CrossEntropy(logits, target)   (:313-322)
  schedule  epoch_behavior[e] >= 0 : train only replaced layer e's switchable parameters (`freeze_except(e)`);
            -1 : train all replaced layers (the reference passes every index to freeze_except);  -2 : unfreeze everything (:227-248)
  update    AdamW with no weight decay on biases / 1-D parameters (timm's create_optimizer_v2 default, :188); in distributed
            runs the gradients of the trainable parameters are summed over ranks with ONE flat all-reduce per step — the reference
            only wraps the model in DDP when sync_bn is set (:153-160), so without it its ranks silently diverge; this hook always
            averages.

What runs where: the norm, its backward and the AdamW update are this library's kernels (cna_l2_recon_loss_grad, cna_l2_norm_backward,
cna_adamw_step); the replaced layers' forward/backward are the library's autograd functions (train mode); everything the
reference model keeps as stock torch modules (BatchNorm, Linear, activations) stays torch.  Data comes from
`approx.hooks._common.make_loader` (timm if installed and a dataset is named, a synthetic loader otherwise).
"""
import os
import time
from collections import OrderedDict

import torch
import torch.distributed as dist
from torch import nn

from approx import backend as B
from ._common import (ArgTable, DEFAULT_DATA_CONFIG, DEFAULT_DATASET_ARGS, DEFAULT_OPTIM_ARGS, DEFAULT_OTHER_ARGS,
                      DEFAULT_SCHED_ARGS, log, make_loader)
from .hook import HOOK, Hook


# ----------------------------------------------------------------------------------------------------------------------
# loss pieces
# ----------------------------------------------------------------------------------------------------------------------
def _same_layout(y_new, y_old):
    """Both tensors as dense buffers in the SAME memory order (channels-last if y_new is, else NCHW): the norm does not care about
    the order, only that element i of one is element i of the other."""
    if y_new.dim() == 4 and not y_new.is_contiguous() and y_new.permute(0, 2, 3, 1).is_contiguous():
        a = y_new.permute(0, 2, 3, 1)
        b = y_old.permute(0, 2, 3, 1)
        return a, (b if b.is_contiguous() else b.contiguous()), True
    return y_new.contiguous(), y_old.contiguous(), False


class _L2Norm(torch.autograd.Function):
    """norms[b] = ||y_new[b] - y_old[b]||_2 with this library's kernels in both directions."""

    @staticmethod
    def forward(ctx, y_new, y_old):
        a, b, cl = _same_layout(y_new.detach(), y_old.detach().to(y_new.dtype))
        # the unit gradient (y_new - y_old) / norm is produced with the norms and saved INSTEAD of y_new: the layers that follow a
        # replaced conv in the reference's models are in-place (nn.ReLU(inplace=True)) and overwrite y_new before backward runs
        norms, unit = B.l2_recon(a, b, 1.0, want_grad=True)
        ctx.save_for_backward(unit)
        ctx.cl = cl
        return norms

    @staticmethod
    def backward(ctx, gn):
        (unit,) = ctx.saved_tensors
        g = B.l2_norm_backward(unit, None, None, gn)
        return (g.permute(0, 3, 1, 2) if ctx.cl else g), None


def l2_norm(y_new, y_old):
    return _L2Norm.apply(y_new, y_old)


def cache_module_output(module, inputs, output):
    # NOT cloned, exactly like the reference (hooks/finetune.py:76-77): when the next module works in place — nn.ReLU(inplace=True)
    # after every conv of the reference's AlexNet — the cached teacher target is what that module leaves behind (relu(y_old)), and
    # the student layers are fitted to THAT.  Kept for result parity with the reference's fine-tuning runs.
    module.cache['ori_output'] = output


def get_l2_error(module, inputs, output):
    module.cache['norm'] = l2_norm(output, module.cache['ori_output'])


# ----------------------------------------------------------------------------------------------------------------------
# optimizer
# ----------------------------------------------------------------------------------------------------------------------
class FusedAdamW:
    """AdamW over a list of parameters through cna_adamw_step (fp32 state per parameter).  Decay rule of timm's
    create_optimizer_v2 (filter_bias_and_bn=True): parameters with ndim <= 1 or named '*.bias' get weight_decay = 0."""

    def __init__(self, named_params, lr, weight_decay, betas=(0.9, 0.999), eps=1e-8):
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.items = []
        for name, p in named_params:
            no_decay = p.ndim <= 1 or name.endswith(".bias")
            self.items.append(dict(name=name, p=p, decay=not no_decay, m=None, v=None, master=None))
        self.t = 0
        self.param_groups = [dict(lr=lr)]

    def zero_grad(self):
        for it in self.items:
            it["p"].grad = None

    def trainable(self):
        return [it for it in self.items if it["p"].requires_grad and it["p"].grad is not None]

    def allreduce_grads(self, group=None):
        """One flat all-reduce(SUM) of every trainable gradient, then division by the world size (mean over ranks)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        its = self.trainable()
        if not its:
            return
        flat = torch.cat([it["p"].grad.detach().float().reshape(-1) for it in its])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat /= dist.get_world_size(group)
        o = 0
        for it in its:
            n = it["p"].numel()
            it["p"].grad.copy_(flat[o:o + n].view_as(it["p"]))
            o += n

    def step(self):
        self.t += 1
        lr = self.param_groups[0]["lr"]
        for it in self.trainable():
            p = it["p"]
            if it["m"] is None:
                it["m"] = torch.zeros(p.numel(), device=p.device, dtype=torch.float32)
                it["v"] = torch.zeros(p.numel(), device=p.device, dtype=torch.float32)
                it["steps"] = 0
            it["steps"] += 1
            g = p.grad.detach().to(torch.float32).contiguous().view(-1)
            if p.dtype == torch.float32 and p.is_contiguous():
                B.adamw_step_(p.data.view(-1), g, it["m"], it["v"], it["steps"], lr, self.betas, self.eps, self.wd,
                              n_decay=p.numel() if it["decay"] else 0)
                torch.autograd.graph.increment_version(p)      # the kernel wrote through the raw pointer: packed-weight caches key on _version
            else:
                if it["master"] is None:
                    it["master"] = p.detach().to(torch.float32).contiguous().view(-1).clone()
                B.adamw_step_(it["master"], g, it["m"], it["v"], it["steps"], lr, self.betas, self.eps, self.wd,
                              n_decay=p.numel() if it["decay"] else 0)
                p.data = it["master"].view_as(p).to(p.dtype)


def _schedule(sche_args, base_lr):
    """lr(epoch): None / 'constant' (the reference's default: create_scheduler with sched=None), 'cosine', 'step'."""
    kind, epochs = sche_args['sched'], sche_args['epochs']
    if kind in (None, 'constant', ''):
        return lambda e: base_lr
    if kind == 'cosine':
        import math
        lo = sche_args['min_lr']
        return lambda e: lo + 0.5 * (base_lr - lo) * (1 + math.cos(math.pi * min(e, epochs) / max(1, epochs)))
    if kind == 'step':
        every = max(1, int(sche_args.get('decay_epochs', max(1, epochs // 3))))
        return lambda e: base_lr * sche_args['decay_rate'] ** (e // every)
    raise ValueError(f"unknown scheduler {kind!r} (None, 'cosine' or 'step')")


# ----------------------------------------------------------------------------------------------------------------------
# the hook
# ----------------------------------------------------------------------------------------------------------------------
@HOOK.register_module()
class L2Reconstruct(Hook):
    def __init__(self, runner, priority, asym=True, l2_weight=1.0, cls_weight=0.0, epoch_behavior=(), no_norm=False,
                 dataset_args={}, optim_args={}, sche_args={}, data_config={}, other_args={}):
        super().__init__(runner, priority)
        self.asym, self.l2_weight, self.cls_weight = asym, l2_weight, cls_weight
        self.epoch_behavior, self.no_norm = list(epoch_behavior), no_norm
        self.model = runner.model
        self.dataset_args = ArgTable(DEFAULT_DATASET_ARGS, dataset_args)
        self.optim_args = ArgTable(DEFAULT_OPTIM_ARGS, optim_args)
        self.sche_args = ArgTable(DEFAULT_SCHED_ARGS, sche_args)
        self.data_config = ArgTable(DEFAULT_DATA_CONFIG, data_config)
        self.other_args = ArgTable(DEFAULT_OTHER_ARGS, other_args)
        self.ori_model = None
        self.history = []                      # per step: dict(loss, norm, total) — what the reference only logs
        if self.asym and not self.no_norm:
            from approx.models import build_model
            self.ori_model = build_model(runner.cfg.model)

    # -- teacher / student set-up (reference :120-151) -----------------------------------------------------------------
    def _prepare(self):
        for sub in self.model.switchable_modules():
            sub.switch_new(remove_old=self.no_norm)
        if self.ori_model is not None:
            runner = self.runner
            for f in runner.filters:
                f.rewind()
            runner.app.rewind()
            self.ori_model.register_switchable(runner.app.src_type, runner.filters)
            self.ori_model.init_weights()
            for idx in range(self.ori_model.length_switchable):
                src = self.ori_model.get_switchable_module(idx)
                self.ori_model.set_switchable_module(idx, runner.app.initialize, src=src)
            for sub in self.ori_model.switchable_modules():
                sub.switch_old(remove_new=True)
                sub.register_forward_hook(cache_module_output)
            for sub in self.model.switchable_modules():
                sub.switch_new(remove_old=True)
            self.ori_model.eval()
            self.ori_model.cuda()
        self.model.cuda()

    def _world(self):
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def after_optimize(self):
        self._prepare()
        num_layers = self.model.length_switchable
        loader_train = make_loader('train', self.dataset_args, self.data_config, self.other_args)
        loader_eval = make_loader('validation', self.dataset_args, self.data_config, self.other_args)
        opt = FusedAdamW(list(self.model.named_parameters()), lr=self.optim_args['lr'],
                         weight_decay=self.optim_args['weight_decay'], eps=self.optim_args['eps'])
        lr_at = _schedule(self.sche_args, self.optim_args['lr'])
        num_epochs = self.sche_args['epochs']
        start_epoch = self.other_args['start_epoch'] or 0
        behavior = list(self.epoch_behavior)
        behavior = (behavior + [-1] * (num_epochs - len(behavior)))[:num_epochs]
        log().info(f'Scheduled epochs: {num_epochs}')
        log().info(f'epoch_behaviors: {behavior}')
        loss_fn = nn.CrossEntropyLoss().cuda()
        eval_metric = self.other_args['eval_metric']
        best, best_epoch = None, None
        out_dir = getattr(self.runner.cfg, "work_dir", None) if (not dist.is_initialized() or dist.get_rank() == 0) else None
        frozen = False
        for epoch in range(start_epoch, num_epochs):
            if behavior[epoch] >= 0:
                self.model.freeze_except(behavior[epoch])
                frozen = True
            elif behavior[epoch] == -1:
                self.model.freeze_except(*range(num_layers))
                frozen = False
            elif frozen:
                self.model.unfreeze()
                frozen = False
            opt.param_groups[0]["lr"] = lr_at(epoch)
            if hasattr(getattr(loader_train, "sampler", None), 'set_epoch'):
                loader_train.sampler.set_epoch(epoch)
            train_metrics = self.train_one_epoch(epoch, loader_train, loss_fn, opt)
            eval_metrics = self.validate(loader_eval, loss_fn)
            metric = eval_metrics[eval_metric]
            better = best is None or (metric < best if eval_metric == 'loss' else metric > best)
            if better:
                best, best_epoch = metric, epoch
            if out_dir:
                self._save(os.path.join(out_dir, "last.pth.tar"), epoch, metric)
                if better:
                    self._save(os.path.join(out_dir, "model_best.pth.tar"), epoch, metric)
            log().info(f"epoch {epoch}: train {dict(train_metrics)}  eval {dict(eval_metrics)}")
        if best is not None:
            log().info('*** Best metric: {0} (epoch {1})'.format(best, best_epoch))

    def _save(self, path, epoch, metric):
        torch.save({'epoch': epoch, 'state_dict': self.model.state_dict(), 'metric': metric}, path)

    # -- one epoch (reference :260-377) --------------------------------------------------------------------------------
    def _teacher_pass(self, inp):
        subs = list(self.model.switchable_modules())
        if self.asym:
            with torch.no_grad():
                self.ori_model(inp)
            for ori_sub, sub in zip(self.ori_model.switchable_modules(), subs):
                sub.cache['ori_output'] = ori_sub.cache['ori_output']
            return
        for sub in subs:                                     # symmetric: the same model with the old layers switched in
            sub.switch_old(remove_new=False)
            sub._forward_hooks = OrderedDict()
            sub.register_forward_hook(cache_module_output)
        self.model.eval()
        with torch.no_grad():
            self.model(inp)
        for sub in subs:
            sub.switch_new(remove_old=False)
            sub._forward_hooks = OrderedDict()
            sub.register_forward_hook(get_l2_error)
        self.model.train()

    def train_one_epoch(self, epoch, loader, loss_fn, optimizer):
        self.model.train()
        subs = list(self.model.switchable_modules())
        if self.asym and not self.no_norm:
            for sub in subs:
                sub._forward_hooks = OrderedDict()
                sub.register_forward_hook(get_l2_error)
        tot_sum, n_seen, t0 = 0.0, 0, time.time()
        last = len(loader) - 1
        for batch_idx, (inp, target) in enumerate(loader):
            inp, target = inp.cuda(), target.cuda()
            if not self.no_norm:
                self._teacher_pass(inp)
            output = self.model(inp)
            loss = loss_fn(output.float(), target)
            total_norm = 0
            if not self.no_norm:
                for sub in subs:
                    total_norm = total_norm + sub.cache['norm']
                total_norm = (total_norm / self.model.length_switchable).mean()
            total_loss = self.l2_weight * total_norm + self.cls_weight * loss
            optimizer.zero_grad()
            total_loss.backward()
            optimizer.allreduce_grads()
            optimizer.step()
            rec = dict(loss=float(loss), norm=float(total_norm) if not self.no_norm else 0.0, total=float(total_loss))
            self.history.append(rec)
            tot_sum += rec["total"] * inp.size(0)
            n_seen += inp.size(0)
            if batch_idx == last or batch_idx % self.other_args['log_interval'] == 0:
                dt = max(time.time() - t0, 1e-9)
                log().info(f"Train: {epoch} [{batch_idx:>4d}/{len(loader)}]  Loss: {rec['loss']:#.4g}  Norm: {rec['norm']:#.4g}  "
                           f"Total: {rec['total']:#.4g}  {n_seen * self._world() / dt:>7.2f} img/s  LR: {optimizer.param_groups[0]['lr']:.3e}")
        if not self.no_norm:
            for sub in subs:
                sub._forward_hooks = OrderedDict()
                sub.cache = {}
        return OrderedDict([('loss', tot_sum / max(1, n_seen))])

    def validate(self, loader, loss_fn):
        self.model.eval()
        loss_sum, top1, top5, n = 0.0, 0, 0, 0
        with torch.no_grad():
            for inp, target in loader:
                inp, target = inp.cuda(), target.cuda()
                out = self.model(inp).float()
                loss_sum += float(loss_fn(out, target)) * inp.size(0)
                k = min(5, out.shape[1])
                pred = out.topk(k, dim=1).indices
                hit = pred.eq(target.view(-1, 1))
                top1 += int(hit[:, :1].sum())
                top5 += int(hit.sum())
                n += inp.size(0)
        stats = torch.tensor([loss_sum, top1, top5, n], dtype=torch.float64, device="cuda")
        if self._world() > 1:
            dist.all_reduce(stats)
        loss_sum, top1, top5, n = stats.tolist()
        n = max(n, 1.0)
        return OrderedDict([('loss', loss_sum / n), ('top1', 100.0 * top1 / n), ('top5', 100.0 * top5 / n)])
