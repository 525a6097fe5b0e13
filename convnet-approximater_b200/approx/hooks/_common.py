"""Shared pieces of the hooks: logger, synthetic data, default argument tables."""
import logging

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

# defaults of the reference's hooks (src/approx/hooks/finetune.py:25-73); the dataset library is not part of this backend, so a
# dataset name of '' or 'synthetic' selects the synthetic loader below
DEFAULT_DATASET_ARGS = dict(name='', root='path/to/dataset', batch_size=64)
DEFAULT_DATA_CONFIG = dict(input_size=(3, 224, 224), interpolation='bicubic', mean=IMAGENET_MEAN, std=IMAGENET_STD, crop_pct=0.875)
DEFAULT_OPTIM_ARGS = dict(opt='adamw', lr=1e-3, momentum=0.9, weight_decay=0.05, eps=1e-8)
DEFAULT_OTHER_ARGS = dict(log_interval=50, num_workers=8, sync_bn=False, dist_bn='reduce', resume='', no_resume_opt=False,
                          start_epoch=None, eval_metric='top1', checkpoint_hist=10)
DEFAULT_SCHED_ARGS = dict(epochs=20, lr_noise=None, lr_noise_pct=0.67, lr_noise_std=1.0, lr_cycle_mul1=1.0, lr_cycle_decay=0.5,
                          lr_cycle_limit=1, sched=None, min_lr=1e-6, warmup_lr=1e-6, warmup_epochs=0, decay_rate=0.1, lr_k_decay=1.0)


def log():
    try:
        from approx.utils.logger import get_logger     # the reference's logger (overlay mode)
        return get_logger()
    except Exception:  # noqa: BLE001
        return logging.getLogger("approx")


class ArgTable(dict):
    """dict with attribute access: defaults overridden by the config's entries (the reference's `combine_config`)."""

    def __init__(self, defaults, overrides=None):
        super().__init__(defaults)
        self.update(overrides or {})

    __getattr__ = dict.__getitem__


class SyntheticLoader:
    """Deterministic stand-in for a timm loader: `num_batches` batches of N(0,1) images and uniform labels per epoch,
    generated on the device (seeded per rank like the reference's `random_seed(seed, rank)`, utils/random.py:6-9)."""

    def __init__(self, batch_size, input_size=(3, 224, 224), num_batches=8, num_classes=10, seed=42, device="cuda", dtype=torch.float32):
        self.batch_size, self.input_size, self.num_batches, self.num_classes = batch_size, tuple(input_size), num_batches, num_classes
        self.seed, self.device, self.dtype = seed, torch.device(device), dtype
        self.sampler = None

    def __len__(self):
        return self.num_batches

    def __iter__(self):
        g = torch.Generator(device=self.device).manual_seed(self.seed)
        for _ in range(self.num_batches):
            x = torch.randn(self.batch_size, *self.input_size, device=self.device, generator=g).to(self.dtype)
            t = torch.randint(0, self.num_classes, (self.batch_size,), device=self.device, generator=g)
            yield x, t


_loader_factory = None


def set_loader_factory(fn):
    """fn(split, dataset_args, data_config, other_args) -> iterable of (input, target).  Lets a deployment plug its own data
    pipeline (e.g. timm's create_dataset/create_loader, as the reference does) into L2Reconstruct / Fps."""
    global _loader_factory
    _loader_factory = fn


def make_loader(split, dataset_args, data_config, other_args=None):
    if _loader_factory is not None:
        return _loader_factory(split, dataset_args, data_config, other_args)
    name = dataset_args.get('name', '')
    if name not in ('', 'synthetic'):
        try:
            from timm.data import create_dataset, create_loader
        except ImportError as e:
            raise RuntimeError(f"dataset {name!r} needs timm (not installed): use name='synthetic' or set_loader_factory()") from e
        ds = create_dataset(split=split, **dataset_args)
        return create_loader(ds, input_size=data_config['input_size'], batch_size=dataset_args['batch_size'],
                             is_training=split == 'train', no_aug=True, mean=data_config['mean'], std=data_config['std'],
                             interpolation=data_config['interpolation'], num_workers=(other_args or {}).get('num_workers', 8),
                             distributed=False)
    return SyntheticLoader(dataset_args['batch_size'], data_config['input_size'],
                           num_batches=dataset_args.get('num_batches', 8), num_classes=dataset_args.get('num_classes', 10),
                           seed=dataset_args.get('seed', 42) + (0 if split == 'train' else 1))
