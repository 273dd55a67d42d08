"""Shim of diffusers.utils: BaseOutput, apply_forward_hook, deprecate, is_accelerate_available, randn_tensor."""
import torch

from . import import_utils  # noqa: F401


class BaseOutput:
    pass


def apply_forward_hook(fn):
    return fn


def deprecate(*a, **k):
    pass


def is_accelerate_available():
    return False


def check_min_version(v):
    pass


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """Appendix A.8 (recalled from diffusers 0.14 utils/torch_utils.py): sample on the generator's device in the requested dtype, then move; a LIST of
    generators (one per batch element) draws each element's (1, ...) slice from its own generator; a one-element list is unwrapped."""
    if isinstance(generator, (list, tuple)):
        if len(generator) == 1:
            generator = generator[0]
        else:
            parts = [torch.randn((1,) + tuple(shape[1:]), generator=g, device=g.device, dtype=dtype) for g in generator]
            return torch.cat(parts, dim=0).to(device or parts[0].device)
    gdev = generator.device if generator is not None else (device or torch.device("cpu"))
    return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device or gdev)
