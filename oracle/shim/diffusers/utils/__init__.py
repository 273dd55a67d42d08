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
    """Appendix A.8: sample on the generator's device in the requested dtype, then move."""
    gdev = generator.device if generator is not None else (device or torch.device("cpu"))
    return torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device or gdev)
