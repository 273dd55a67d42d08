"""Shim of diffusers.pipeline_utils.DiffusionPipeline (register_modules, device, to, progress_bar, numpy_to_pil)."""
import contextlib

import torch

from .configuration_utils import ConfigMixin


class _Bar:
    def update(self, n=1):
        pass


class DiffusionPipeline(ConfigMixin):
    def register_modules(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)
        self.register_to_config(**{k: (type(v).__module__, type(v).__name__) for k, v in kw.items()})

    @property
    def device(self):
        for v in self.__dict__.values():
            if isinstance(v, torch.nn.Module):
                return next(v.parameters()).device
        return torch.device("cpu")

    def to(self, device):
        for v in self.__dict__.values():
            if isinstance(v, torch.nn.Module):
                v.to(device)
        return self

    @contextlib.contextmanager
    def progress_bar(self, iterable=None, total=None):
        yield _Bar()

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None]
        images = (images * 255).round().astype("uint8")
        return [Image.fromarray(im) for im in images]
