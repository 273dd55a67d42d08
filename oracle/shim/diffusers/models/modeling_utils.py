"""Shim of diffusers.models.modeling_utils.ModelMixin."""
import torch


class ModelMixin(torch.nn.Module):
    def __getattr__(self, name):
        # diffusers falls back to config entries for unknown attributes
        # (AutoencoderKL.py:114 reads `self.block_out_channels`)
        try:
            return super().__getattr__(name)
        except AttributeError:
            d = self.__dict__.get("_internal_dict", {})
            if name in d:
                return d[name]
            raise

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device
