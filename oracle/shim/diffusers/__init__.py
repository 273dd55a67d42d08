"""ORACLE SHIM (test infrastructure only; used in the build container, never shipped on the product path).

Minimal stand-in for the `diffusers==0.14.0` symbols that the reference's own files import
(/root/reference/src/vto_pipelines/tryon_pipe.py:14-21, src/models/vae.py:21-23,
src/models/AutoencoderKL.py:12-14, hubconf.py:12), so those files can be imported UNMODIFIED
from /root/reference and executed on CPU to pin oracle/ladi_oracle.  The arithmetic lives in
oracle/ladi_oracle (restated from SURVEY.md Appendix A); this package only adapts names.
"""
__version__ = "0.14.0"
from .models import UNet2DConditionModel  # noqa: E402,F401
