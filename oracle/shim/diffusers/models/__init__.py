from ladi_oracle.unet import UNet2DConditionModel  # noqa: F401
from ladi_oracle.vae import AutoencoderKL  # noqa: F401  (stock class; the reference uses its own fork)
