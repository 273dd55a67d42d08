from ladi_oracle.parts import prepare_mask_and_masked_image  # noqa: F401
