from ladi_oracle.parts import DDIMScheduler  # noqa: F401


class PNDMScheduler:  # names only (tryon_pipe.py:19 type annotations)
    pass


class LMSDiscreteScheduler:
    pass
