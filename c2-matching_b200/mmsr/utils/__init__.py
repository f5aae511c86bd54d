from .util import tensor2img  # noqa: F401
