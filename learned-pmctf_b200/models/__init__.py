from .pWave import pWave  # noqa: F401
