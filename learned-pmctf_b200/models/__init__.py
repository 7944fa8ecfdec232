from .pWave import pWave  # noqa: F401
from . import MLCodec_CXX, MLCodec_rans  # noqa: F401  (drop-ins for the reference's two pybind11 extensions)
