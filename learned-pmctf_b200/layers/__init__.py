from .layers import ClampNoGradient, RoundNoGradient, conv3x3  # noqa: F401
from .lifting_1d import PredictUpdate, iWave1D, merge, split  # noqa: F401
from .wavelet_transform import LiftingScheme2D  # noqa: F401
from .postprocessing import PostProcess, ResBlock  # noqa: F401
from .context_fusion_4step import ContextFusionFourStep, ContextResidual, DepthConvBlock  # noqa: F401
