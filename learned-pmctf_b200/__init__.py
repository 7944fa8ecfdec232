"""learned_pmctf_b200 -- B200 (sm_100a) implementation of Learned-pMCTF's motion-compensated
temporal lifting + pWave++ spatial lifting hot path behind the reference's nn.Module surface.

Layout mirrors the reference package for the modules on the path:
    layers.video.video_net                       flow_warp, bilineardownsacling, ME_Spynet (SpyNet)
    layers.context_fusion_4step / context_fusion / long_context   entropy-parameter networks (section 8f row 1)
    layers.video.wavelet_transform_temporal_mctf TemporalLifting
    layers.lifting_1d                            PredictUpdate, iWave1D, split, merge
    layers.wavelet_transform                     LiftingScheme2D
    layers.postprocessing                        PostProcess (de-quantisation filter)
    entropy_models.entropy_models                EntropyCoder, GaussianEncoder (rANS boundary)
    entropy_models.gaussian_model                CompressionModel
    models.MLCodec_rans / models.MLCodec_CXX     drop-ins for the reference's two pybind11 extensions (native coder behind the C ABI)
    models.pWave                                 pWave (transform + quantiser)
    models.video.pMCTF_L                         pMCTF (forward_MCTF / inverse_MCTF), accelerate()
    gop                                          dyadic GOP schedule (test_pMCTF_flex.py:131-291)
    parallel                                     GOP sharding + NCCL stats gather
"""
from . import _native, ops  # noqa: F401
from .layers import LiftingScheme2D, PostProcess, PredictUpdate, iWave1D  # noqa: F401
from .layers.video.video_net import ME_Spynet, MEBasic, bilineardownsacling, flow_warp  # noqa: F401
from .layers.video.wavelet_transform_temporal_mctf import TemporalLifting  # noqa: F401
from .models.pWave import pWave  # noqa: F401
from .models.video.pMCTF_L import accelerate, pMCTF  # noqa: F401

__all__ = ["flow_warp", "bilineardownsacling", "PredictUpdate", "iWave1D", "LiftingScheme2D", "TemporalLifting", "PostProcess",
           "pWave", "pMCTF", "accelerate", "ops"]
