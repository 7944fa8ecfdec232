"""Separable 2-D lifting (reference: pMCTF/layers/wavelet_transform.py:8-57).

Rows first, then columns with the SAME iWave1D (lift_v is lift_h) applied to transposed views.
The kernels read and write those views through strides, so no permute/contiguous copies, no
split/merge copies, and the two column transforms run as one batched launch sequence."""
import torch.nn as nn

from .. import ops
from .lifting_1d import iWave1D


class LiftingScheme2D(nn.Module):
    def __init__(self, non_separable=False, bitdepth=8, lossy=True, in_channels=1, haar=False):
        super().__init__()
        if haar or non_separable:
            raise NotImplementedError("Haar / non-separable variants have no callers in the reference (SURVEY.md section 2.1 #3)")
        self.bitdepth = bitdepth
        self.dynamic_range = float(2 ** bitdepth)
        self.non_separable = non_separable
        self.lift_h = iWave1D(bitdepth=bitdepth, lossy=lossy, in_channels=in_channels)
        self.lift_v = self.lift_h  # wavelet_transform.py:20-21 -> duplicated state_dict keys
        self.ll_subband = None

    def forward_lift_2d(self, x):
        """-> {'ll','lh','hl','hh','l','h'} (wavelet_transform.py:25-43)."""
        from .. import train
        if train.needs_grad(x, self):
            return train.lift2d_forward(self, x)
        return ops.lift2d_forward(x, self.lift_h.descriptor(), want_lh_rows=True)

    def forward_lift_2d_bands(self, x):
        """Same without materialising the row-pass 'l'/'h' entries (nothing reads them)."""
        from .. import train
        if train.needs_grad(x, self):
            return train.lift2d_forward(self, x)
        return ops.lift2d_forward(x, self.lift_h.descriptor(), want_lh_rows=False)

    def backward_lift_2d(self, subbands, ll_div=1.0, q=1.0):
        """wavelet_transform.py:45-57; ll_div / q fuse dequantize_subbands (pWave.py:191-202)."""
        from .. import train
        if train.needs_grad(subbands["ll"], subbands["lh"], subbands["hl"], subbands["hh"], self):
            sb = {"ll": subbands["ll"] / ll_div, "lh": subbands["lh"] / q, "hl": subbands["hl"] / q, "hh": subbands["hh"] / q}
            return train.lift2d_backward(self, sb)
        return ops.lift2d_backward(subbands["ll"], subbands["lh"], subbands["hl"], subbands["hh"],
                                   self.lift_h.descriptor(), ll_div, q)
