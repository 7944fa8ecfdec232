"""Temporal predict / update filters (reference: pMCTF/layers/video/wavelet_transform_temporal_mctf.py:11-45)."""
import math

import torch
import torch.nn as nn

from ... import _native as nat
from ... import ops
from ..lifting_1d import PredictUpdate, _HotPathModule, _PackedWeights


class TemporalLifting(_HotPathModule):
    def __init__(self, bitdepth=8, lossy=True, in_channels=1):
        super().__init__()
        self.bitdepth = bitdepth
        self.dynamic_range = float(2 ** bitdepth)
        self.scale = 0.1
        self.in_channels = in_channels
        self.lossy = lossy
        self.P_t = PredictUpdate(in_channels)
        self.U_t = PredictUpdate(in_channels)
        self.scale_p = torch.tensor(1 / math.sqrt(2), requires_grad=True)  # not in the state_dict (:24-25)
        self.scale_u = torch.tensor(0.5, requires_grad=True)
        self._pack = _PackedWeights()

    def descriptor(self) -> nat.Temporal:
        packed = self._pack.get([self.P_t, self.U_t])
        self._keep = packed
        return nat.Temporal(packed.data_ptr(), packed.data_ptr() + 4 * nat.PU_PACKED_FLOATS,
                            float(self.scale_p.detach()), float(self.scale_u.detach()), int(self.lossy),
                            ops.conv_mode_code(self.conv_mode))

    def predict_filter(self, x):
        """(x + 0.1 * P_t(x)) * scale_p   (:27-35)"""
        from ... import train
        if train.needs_grad(x, self):
            return train.temporal_filter(self, x, 0)
        return ops.temporal_filter(x, self.descriptor(), 0)

    def update_filter(self, x):
        """(x + 0.1 * U_t(x)) * scale_u   (:37-45)"""
        from ... import train
        if train.needs_grad(x, self):
            return train.temporal_filter(self, x, 1)
        return ops.temporal_filter(x, self.descriptor(), 1)
