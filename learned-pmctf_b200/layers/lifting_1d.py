"""PredictUpdate, split/merge and the learned 1-D lifting iWave1D
(reference: pMCTF/layers/lifting_1d.py:10-189) on the B200 kernels.

The nn.Module tree, parameter names, shapes and initial values are the reference's, so its
state_dicts load unchanged; forward computation goes to the fused CUDA step kernel
(csrc/pmctf_kernels.cu) through the C ABI."""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _native as nat
from .. import ops
from .layers import conv3x3


def split(x):
    """Polyphase split along rows (lifting_1d.py:10-13); views, no copy."""
    return x[:, :, ::2, :], x[:, :, 1::2, :]


def merge(x_e, x_o):
    """Row interleave (lifting_1d.py:16-22)."""
    n, c, h, w = x_e.shape
    x = torch.empty((n, c, 2 * h, w), dtype=x_e.dtype, device=x_e.device)
    x[:, :, ::2, :] = x_e
    x[:, :, 1::2, :] = x_o
    return x


class _PackedWeights:
    """Device-side repack of PredictUpdate weights, refreshed when any parameter changes.  The key is (storage address,
    tensor version, device) per parameter.  Writes that bypass the version counter (`p.data.copy_(...)`, as some EMA /
    checkpoint code does) are not visible to it: call `invalidate()` (or the owning module's `invalidate_packed()`, which
    load_state_dict does automatically) after such a write.  The block registered with the C library is released before its
    memory is dropped (pmctf_release_pu_weights), so a recycled address can never be launched with stale parameters."""

    def __init__(self):
        self.key = None
        self.buf = None
        self.blocks = 0

    def invalidate(self):
        self.key = None

    def _release(self):
        if self.buf is not None:
            try:
                ops.release_pu(self.buf, self.blocks)
            except Exception:   # interpreter shutdown: the library may be gone already
                pass
            self.buf = None

    def __del__(self):
        self._release()

    # caches hold device pointers: never pickled / deep-copied (torch.save(model), copy.deepcopy(model))
    def __getstate__(self):
        return {"packed": None}   # non-empty: pickle skips __setstate__ for a falsy state

    def __setstate__(self, state):
        self.key, self.buf, self.blocks = None, None, 0

    def __deepcopy__(self, memo):
        return _PackedWeights()

    def get(self, pus):
        params = [p for pu in pus for p in pu.ordered_params()]
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in params)
        if key != self.key:
            dev = params[0].device
            if self.buf is not None and (self.buf.device != dev or self.blocks != len(pus)):
                self._release()
            if self.buf is None:
                self.buf = torch.empty(len(pus) * nat.PU_PACKED_FLOATS, dtype=torch.float32, device=dev)
                self.blocks = len(pus)
            for i, pu in enumerate(pus):   # a repack at the same address replaces the registered parameters
                ops.pack_pu(pu.ordered_params(), self.buf[i * nat.PU_PACKED_FLOATS:(i + 1) * nat.PU_PACKED_FLOATS])
            self.key = key
        return self.buf


class _HotPathModule(nn.Module):
    """Shared plumbing of the modules that own packed weights: `conv_mode` (None = process default, 'ffma' | 'tensor' =
    this module's calls only), cache invalidation on load_state_dict, and pickling without device pointers."""

    conv_mode = None

    def invalidate_packed(self):
        for m in self.modules():
            pk = m.__dict__.get("_pack")
            if pk is not None:
                pk.invalidate()
            if "_desc" in m.__dict__:
                m.__dict__["_desc"] = None

    def _load_from_state_dict(self, *args, **kwargs):
        super()._load_from_state_dict(*args, **kwargs)
        pk = self.__dict__.get("_pack")
        if pk is not None:
            pk.invalidate()
        if "_desc" in self.__dict__:
            self.__dict__["_desc"] = None

    def __getstate__(self):
        st = dict(self.__dict__)
        if "_desc" in st:
            st["_desc"] = None      # ctypes struct with device pointers
        st.pop("_keep", None)
        return st


class PredictUpdate(_HotPathModule):
    """conv1 -> tanh -> conv2 -> tanh -> conv3, + conv1, -> conv4; 1-16-16-16-1 channels (lifting_1d.py:25-49)."""

    def __init__(self, in_ch):
        super().__init__()
        if in_ch != 1:
            raise NotImplementedError("the B200 PredictUpdate kernel is single-channel (the only use in the reference)")
        self.in_channels = in_ch
        num_ch = 16
        self.conv1 = conv3x3(in_ch, num_ch)
        self.conv2 = conv3x3(num_ch, num_ch)
        self.conv3 = conv3x3(num_ch, num_ch)
        self.conv4 = conv3x3(num_ch, 1)
        self._pack = _PackedWeights()

    def ordered_params(self):
        return [self.conv1.weight, self.conv1.bias, self.conv2.weight, self.conv2.bias,
                self.conv3.weight, self.conv3.bias, self.conv4.weight, self.conv4.bias]

    def packed(self) -> torch.Tensor:
        return self._pack.get([self])

    def forward(self, x):
        from .. import train
        if train.needs_grad(x, self):  # autograd: differentiable conv kernels (csrc/pmctf_train.cu)
            return train.predict_update(self, x)
        return ops.predict_update(x, self.packed(), conv_mode=self.conv_mode)


def _skip_conv(init):
    """(3,1) depth-wise conv holding the two learnable skip taps + bias (lifting_1d.py:71-89 via
    convs.get_conv2d :117-138: padding=False, groups=in_channels, weights replaced by init)."""
    conv = nn.Conv2d(1, 1, kernel_size=(3, 1), stride=(1, 1), padding=(0, 0), groups=1)
    conv.weight.data = torch.tensor(init, dtype=torch.float32).view(1, 1, 3, 1)
    return conv


class iWave1D(_HotPathModule):
    """Prediction-first learned lifting along rows with bior4.4-initialised skip taps (lifting_1d.py:52-189)."""

    def __init__(self, in_channels=1, bitdepth=8, lossy=True):
        super().__init__()
        if in_channels != 1:
            raise NotImplementedError("single-channel only (as used by pWave++)")
        self.bitdepth = bitdepth
        self.dynamic_range = float(2 ** bitdepth)
        self.in_channels = in_channels
        self.lossy = lossy
        c = self.lifting_coeffs = [-1.586134342059924, -0.052980118572961, 0.882911075530934, 0.443506852043971,
                                   0.869864451624781, 1.149604398860241]
        self.conv_P1 = _skip_conv([0.0, c[0], c[0]])
        self.conv_U1 = _skip_conv([c[1], c[1], 0.0])
        self.conv_P2 = _skip_conv([0.0, c[2], c[2]])
        self.conv_U2 = _skip_conv([c[3], c[3], 0.0])
        self.P_1 = PredictUpdate(in_channels)
        self.P_2 = PredictUpdate(in_channels)
        self.U_1 = PredictUpdate(in_channels)
        self.U_2 = PredictUpdate(in_channels)
        self.scaling = True
        self.scale_l = torch.tensor(c[5], requires_grad=True)  # outside the state_dict, as in the reference
        self.scale_h = torch.tensor(c[4], requires_grad=True)
        self._pack = _PackedWeights()
        self._tap_key = None
        self._desc = None

    def descriptor(self) -> nat.IWave:
        """C-ABI parameter block; the 16 skip scalars are read back to the host once per weight version."""
        skips = [self.conv_P1, self.conv_U1, self.conv_P2, self.conv_U2]
        key = tuple((m.weight.data_ptr(), m.weight._version, m.bias.data_ptr(), m.bias._version) for m in skips)
        packed = self._pack.get([self.P_1, self.U_1, self.P_2, self.U_2])  # application order
        mode = ops.conv_mode_code(self.conv_mode)
        if key != self._tap_key or self._desc is None or self._desc.pu_packed != packed.data_ptr() or self._desc.conv_mode != mode:
            vals = torch.cat([torch.cat([m.weight.detach().reshape(3), m.bias.detach().reshape(1)]) for m in skips]).tolist()
            d = nat.IWave()
            for i in range(4):
                for j in range(3):
                    d.tap[i][j] = vals[4 * i + j]
                d.bias[i] = vals[4 * i + 3]
            d.pu_packed = packed.data_ptr()
            d.scale_l, d.scale_h = float(self.scale_l.detach()), float(self.scale_h.detach())
            d.dynamic_range = self.dynamic_range
            d.lossy = int(self.lossy)
            d.conv_mode = mode
            self._tap_key, self._desc = key, d
        return self._desc

    def forward_lift(self, x):
        from .. import train
        if train.needs_grad(x, self):
            return train.iwave1d_forward(self, x)
        return ops.iwave1d_forward(x, self.descriptor())

    def backward_lift(self, l, h):
        from .. import train
        if train.needs_grad(l, h, self):
            return train.iwave1d_backward(self, l, h)
        return ops.iwave1d_backward(l, h, self.descriptor())
