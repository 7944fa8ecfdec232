"""Host mirror of pMCTF/entropy_models/entropy_models.py for the classes pWave++ uses (EntropyCoder :9-55, AEHelper :82-102,
GaussianEncoder :201-284) on the native coder of csrc/pmctf_rans.cu.

Same class names, methods and table construction; what changes is the staging: `GaussianEncoder.encode` / `decode_stream`
on CUDA tensors turn symbols and scales into int16 symbols + int16 table indexes in ONE kernel pass writing straight into
pinned host memory, followed by one stream synchronisation, instead of the reference's two `.to(int16).cpu()` round trips
per coded step (entropy_models.py:37-40)."""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import _native as nat
from ..models import MLCodec_CXX, MLCodec_rans


class EntropyCoder:
    def __init__(self, ec_thread: bool = False, stream_part: int = 1):
        self.encoder = MLCodec_rans.RansEncoder(ec_thread, stream_part)
        self.decoder = MLCodec_rans.RansDecoder(stream_part)

    @staticmethod
    def pmf_to_quantized_cdf(pmf, precision: int = 16):
        return torch.IntTensor(MLCodec_CXX.pmf_to_quantized_cdf(pmf.tolist(), precision))

    @staticmethod
    def pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
        rows = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            q = EntropyCoder.pmf_to_quantized_cdf(torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0), 16)
            rows[i, : q.size(0)] = q
        return rows

    def reset(self):
        self.encoder.reset()

    def encode_with_indexes(self, symbols, indexes, cdf, cdf_length, offset):
        if torch.is_tensor(symbols):
            symbols = symbols.clamp(-30000, 30000).to(torch.int16).cpu().numpy()
        if torch.is_tensor(indexes):
            indexes = indexes.to(torch.int16).cpu().numpy()
        self.encoder.encode_with_indexes(symbols, indexes, cdf, cdf_length, offset)

    def flush(self):
        self.encoder.flush()

    def get_encoded_stream(self) -> bytes:
        return self.encoder.get_encoded_stream().tobytes()

    def set_stream(self, stream):
        self.decoder.set_stream(np.frombuffer(stream, dtype=np.uint8))

    def decode_stream(self, indexes, cdf, cdf_length, offset):
        if torch.is_tensor(indexes):
            indexes = indexes.to(torch.int16).cpu().numpy()
        return torch.Tensor(self.decoder.decode_stream(indexes, cdf, cdf_length, offset))


class AEHelper:
    def __init__(self):
        super().__init__()
        self.entropy_coder = None
        self._offset = None
        self._quantized_cdf = None
        self._cdf_length = None

    def set_entropy_coder(self, coder):
        self.entropy_coder = coder

    def set_cdf_info(self, quantized_cdf, cdf_length, offset):
        self._quantized_cdf = quantized_cdf.cpu().numpy()
        self._cdf_length = cdf_length.reshape(-1).int().cpu().numpy()
        self._offset = offset.reshape(-1).int().cpu().numpy()

    def get_cdf_info(self):
        return self._quantized_cdf, self._cdf_length, self._offset


class _Staging:
    """Grow-only pinned host buffers for the int16 symbols / indexes of one coded step."""

    def __init__(self):
        self.sym = self.idx = None

    def get(self, n: int):
        if self.sym is None or self.sym.numel() < n:
            cap = max(n, 1 << 16)
            self.sym = torch.empty(cap, dtype=torch.int16, pin_memory=True)
            self.idx = torch.empty(cap, dtype=torch.int16, pin_memory=True)
        return self.sym, self.idx


class GaussianEncoder(AEHelper):
    def __init__(self, distribution: str = "laplace"):
        super().__init__()
        if distribution not in ("laplace", "gaussian"):
            raise ValueError(distribution)
        self.distribution = distribution
        self.cdf_distribution = torch.distributions.laplace.Laplace if distribution == "laplace" else torch.distributions.normal.Normal
        self.scale_min = 0.01 if distribution == "laplace" else 0.11
        self.scale_max = 64.0
        self.scale_level = 256
        self.scale_table = self.get_scale_table(self.scale_min, self.scale_max, self.scale_level)
        self.log_scale_min = math.log(self.scale_min)
        self.log_scale_max = math.log(self.scale_max)
        self.log_scale_step = (self.log_scale_max - self.log_scale_min) / (self.scale_level - 1)
        self._staging = _Staging()

    @staticmethod
    def get_scale_table(min_val, max_val, levels):
        return torch.exp(torch.linspace(math.log(min_val), math.log(max_val), levels))

    def update(self, force: bool = False, entropy_coder=None):
        """One table per scale level: the integer support [-c, c] that holds all but 1e-4 of the upper tail, its pmf, and the
        two tails together as the escape symbol (entropy_models.py:227-264)."""
        if entropy_coder is not None:
            self.entropy_coder = entropy_coder
        if not force and self._offset is not None:
            return
        table = self.scale_table
        whole = self.cdf_distribution(torch.zeros_like(table), table)
        centre = torch.full_like(table, 50.0)
        for i in range(50, 1, -1):   # smallest i >= 2 whose cdf exceeds 0.9999, else 50
            centre = torch.where(whole.cdf(torch.full_like(table, float(i))) > 0.9999, torch.full_like(table, float(i)), centre)
        centre = centre.int()
        length = 2 * centre + 1
        width = int(length.max().item())
        grid = (torch.arange(width) - centre[:, None]).float()
        dist = self.cdf_distribution(torch.zeros_like(grid), torch.zeros_like(grid) + table[:, None])
        upper, lower = dist.cdf(grid + 0.5), dist.cdf(grid - 0.5)
        rows = EntropyCoder.pmf_to_cdf(upper - lower, 2 * lower[:, :1], length, width)
        self.set_cdf_info(rows, length + 2, -centre)

    def build_indexes(self, scales):
        scales = torch.maximum(scales, torch.zeros_like(scales) + 1e-5)
        idx = (torch.log(scales) - self.log_scale_min) / self.log_scale_step
        return idx.clamp_(0, self.scale_level - 1).int()

    # ---- device staging ---------------------------------------------------------------------------------------------
    def stage(self, x, scales):
        """CUDA tensors -> (int16 symbols, int16 table indexes) as numpy views of pinned host memory; x may be None."""
        if not scales.is_cuda:
            raise RuntimeError("stage(): CUDA tensors only (the host path is build_indexes + EntropyCoder.encode_with_indexes)")
        sc = scales.contiguous().float()
        if sc.data_ptr() % 16:          # a slice of a larger tensor: the kernel reads 16-byte vectors
            sc = sc.clone()
        n = sc.numel()
        sym_h, idx_h = self._staging.get(n)
        xs = None
        if x is not None:
            xs = x.contiguous().float()
            if xs.data_ptr() % 16:
                xs = xs.clone()
            if xs.numel() != n or xs.device != sc.device:
                raise RuntimeError("symbols and scales must have the same number of elements on one device")
        st = torch.cuda.current_stream(sc.device)
        with torch.cuda.device(sc.device):
            nat.check(nat.lib().pmctf_gaussian_symbolize(xs.data_ptr() if xs is not None else None, sc.data_ptr(), n,
                                                         float(np.float32(self.log_scale_min)), float(np.float32(self.log_scale_step)),
                                                         self.scale_level, sym_h.data_ptr() if xs is not None else None,
                                                         idx_h.data_ptr(), st.cuda_stream), "gaussian_symbolize")
        st.synchronize()   # the one host synchronisation of the coded step
        return (sym_h[:n].numpy() if xs is not None else None), idx_h[:n].numpy()

    def encode(self, x, scales):
        if scales.is_cuda:
            sym, idx = self.stage(x, scales)
            return self.entropy_coder.encode_with_indexes(sym, idx, *self.get_cdf_info())
        return self.entropy_coder.encode_with_indexes(x.reshape(-1), self.build_indexes(scales).reshape(-1), *self.get_cdf_info())

    def decode_stream(self, scales, dtype, device):
        if scales.is_cuda:
            _, idx = self.stage(None, scales)
        else:
            idx = self.build_indexes(scales).reshape(-1)
        val = self.entropy_coder.decode_stream(idx, *self.get_cdf_info())
        return val.reshape(scales.shape).to(dtype).to(device)


# ---- factorised prior of the MV hyper-latent (entropy_models.py:58-199) -------------------------------------------------------
import torch.nn.functional as F  # noqa: E402
from torch import nn  # noqa: E402


class Bitparm(nn.Module):
    def __init__(self, channel, final=False):
        super().__init__()
        self.final = final
        self.h = nn.Parameter(torch.nn.init.normal_(torch.empty(channel).view(1, -1, 1, 1), 0, 0.01))
        self.b = nn.Parameter(torch.nn.init.normal_(torch.empty(channel).view(1, -1, 1, 1), 0, 0.01))
        self.a = None if final else nn.Parameter(torch.nn.init.normal_(torch.empty(channel).view(1, -1, 1, 1), 0, 0.01))

    def forward(self, x):
        x = x * F.softplus(self.h) + self.b
        return x if self.final else x + torch.tanh(x) * torch.tanh(self.a)


class BitEstimator(AEHelper, nn.Module):
    """Per-channel learned CDF (four Bitparm stages + sigmoid); update() tabulates it into rANS tables whose support is where the
    CDF leaves [1e-4, 1 - 1e-4] (entropy_models.py:123-176)."""

    def __init__(self, channel):
        super().__init__()
        self.f1, self.f2, self.f3, self.f4 = Bitparm(channel), Bitparm(channel), Bitparm(channel), Bitparm(channel, True)
        self.channel = channel

    def forward(self, x):
        return self.get_cdf(x)

    def get_logits_cdf(self, x):
        return self.f4(self.f3(self.f2(self.f1(x))))

    def get_cdf(self, x):
        return torch.sigmoid(self.get_logits_cdf(x))

    def update(self, force=False, entropy_coder=None):
        if entropy_coder is not None:
            self.entropy_coder = entropy_coder
        if not force and self._offset is not None:
            return
        with torch.no_grad():
            device = next(self.parameters()).device
            zero = torch.zeros(self.channel, device=device)
            minima, maxima = zero + 50, zero + 50
            for i in range(50, 1, -1):
                lo = torch.squeeze(self.forward((zero - i)[None, :, None, None]))
                minima = torch.where(lo < zero + 0.0001, zero + i, minima)
            for i in range(50, 1, -1):
                hi = torch.squeeze(self.forward((zero + i)[None, :, None, None]))
                maxima = torch.where(hi > zero + 0.9999, zero + i, maxima)
            minima, maxima = minima.int(), maxima.int()
            pmf_start = zero - minima
            pmf_length = maxima + minima + 1
            max_length = pmf_length.max()
            samples = torch.arange(max_length, device=device)[None, :] + pmf_start[:, None, None]
            lower = self.forward(samples - 0.5).squeeze(0)
            upper = self.forward(samples + 0.5).squeeze(0)
            pmf = (upper - lower)[:, 0, :]
            tail_mass = lower[:, 0, :1] + (1.0 - upper[:, 0, -1:])
            self.set_cdf_info(EntropyCoder.pmf_to_cdf(pmf, tail_mass, pmf_length, max_length), pmf_length + 2, -minima)

    @staticmethod
    def build_indexes(size):
        N, Cc, H, W = size
        return torch.arange(Cc, dtype=torch.int).view(1, -1, 1, 1).repeat(N, 1, H, W)

    @staticmethod
    def build_indexes_np(size):
        return BitEstimator.build_indexes(size).cpu().numpy()

    def encode(self, x):
        idx = self.build_indexes(x.size())
        return self.entropy_coder.encode_with_indexes(x.reshape(-1), idx.reshape(-1), *self.get_cdf_info())

    def decode_stream(self, size, dtype, device):
        idx = self.build_indexes((1, self.channel, size[0], size[1]))
        val = self.entropy_coder.decode_stream(idx.reshape(-1), *self.get_cdf_info())
        return val.reshape(idx.shape).to(dtype).to(device)
