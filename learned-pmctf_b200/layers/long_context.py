"""The inter-subband long-term context of pWave++ (reference: pMCTF/layers/long_context.py:8-226): three stacked convolutional
LSTM cells walk over the subbands in coding order and hand each four-step model its `context` plane; between decomposition levels
the states are upsampled by nearest x2 + 3x3 convolution.

58 kFLOP per coefficient (1 % of the four-step network it feeds): host-side torch code on stock convolutions with the
reference's module tree and parameter names (LSTM{1,2,3}.conv_{in,hidden}, deconv_{h,c}{1,2,3}.{lvl}.conv)."""
from __future__ import annotations

import torch
import torch.nn as nn


class LSTM2D(nn.Module):
    """One cell.  All three gates and the candidate are functions of the SAME pre-activation conv_in(x) + conv_hidden(h)
    (long_context.py:16-34): c' = sigmoid(a) * c + sigmoid(a) * tanh(a), h' = sigmoid(a) * tanh(c')."""

    def __init__(self, input_channels, hidden_size):
        super().__init__()
        self.conv_in = nn.Conv2d(input_channels, hidden_size, 3, padding=1)
        self.conv_hidden = nn.Conv2d(hidden_size, hidden_size, 3, padding=1)

    def forward(self, x, hidden, cell_state):
        if x.is_cuda and not torch.is_grad_enabled():      # the gate arithmetic as one kernel instead of eight ATen launches
            from .. import _native as nat
            from .. import ops
            a_in, a_hid = self.conv_in(x).contiguous(), self.conv_hidden(hidden).contiguous()
            c = cell_state.expand_as(a_in).contiguous()    # the cell state of the last cell broadcasts (long_context.py:165-169)
            h_out, c_out = torch.empty_like(a_in), torch.empty_like(a_in)
            ops._launch(x.device, "lstm_gates", nat.lib().pmctf_lstm_gates, a_in.data_ptr(), a_hid.data_ptr(), c.data_ptr(), h_out.data_ptr(),
                        c_out.data_ptr(), a_in.numel())
            return h_out, c_out
        a = self.conv_in(x) + self.conv_hidden(hidden)
        gate = torch.sigmoid(a)
        cell_state = gate * cell_state + gate * torch.tanh(a)
        return gate * torch.tanh(cell_state), cell_state


class UpsampleModule(nn.Module):
    def __init__(self, num_channels, mode="nearest"):
        super().__init__()
        if mode != "nearest":
            raise NotImplementedError("pWave++ uses the nearest-neighbour variant (long_context.py:46,83-94)")
        self.up = nn.Upsample(scale_factor=2, mode="nearest")
        self.conv = nn.Conv2d(num_channels, num_channels, 3, padding=1)

    def forward(self, x):
        return self.conv(self.up(x))


class SubbandContext(nn.Module):
    def __init__(self, in_channels=1, decomp_levels=4, ctx_per_band=True, init_states=False):
        super().__init__()
        self.decomp_levels, self.init_states, self.ctx_per_band, self.in_channels = decomp_levels, init_states, ctx_per_band, in_channels
        self.out_channels = 3 * in_channels if ctx_per_band else in_channels
        self.hidden_size = hidden = 32
        self.LSTM1 = LSTM2D(in_channels, hidden)
        self.LSTM2 = LSTM2D(hidden, hidden)
        self.LSTM3 = LSTM2D(hidden, self.out_channels)
        self.sequential_init = False
        self.lstm1 = self.lstm2 = self.lstm3 = None
        if decomp_levels > 1:
            for name, ch in (("h1", hidden), ("c1", hidden), ("h2", hidden), ("c2", hidden), ("h3", self.out_channels), ("c3", self.out_channels)):
                setattr(self, "deconv_" + name, nn.ModuleList(UpsampleModule(ch) for _ in range(decomp_levels - 1)))

    def context_one_band(self, x, lstm1, lstm2, lstm3):
        h1, c1 = self.LSTM1(x, *lstm1)
        h2, c2 = self.LSTM2(h1, *lstm2)
        h3, c3 = self.LSTM3(h2, *lstm3)
        return [h1, c1], [h2, c2], [h3, c3]

    def _upsample_states(self, idx):
        self.lstm1 = [self.deconv_h1[idx](self.lstm1[0]), self.deconv_c1[idx](self.lstm1[1])]
        self.lstm2 = [self.deconv_h2[idx](self.lstm2[0]), self.deconv_c2[idx](self.lstm2[1])]
        self.lstm3 = [self.deconv_h3[idx](self.lstm3[0]), self.deconv_c3[idx](self.lstm3[1])]

    def init_sequential(self, subband_size, device, init_states=None):
        if init_states:
            self.lstm1, self.lstm2, self.lstm3 = init_states["lstm1"], init_states["lstm2"], init_states["lstm3"]
        else:
            n, _, h, w = subband_size
            z = lambda ch: torch.zeros((n, ch, h, w), dtype=torch.float32, device=device)  # noqa: E731
            self.lstm3 = [z(3 if self.ctx_per_band else subband_size[1]), z(subband_size[1])]   # the cell state broadcasts (long_context.py:165-169)
            self.lstm1 = [z(self.hidden_size), z(self.hidden_size)]
            self.lstm2 = [z(self.hidden_size), z(self.hidden_size)]
        self.sequential_init = True

    def forward_one_subband(self, subband, subband_name, lvl):
        """`subband`: the band reconstructed just before; returns the context for the NEXT band in coding order.  After the last
        band of a level (hh) the states move up to the next finer level (long_context.py:203-226)."""
        self.lstm1, self.lstm2, self.lstm3 = self.context_one_band(subband, self.lstm1, self.lstm2, self.lstm3)
        if subband_name == "hh" and lvl > 0:
            self._upsample_states(lvl - 1)
        init_next = None
        if subband_name == "ll" and lvl == self.decomp_levels - 1:
            init_next = {k: [s.detach().clone() for s in getattr(self, k)] for k in ("lstm1", "lstm2", "lstm3")}
        return {"context": self.lstm3[0], "init_next": init_next}

    def forward(self, subband_dict, init_states=None):
        """all contexts of a decomposition at once (training form, long_context.py:104-157)"""
        top = self.decomp_levels - 1
        ll = subband_dict[top]["ll"]
        self.init_sequential(list(ll.size()), ll.device, init_states)
        hidden = {i: {} for i in range(self.decomp_levels)}
        hidden[top]["lh"] = self.forward_one_subband(ll, "ll", top)["context"]
        for lvl in range(top, -1, -1):
            hidden[lvl]["hl"] = self.forward_one_subband(subband_dict[lvl]["lh"], "lh", lvl)["context"]
            hidden[lvl]["hh"] = self.forward_one_subband(subband_dict[lvl]["hl"], "hl", lvl)["context"]
            ctx = self.forward_one_subband(subband_dict[lvl]["hh"], "hh", lvl)["context"]
            if lvl > 0:
                hidden[lvl - 1]["lh"] = ctx
        return hidden
