"""A stand-in for the reference's model objects on machines where /root/reference does not exist (the GPU box).

The classes below have the module tree, parameter names / shapes and hot-path attributes of the reference's
`TemporalLifting`, `iWave1D`, `LiftingScheme2D`, `pWave` and `pMCTF` (pMCTF/layers/lifting_1d.py,
layers/wavelet_transform.py, layers/video/wavelet_transform_temporal_mctf.py, models/pWave.py, models/video/pMCTF_L.py),
built from plain torch modules, and their entry points `pWave.forward / forward_one_channel` and
`pMCTF.forward_one_stage` drive the hot path in exactly the reference's CALL SEQUENCE
(pWave.py:231-312, pMCTF_L.py:332-379).  Written for the tests: the out-of-scope networks are replaced by the simplest
stand-ins with the same call signatures (zero-mean context model, no-op PostProcess, no entropy estimate), so that
`accelerate()` can be exercised on a GPU exactly as it would be on the real objects -- it rebinds methods and swaps
submodules by NAME, which is all that matters to it.  Nothing here is used by the product."""
import torch
import torch.nn as nn
import torch.nn.functional as F


# the reference's model file binds these two names at import (pMCTF_L.py:14); accelerate() patches them in the module of
# the model's class -- here: this module.  Stock-torch versions, only used before accelerate().
def flow_warp(im, flow):
    raise RuntimeError("test double: flow_warp must have been replaced by accelerate()")


def bilineardownsacling(x):
    return F.interpolate(x, scale_factor=0.5, mode="bilinear", align_corners=False)


def _conv(cin, cout):
    return nn.Conv2d(cin, cout, 3, padding=1)


class PredictUpdate(nn.Module):
    def __init__(self, in_ch=1):
        super().__init__()
        self.conv1, self.conv2, self.conv3, self.conv4 = _conv(in_ch, 16), _conv(16, 16), _conv(16, 16), _conv(16, in_ch)


class TemporalLifting(nn.Module):
    def __init__(self, lossy=True):
        super().__init__()
        self.lossy = lossy
        self.P_t, self.U_t = PredictUpdate(), PredictUpdate()


class iWave1D(nn.Module):
    def __init__(self):
        super().__init__()
        for n in ("conv_P1", "conv_U1", "conv_P2", "conv_U2"):
            setattr(self, n, nn.Conv2d(1, 1, (3, 1)))
        for n in ("P_1", "P_2", "U_1", "U_2"):
            setattr(self, n, PredictUpdate())


class LiftingScheme2D(nn.Module):
    def __init__(self):
        super().__init__()
        self.lift_h = iWave1D()
        self.lift_v = self.lift_h


class _ZeroMeanContext(nn.Module):
    """Signature of ContextFusionFourStep.forward (context_fusion_4step.py:96-140): -> (s_res, s_q, s_hat, scales)."""

    def forward(self, s_curr, context=None, prev_subband=None):
        s_q = torch.round(s_curr)
        return s_q, s_q, s_q, torch.ones_like(s_q)


class pWave(nn.Module):
    def __init__(self, bitdepth=8, decomp_levels=4, lossy=True):
        super().__init__()
        self.bitdepth, self.decomp_levels, self.lossy = bitdepth, decomp_levels, lossy
        self.dynamic_range = float(2 ** bitdepth)
        self.clip_value = 8192.0
        self.wavelet_transform = LiftingScheme2D()
        self.QP = nn.Parameter(torch.tensor([1 / 32, 1 / 2]).view(2, 1, 1, 1))
        self.QP_ll = nn.Parameter(torch.tensor([1 / 16, 1 / 2]).view(2, 1, 1, 1))
        self.context_fusion = nn.ModuleDict({str(l): nn.ModuleDict({b: _ZeroMeanContext() for b in ("lh", "hl", "hh")})
                                             for l in range(decomp_levels)})
        self.dequantModule = nn.Identity()
        self.calls = []                                                        # names of the hot-path methods, in call order

    @staticmethod
    def get_qp_num():
        return 21

    def get_one_q_scale(self, q_scale, q_index):
        min_q, max_q = q_scale[0:1, :, :, :], q_scale[1:2, :, :, :]
        step = (torch.log(max_q) - torch.log(min_q)) / (self.get_qp_num() - 1)
        return torch.exp(torch.log(min_q) + step * q_index)

    def get_curr_q(self, q_scale, q_index):
        return self.get_one_q_scale(q_scale, q_index)

    def forward(self, x, q_index=None, qp_scale=None):                         # pWave.py:231-242
        if q_index is not None:
            qp, qp_ll = self.get_curr_q(self.QP, q_index), self.get_curr_q(self.QP_ll, q_index)
            if qp_scale is not None:
                qp, qp_ll = qp * qp_scale, qp_ll * qp_scale
            return self.forward_one_channel(x, qp, qp_ll)
        return self.forward_one_channel(x)

    def forward_one_channel(self, x, q_scale=None, q_scale_ll=None):           # pWave.py:244-312, call sequence
        if q_scale is None:
            q_scale, q_scale_ll = self.QP[-1], self.QP_ll[-1]
        top = self.decomp_levels - 1
        self.calls.append("encode")
        y = self.encode(x)
        subbands_hat = {lvl: {} for lvl in range(self.decomp_levels)}
        self.calls.append("quantize_subband")
        ll_hat = torch.round(self.quantize_subband(y[top]["ll"], q_scale_ll))
        subbands_hat[top]["ll"] = ll_hat
        for lvl in range(top, -1, -1):
            for b in ("lh", "hl", "hh"):
                self.calls.append("quantize_subband")
                s_curr = self.quantize_subband(y[lvl][b], q_scale)
                _, _, s_hat, _ = self.context_fusion[str(lvl)][b](s_curr, context=None, prev_subband=None)
                subbands_hat[lvl][b] = s_hat
        self.calls.append("dequantize_subbands")
        rec = self.dequantize_subbands(subbands_hat, q_scale, q_scale_ll)
        self.calls.append("decode")
        x_hat = self.decode(rec)
        if self.lossy:
            x_hat = self.dequantModule(x_hat / self.dynamic_range) * self.dynamic_range
        return {"x_hat": x_hat, "subbands": subbands_hat, "mse": torch.mean((x - x_hat) ** 2)}


class pMCTF(nn.Module):
    def __init__(self, num_me_stages=4, lossy=True, quant_stage=True):
        super().__init__()
        self.num_me_stages, self.lossy, self.quant_stage = num_me_stages, lossy, quant_stage
        self.lp_coder, self.hp_coder = pWave(lossy=lossy), pWave(lossy=lossy)
        self.temporal_filtering = nn.ModuleList([TemporalLifting(lossy) for _ in range(num_me_stages)])
        self.hp_q_scale = nn.ParameterList([nn.Parameter(torch.tensor([1.0, 0.7 - 0.1 * i]).view(2, 1, 1, 1)) for i in range(num_me_stages)])

    @staticmethod
    def get_qp_num():
        return 21

    def get_one_q_scale(self, q_scale, q_index):
        min_q, max_q = q_scale[0:1], q_scale[1:2]
        step = (torch.log(max_q) - torch.log(min_q)) / (self.get_qp_num() - 1)
        return torch.exp(torch.log(min_q) + step * q_index)

    def get_curr_q(self, q_scale, q_index):
        return self.get_one_q_scale(q_scale, q_index)

    def forward_one_stage(self, ref_frame, cur_frame, q_index, code_lt, dpb, mv_hat=None, stage_idx=0):   # pMCTF_L.py:332-379
        assert mv_hat is not None, "the double has no motion estimation"
        mv_hat = bilineardownsacling(mv_hat) / 2
        L_t, H_t, pred_frame, inv_pred_frame = self.forward_MCTF(ref_frame, cur_frame, mv_hat, stage_idx)
        qp_scale = self.get_curr_q(self.hp_q_scale[stage_idx], q_index) if self.quant_stage else None
        res_H = self.hp_coder.forward(H_t, q_index, qp_scale=qp_scale)
        ret = {"H_t": res_H["x_hat"], "mse_H": res_H["mse"], "mv_hat": mv_hat, "subbands_H": res_H["subbands"],
               "me_mse": torch.mean((pred_frame - cur_frame) ** 2)}
        if code_lt:
            res_L = self.lp_coder.forward(L_t, q_index)
            ret["L_t"], ret["subbands_L"] = res_L["x_hat"], res_L["subbands"]
        else:
            ret["L_t"] = L_t
        return ret
