"""pMCTF-L temporal lifting on the B200 kernels (reference: pMCTF/models/video/pMCTF_L.py:29-330).

`pMCTF` below is the stand-alone hot-path subset (temporal_filtering, hp_q_scale, lp_coder /
hp_coder transforms) with the reference's constructor and parameter names.  `accelerate(model)`
grafts the same methods onto an instance of the reference's own pMCTF, so the full codec
(test_pMCTF_flex.py) runs with motion estimation, entropy models and post-filter in stock torch
and the lifting path on these kernels -- see INTEGRATION.md."""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from ... import ops
from ...layers.video.video_net import bilineardownsacling, flow_warp
from ...layers.video.wavelet_transform_temporal_mctf import TemporalLifting
from ..pWave import pWave, pWaveTransform


class MCTFMixin:
    num_me_stages: int
    lossy: bool

    def motion_compensation(self, ref_frame, mv):
        return flow_warp(ref_frame, mv)  # pMCTF_L.py:191-193

    @staticmethod
    def get_qp_num():
        return 21

    def get_one_q_scale(self, q_scale, q_index):  # pMCTF_L.py:195-200
        min_q, max_q = q_scale[0:1], q_scale[1:2]
        step = (torch.log(max_q) - torch.log(min_q)) / (self.get_qp_num() - 1)
        return torch.exp(torch.log(min_q) + step * q_index)

    def get_curr_q(self, q_scale, q_index):  # pMCTF_L.py:202-209
        if isinstance(q_index, list):
            return torch.cat([self.get_one_q_scale(q_scale, i) for i in q_index], dim=0)
        return self.get_one_q_scale(q_scale, q_index)

    def _temporal(self, stage_idx):
        return self.temporal_filtering[min(self.num_me_stages - 1, stage_idx)].descriptor()

    def hp_qp_scale(self, stage_idx, q_index):
        """Temporal-layer-adaptive step scaling (pMCTF_L.py:343-347,404-408)."""
        if not self.quant_stage:
            return None
        return self.get_curr_q(self.hp_q_scale[stage_idx], q_index)

    @staticmethod
    def mse(x, y):
        return torch.mean((x - y) ** 2)

    # --- entry points (pMCTF_L.py:294, 332-379) ------------------------------------------------------------
    def forward(self, ref_frame, cur_frame, q_index, code_lt, dpb, stage_idx=0):
        return self.forward_one_stage(ref_frame, cur_frame, q_index, code_lt, dpb, stage_idx=stage_idx)

    def forward_one_stage(self, ref_frame, cur_frame, q_index, code_lt, dpb, mv_hat=None, stage_idx=0, me_downsample=1):
        """pMCTF.forward_one_stage (pMCTF_L.py:332-379), same signature and return keys, for the hot path: the motion field is
        an INPUT here (`mv_hat=`; the reference's own signature already accepts it and then skips SpyNet + the MV codec,
        :333-336 -- note that it halves and 2x2-averages a given field exactly as below).  The keys that only the MV
        codec / entropy model can fill are None / NaN (see pWave.forward_one_channel)."""
        if mv_hat is None:
            raise NotImplementedError("motion estimation + MV coding (SpyNet, SURVEY.md section 8f row 4) are not part of the hot path: "
                                      "pass mv_hat= (the reference's forward_one_stage accepts it, pMCTF_L.py:333-336)")
        bpp_mv_y, bpp_mv_z = None, None
        ref_mv = {"mv_feature": None, "mv_y_hat": None}
        mv_hat = bilineardownsacling(mv_hat) / 2                      # pMCTF_L.py:336
        L_t, H_t, pred_frame, inv_pred_frame = self.forward_MCTF(ref_frame, cur_frame, mv_hat, stage_idx)
        qp_scale = self.get_curr_q(self.hp_q_scale[stage_idx], q_index) if self.quant_stage else None
        res_H = self.hp_coder.forward(H_t, q_index, qp_scale=qp_scale)
        ret = {"bpp_mv_y": bpp_mv_y, "bpp_mv_z": bpp_mv_z, "bpp_me": None, "me_mse": self.mse(pred_frame, cur_frame),
               "bpp": res_H["bpp_total"], "bpp_H": res_H["bpp_total"], "bit_H": res_H["bits_total"], "bit_ME": None,
               "mse_H": res_H["mse"], "mv_hat": mv_hat, "dpb": {"mv_feature": ref_mv["mv_feature"], "ref_mv_y": ref_mv["mv_y_hat"]},
               "H_t": res_H["x_hat"]}
        if code_lt:
            res_L = self.lp_coder.forward(L_t, q_index)
            ret["bpp_L"], ret["bit_L"], ret["mse_L"] = res_L["bpp_total"], res_L["bits_total"], res_L["mse"]
            ret["me_mse_inv"] = self.mse(inv_pred_frame, ref_frame)
            ret["L_t"] = res_L["x_hat"]
        else:
            ret["L_t"] = L_t
        ret["bit"] = ret["bpp"] * (ref_frame.size(2) * ref_frame.size(3))
        return ret

    def forward_MCTF(self, ref_frame, cur_frame, mv_hat, stage_idx=0, mv_down=False, want_pred=True, **out):
        """H = cur - P(warp(ref, mv)); L = ref + U(warp(H, -mv)) -> (L_t, H_t, pred, inv)  (pMCTF_L.py:297-312).
        Two launches: each fuses warp + PredictUpdate CNN + lifting arithmetic.  `mv_down=True`
        takes the luma motion field and applies the chroma 2x2-mean/2 on the fly (pMCTF_L.py:401)."""
        from ... import train
        if train.needs_grad(ref_frame, cur_frame, mv_hat, self.temporal_filtering):
            return train.forward_mctf(self, ref_frame, cur_frame, mv_hat, stage_idx, mv_down)
        return ops.forward_mctf(ref_frame, cur_frame, mv_hat, self._temporal(stage_idx), mv_down, want_pred, **out)

    def inverse_MCTF(self, L_t, H_t, mv_hat, downscale=False, stage_idx=0, **out):
        """ref = L - U(warp(H, -mv)); cur = H + P(warp(ref, mv))  (pMCTF_L.py:314-330)."""
        from ... import train
        if train.needs_grad(L_t, H_t, mv_hat, self.temporal_filtering):
            ref, cur = train.inverse_mctf(self, L_t, H_t, mv_hat, downscale, stage_idx)
            if out.get("out_ref") is not None:  # caller-provided strided outputs (GopCodec.synthesis)
                out["out_ref"].copy_(ref), out["out_cur"].copy_(cur)
            return ref, cur
        return ops.inverse_mctf(L_t, H_t, mv_hat, self._temporal(stage_idx), downscale, **out)


class pMCTF(MCTFMixin, nn.Module):
    def __init__(self, bitdepth=8, decomp_levels=4, lossy=True, two_stage_me=True, num_me_stages=2, quant_stage=True,
                 postprocess=False, entropy_model=False, **kwargs):
        super().__init__()
        self.bitdepth = bitdepth
        self.dynamic_range = 2 ** bitdepth - 1
        self.lossy = lossy
        self.lp_coder = pWave(bitdepth, decomp_levels, lossy, postprocess=postprocess, entropy_model=entropy_model)
        self.hp_coder = pWave(bitdepth, decomp_levels, lossy, postprocess=postprocess, entropy_model=entropy_model)
        self.temporal_filtering = nn.ModuleList([TemporalLifting(lossy=lossy) for _ in range(num_me_stages)])
        self.quant_stage = quant_stage
        if quant_stage:
            self.hp_q_scale = nn.ParameterList([nn.Parameter(torch.ones((2, 1, 1, 1))) for _ in range(num_me_stages)])
        self.two_stage_me = two_stage_me
        self.num_me_stages = num_me_stages

    def load_reference_state_dict(self, sd):
        """Load the hot-path entries of a full reference checkpoint (3224 keys for num_me_stages=4);
        every key this model owns must be present."""
        own = self.state_dict()
        missing = [k for k in own if k not in sd]
        if missing:
            raise KeyError(f"reference state_dict lacks hot-path keys: {missing[:5]} ...")
        return self.load_state_dict({k: sd[k] for k in own}, strict=True)


# grafted onto the reference's objects by accelerate(): the hot-path methods and the helpers they (and GopCodec) call.  The
# reference's own forward / forward_one_channel / forward_one_stage / compress / decompress bodies stay: they sequence the
# out-of-scope networks around these.
_PWAVE_METHODS = ("post_process", "encode", "decode", "decode_dequant", "encode_bands", "quantize_subband", "quantize_subbands",
                  "dequantize_subbands", "dequantize_subband", "spatial_wavelet_dec", "_q_float", "code_planes", "_train", "_round",
                  "q_pair", "_step_table", "band_layout")
_MCTF_METHODS = ("motion_compensation", "forward_MCTF", "inverse_MCTF", "_temporal", "hp_qp_scale")


def accelerate(ref_model):
    """Graft the B200 hot path onto an instance of the REFERENCE pMCTF (pMCTF_L.py:29): its
    TemporalLifting / LiftingScheme2D submodules are replaced by ours with the SAME parameter
    tensors (state_dict keys and values unchanged), and the hot-path methods are rebound.  The
    four-step entropy-parameter networks, PostProcess and SpyNet are replaced the same way; the MV codec, the LL model and the
    ConvLSTM context stay the reference's stock torch modules."""
    from ...layers import LiftingScheme2D

    def adopt(dst: nn.Module, src: nn.Module):
        src_params = dict(src.named_parameters(remove_duplicate=False))
        for name, _ in list(dst.named_parameters(remove_duplicate=False)):
            mod, leaf = dst, name
            *path, leaf = name.split(".")
            for p in path:
                mod = getattr(mod, p)
            mod._parameters[leaf] = src_params[name]
        return dst

    for i, tl in enumerate(ref_model.temporal_filtering):
        ref_model.temporal_filtering[i] = adopt(TemporalLifting(lossy=tl.lossy).to(next(tl.parameters()).device), tl)
    for coder in (ref_model.lp_coder, ref_model.hp_coder):
        wt = coder.wavelet_transform
        new = LiftingScheme2D(bitdepth=coder.bitdepth, lossy=coder.lossy).to(next(wt.parameters()).device)
        coder.wavelet_transform = adopt(new, wt)
        dq = getattr(coder, "dequantModule", None)
        if dq is not None and all(hasattr(dq, n) for n in ("resBlocks", "conv1", "conv2", "conv3")):   # the reference's PostProcess
            from ...layers.postprocessing import PostProcess
            coder.dequantModule = adopt(PostProcess().to(next(dq.parameters()).device), dq)
        cf = getattr(coder, "context_fusion", None)
        if cf is not None:   # the reference's ContextFusionFourStep modules (context_fusion_4step.py:23) -> the tensor-core ones
            from ...layers.context_fusion_4step import ContextFusionFourStep
            for lvl in list(cf.keys()):
                for band in ("lh", "hl", "hh"):
                    old = cf[lvl][band]
                    if all(hasattr(old, n) for n in ("y_hierarchical_prior_enc", "conv1_context", "y_spatial_prior_3_out")) and old.num_ch == 112:
                        new = ContextFusionFourStep(ctx_channels=old.ctx_channels, lossy=old.lossy).to(next(old.parameters()).device)
                        cf[lvl][band] = adopt(new, old).train(old.training)
        for m in _PWAVE_METHODS:
            setattr(coder, m, types.MethodType(getattr(pWaveTransform, m), coder))
    of = getattr(ref_model, "optic_flow", None)
    if of is not None and hasattr(of, "moduleBasic") and all(hasattr(b, "conv5") for b in of.moduleBasic):   # the reference's ME_Spynet
        from ...layers.video.video_net import ME_Spynet
        ref_model.optic_flow = adopt(ME_Spynet(L=of.L).to(next(of.parameters()).device), of).train(of.training)
    for m in _MCTF_METHODS:
        setattr(ref_model, m, types.MethodType(getattr(MCTFMixin, m), ref_model))
    import sys
    mod = sys.modules.get(type(ref_model).__module__)
    if mod is not None:  # the model file binds these names at import (pMCTF_L.py:14)
        mod.flow_warp, mod.bilineardownsacling = flow_warp, bilineardownsacling
    return ref_model
