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
        p = self.hp_q_scale[stage_idx]
        if torch.is_grad_enabled() or isinstance(q_index, list):
            return self.get_curr_q(p, q_index)
        cache = self.__dict__.setdefault("_hpq_cache", {})     # the same tensor object per (stage, q_index, parameter version): see pWave.q_pair
        key = (stage_idx, q_index, p.data_ptr(), p._version)
        if key not in cache:
            if len(cache) > 128:
                cache.clear()
            cache[key] = self.get_curr_q(p, q_index)
        return cache[key]

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
            if not hasattr(self, "mv_encoder"):
                raise NotImplementedError("this model was built without the motion path (pMCTF(motion=True, entropy_model=True)): "
                                          "pass mv_hat= (the reference's forward_one_stage accepts it, pMCTF_L.py:333-336)")
            mv_hat, ref_mv, bpp_mv_y, bpp_mv_z = self.compute_and_code_motion(ref_frame, cur_frame, q_index, dpb, stage_idx=stage_idx,
                                                                              me_downsample=me_downsample)
        else:
            bpp_mv_y, bpp_mv_z = None, None
            ref_mv = {"mv_feature": None, "mv_y_hat": None}
            mv_hat = bilineardownsacling(mv_hat) / 2                      # pMCTF_L.py:336
        L_t, H_t, pred_frame, inv_pred_frame = self.forward_MCTF(ref_frame, cur_frame, mv_hat, stage_idx)
        qp_scale = self.hp_qp_scale(stage_idx, q_index)
        res_H = self.hp_coder.forward(H_t, q_index, qp_scale=qp_scale)
        coded_mv = bpp_mv_z is not None
        ret = {"bpp_mv_y": bpp_mv_y, "bpp_mv_z": bpp_mv_z, "bpp_me": bpp_mv_z + bpp_mv_y if coded_mv else None,
               "me_mse": self.mse(pred_frame, cur_frame),
               "bpp": res_H["bpp_total"] + bpp_mv_z + bpp_mv_y if coded_mv else res_H["bpp_total"], "bpp_H": res_H["bpp_total"],
               "bit_H": res_H["bits_total"],
               "bit_ME": (bpp_mv_y + bpp_mv_z) * (ref_frame.size(2) * ref_frame.size(3)) if coded_mv else None,
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

    # ---- motion estimation + MV coding (pMCTF_L.py:211-292, 397-524): host-side sequencing of SpyNet (tensor cores) and the MV codec ----
    @staticmethod
    def _rounded_q(q):
        """stream_helper.get_rounded_q (:35-39): the step that travels implicitly with q_index, rounded to 1/100"""
        import numpy as np
        q = np.clip(q, 0.01, 655.0)
        idx = int(np.round(q * 100))
        return idx / 100, idx

    def get_mv_y_q(self, q_index, stage_idx=0, inference=False):
        enc = self.get_curr_q(self.mv_y_q_scale_enc[stage_idx], q_index)
        dec = self.get_curr_q(self.mv_y_q_scale_dec[stage_idx], q_index)
        if inference:
            enc, dec = self._rounded_q(enc.cpu())[0], self._rounded_q(dec.cpu())[0]
        return enc, dec

    def mv_prior_param_decoder(self, mv_z_hat, dpb, me_num):
        params = self.mv_hyper_prior_decoder[me_num](mv_z_hat)
        if dpb["ref_mv_y"] is None:
            params = self.mv_y_prior_fusion_adaptor_0[me_num](params)
        else:
            params = self.mv_y_prior_fusion_adaptor_1[me_num](torch.cat((params, dpb["ref_mv_y"]), dim=1))
        return self.mv_y_prior_fusion[me_num](params)

    def _me_inputs(self, ref_frame, cur_frame, me_downsample, first_only):
        from ...layers.video.video_net import bilinearupsacling  # noqa: F401
        cur = cur_frame[0] if first_only else cur_frame
        ref = ref_frame[0] if first_only else ref_frame
        mv_cur, mv_ref = cur.tile((1, 3, 1, 1)) / self.dynamic_range, ref.tile((1, 3, 1, 1)) / self.dynamic_range
        if me_downsample > 1:
            import torch.nn.functional as F
            size = (mv_cur.size(2) // me_downsample, mv_cur.size(3) // me_downsample)
            mv_cur = F.interpolate(mv_cur, size, mode="bilinear", align_corners=False)
            mv_ref = F.interpolate(mv_ref, size, mode="bilinear", align_corners=False)
        return mv_cur, mv_ref

    def _adaptors(self, me_num):
        return (self.mv_y_spatial_prior_adaptor_1[me_num], self.mv_y_spatial_prior_adaptor_2[me_num],
                self.mv_y_spatial_prior_adaptor_3[me_num], self.mv_y_spatial_prior[me_num])

    def compute_and_code_motion(self, ref_frame, cur_frame, q_index, dpb, stage_idx=0, me_downsample=1):
        """Flow by SpyNet on the luma planes, analysis into the 1/16-resolution latent, hyper-prior, four-part prior coding (rate
        estimate) and synthesis of the DECODED motion field the lifting uses (pMCTF_L.py:243-292)."""
        from ...layers.video.video_net import bilinearupsacling
        me_num = min(self.num_me_stages - 1, stage_idx)
        q_enc, q_dec = self.get_mv_y_q(q_index, me_num)
        first_only = not (self.training and cur_frame.size(0) != 3)
        mv_cur, mv_ref = self._me_inputs(ref_frame, cur_frame, me_downsample, first_only)
        est_mv = self.optic_flow(mv_cur, mv_ref)
        mv_y = self.mv_encoder[me_num](est_mv, dpb["mv_feature"], q_enc)
        mv_z = self.mv_hyper_prior_encoder[me_num](mv_y)
        mv_z_hat = self.mv_coder.quant(mv_z)
        params = self.mv_prior_param_decoder(mv_z_hat, dpb, me_num)
        y_res, y_q, y_hat, scales_hat = self.mv_coder.forward_four_part_prior(mv_y, params, *self._adaptors(me_num))
        mv_hat, mv_feature = self.mv_decoder[me_num](y_hat, q_dec)
        if me_downsample > 1:
            mv_hat = bilinearupsacling(mv_hat, factor=me_downsample) * me_downsample
        y_for_bit, z_for_bit = (self.em.add_noise(y_res), self.em.add_noise(mv_z)) if self.training else (y_q, mv_z_hat)
        bits_y = self.em.get_y_laplace_bits(y_for_bit, scales_hat)
        bits_z = self.em.get_z_bits(z_for_bit, self.mv_bit_est[me_num])
        px = ref_frame.size(2) * ref_frame.size(3)
        bpp_y, bpp_z = torch.sum(bits_y, dim=(1, 2, 3)) / px, torch.sum(bits_z, dim=(1, 2, 3)) / px
        red = torch.mean if self.training else torch.sum
        return mv_hat, {"mv_feature": mv_feature, "mv_y_hat": y_hat}, red(bpp_y), red(bpp_z)

    @torch.no_grad()
    def compress_mv(self, ref_frame, cur_frame, dpb, stage_idx=0, q_index=0, me_downsample=1):
        """The motion bitstream of one frame pair (pMCTF_L.py:448-497): hyper-latent through the factorised prior, the four
        masked planes of the latent through the Laplace tables."""
        from ...layers.video.video_net import bilinearupsacling
        me_num = min(self.num_me_stages - 1, stage_idx)
        q_enc, q_dec = self.get_mv_y_q(q_index, me_num, inference=True)
        mv_cur, mv_ref = self._me_inputs(ref_frame, cur_frame, me_downsample, first_only=False)
        est_mv = self.optic_flow(mv_cur, mv_ref)
        mv_y = self.mv_encoder[me_num](est_mv, dpb["mv_feature"], q_enc)
        mv_z_hat = torch.round(self.mv_hyper_prior_encoder[me_num](mv_y))
        params = self.mv_prior_param_decoder(mv_z_hat, dpb, me_num)
        out = self.mv_coder.compress_four_part_prior(mv_y, params, *self._adaptors(me_num))
        mv_hat, mv_feature = self.mv_decoder[me_num](out[8], q_dec)
        if me_downsample > 1:
            mv_hat = bilinearupsacling(mv_hat, factor=me_downsample) * me_downsample
        self.em.entropy_coder.reset()
        self.mv_bit_est[me_num].encode(mv_z_hat)
        for k in range(4):
            self.em.gaussian_encoder.encode(out[k], out[4 + k])
        self.em.entropy_coder.flush()
        return {"bit_stream": self.em.entropy_coder.get_encoded_stream(), "mv_hat": mv_hat, "mv_feature": mv_feature, "mv_y_hat": out[8]}

    @torch.no_grad()
    def decompress_mv(self, string, dtype, height, width, dpb, stage_idx=0, q_index=0, me_downsample=1):
        """pMCTF_L.py:499-524"""
        from ...layers.video.video_net import bilinearupsacling
        me_num = min(self.num_me_stages - 1, stage_idx)
        _, q_dec = self.get_mv_y_q(q_index, me_num, inference=True)
        self.em.entropy_coder.set_stream(string)
        device = next(self.parameters()).device
        zh, zw = int((height + 63) // 64 * 64 / 64 + 0.5), int((width + 63) // 64 * 64 / 64 + 0.5)    # stream_helper.get_downsampled_shape
        mv_z_hat = self.mv_bit_est[me_num].decode_stream((zh, zw), dtype, device).to(device)
        params = self.mv_prior_param_decoder(mv_z_hat, dpb, me_num)
        y_hat = self.mv_coder.decompress_four_part_prior(params, *self._adaptors(me_num), gaussian_encoder=self.em.gaussian_encoder)
        mv_hat, mv_feature = self.mv_decoder[me_num](y_hat, q_dec)
        if me_downsample > 1:
            mv_hat = bilinearupsacling(mv_hat, factor=me_downsample) * me_downsample
        return {"mv_hat": mv_hat, "mv_feature": mv_feature, "mv_y_hat": y_hat}

    @torch.no_grad()
    def compress_one_stage(self, ref_frame, cur_frame, code_lt, mv_hat, ischroma, sideinfo=None, file_name=None, stage_idx=0, q_index=0,
                           skip_decoding=False):
        """pMCTF_L.py:397-420: lifting with the decoded field, then the H (and, at the last stage, L) frame through pWave.compress"""
        import os.path as osp
        if ischroma:
            mv_hat = bilineardownsacling(mv_hat) / 2
        L_t, H_t, _, _ = self.forward_MCTF(ref_frame, cur_frame, mv_hat, stage_idx)
        H_hat = self.hp_coder.compress(H_t, sideinfo, file_name, q_index=q_index, skip_decoding=skip_decoding,
                                       qp_scale=self.hp_qp_scale(stage_idx, q_index))
        L_hat = None
        if code_lt:
            name_l = file_name.replace(osp.basename(file_name), "0_C_main.bin" if ischroma else "0_main.bin")
            L_hat = self.lp_coder.compress(L_t, sideinfo, name_l, q_index=q_index, skip_decoding=skip_decoding)
        return {"L_t": L_t, "H_t": H_t, "H_t_hat": H_hat, "L_t_hat": L_hat}

    @torch.no_grad()
    def decompress_one_stage(self, file_name, code_lt, ischroma, psize=128, q_index=0, stage_idx=0):
        """pMCTF_L.py:422-439"""
        import os.path as osp
        pad = psize // 2 if ischroma else psize
        H_t = self.hp_coder.decompress(file_name, padding=pad, q_index=q_index, qp_scale=self.hp_qp_scale(stage_idx, q_index))
        L_t = None
        if code_lt:
            name_l = file_name.replace(osp.basename(file_name), "0_C_main.bin" if ischroma else "0_main.bin")
            L_t = self.lp_coder.decompress(name_l, padding=pad, q_index=q_index)
        return {"L_t": L_t, "H_t": H_t}

    def update(self, force=False):
        """entropy-coder tables of the whole model (pMCTF_L.py:441-446)"""
        if hasattr(self, "em"):
            self.em.update(force)
            for est in self.mv_bit_est:
                est.update(force, entropy_coder=self.em.entropy_coder)
        for coder in (self.lp_coder, self.hp_coder):
            if coder.has_entropy_model():
                coder.update(force)

    def encode_one_stage(self, ref_frame, cur_frame, code_lt, dpb, output_path=None, pic_width=None, pic_height=None, psize=128,
                         skip_decoding=False, stage_idx=0, q_index=0, me_downsample=1):
        """One frame pair, luma then chroma with the luma's decoded field (pMCTF_L.py:525-637).  output_path None: the forward
        (rate-estimate) path -- the reference's own version of this branch reads keys forward_one_stage does not return
        (`result["mv_feature"]`) and cannot run; here it takes them from `result["dpb"]`.  With an output path: the
        motion stream (`*_mv.bin`, stream_helper.encode_p), the luma and chroma frame streams, and -- unless skip_decoding -- the
        decoder's reconstruction read back from those files."""
        import os
        import struct
        import time
        ref_y, ref_c = ref_frame
        cur_y, cur_c = cur_frame
        if output_path is None:
            r = self.forward_one_stage(ref_y, cur_y, q_index, code_lt, dpb, stage_idx=stage_idx, me_downsample=me_downsample)
            rc = self.forward_one_stage(ref_c, cur_c, q_index, code_lt, dpb, mv_hat=r["mv_hat"], stage_idx=stage_idx, me_downsample=me_downsample)
            return {"L_t": r["L_t"], "H_t": r["H_t"], "L_tc": rc["L_t"], "H_tc": rc["H_t"],
                    "bit_L": r["bit_L"] + rc["bit_L"] if code_lt else None, "bit_H": r["bit_H"] + rc["bit_H"],
                    "bit_Lc": rc["bit_L"] if code_lt else None, "bit_Hc": rc["bit_H"], "bit_ME": r["bit_ME"], "mv_hat": r["mv_hat"],
                    "dpb": r["dpb"], "decoding_time": 0, "encoding_time": 0}
        t0 = time.time()
        mv_out = output_path.replace(".bin", "_mv.bin")
        enc_mv = self.compress_mv(ref_y, cur_y, dpb, stage_idx=stage_idx, q_index=q_index, me_downsample=me_downsample)
        with open(mv_out, "wb") as f:                       # stream_helper.encode_p (:181-186): q index (unused, 0), length, stream
            f.write(struct.pack(">H", 0) + struct.pack(">I", len(enc_mv["bit_stream"])) + enc_mv["bit_stream"])
        mv_hat, mv_feature, mv_y_hat = enc_mv["mv_hat"], enc_mv["mv_feature"], enc_mv["mv_y_hat"]
        out_y = self.compress_one_stage(ref_y, cur_y, code_lt, mv_hat, ischroma=False, sideinfo=[1, 1, pic_height, pic_width],
                                        stage_idx=stage_idx, file_name=output_path, q_index=q_index, skip_decoding=skip_decoding)
        name_c = output_path.replace(".bin", "_C_main.bin")
        out_c = self.compress_one_stage(ref_c.to(ref_y.device), cur_c.to(ref_y.device), code_lt, mv_hat, ischroma=True,
                                        sideinfo=[1, 2, pic_height // 2, pic_width // 2], file_name=name_c, stage_idx=stage_idx,
                                        q_index=q_index, skip_decoding=skip_decoding)
        encoding_time = time.time() - t0
        size = lambda n: os.path.getsize(n) * 8.0  # noqa: E731
        base = os.path.basename(output_path)
        bits_H, bits_Hc, bits_me = size(output_path), size(name_c), size(mv_out)
        bits_L = size(output_path.replace(base, "0_main.bin")) if code_lt else None
        bits_Lc = size(output_path.replace(base, "0_C_main.bin")) if code_lt else None
        decoding_time = 0
        if not skip_decoding:
            t0 = time.time()
            blob = open(mv_out, "rb").read()
            (n,) = struct.unpack(">I", blob[2:6])
            dec_mv = self.decompress_mv(blob[6:6 + n], ref_y.dtype, ref_y.size(2), ref_y.size(3), dpb, stage_idx=stage_idx, q_index=q_index)
            mv_hat, mv_feature = dec_mv["mv_hat"], dec_mv["mv_feature"]
            dy = self.decompress_one_stage(output_path, code_lt, ischroma=False, psize=psize, q_index=q_index, stage_idx=stage_idx)
            dc = self.decompress_one_stage(name_c, code_lt, ischroma=True, psize=psize, q_index=q_index, stage_idx=stage_idx)
            decoding_time = time.time() - t0
            L_t, H_t = (dy["L_t"]["x_hat"] if code_lt else out_y["L_t"]), dy["H_t"]["x_hat"]
            L_tc, H_tc = (dc["L_t"]["x_hat"] if code_lt else out_c["L_t"]), dc["H_t"]["x_hat"]
        else:
            L_t, H_t = (out_y["L_t_hat"] if code_lt else out_y["L_t"]), out_y["H_t_hat"]
            L_tc, H_tc = (out_c["L_t_hat"] if code_lt else out_c["L_t"]), out_c["H_t_hat"]
        return {"L_t": L_t, "H_t": H_t, "L_tc": L_tc, "H_tc": H_tc, "bit_H": bits_H + bits_Hc, "bit_L": bits_L + bits_Lc if code_lt else None,
                "bit_Lc": bits_Lc, "bit_Hc": bits_Hc, "bit_ME": bits_me, "mv_hat": mv_hat, "dpb": {"mv_feature": mv_feature, "ref_mv_y": mv_y_hat},
                "decoding_time": decoding_time, "encoding_time": encoding_time}

    @torch.no_grad()
    def code_gop_forward(self, frames_y, frames_c, q_index=12, bin_folder=None, skip_decoding=True, batched=False):
        """The reference's GOP loop (test_pMCTF_flex.py:131-291) on this model: dyadic temporal analysis with motion estimated and
        coded pair by pair (encode_one_stage), then the temporal synthesis.  frames_y: list of G padded luma planes [1,1,H,W],
        frames_c: list of G chroma pairs [2,1,H/2,W/2].  bin_folder None: the forward (rate-estimate) path.  Returns the
        reconstructed frames and the per-frame bit counts.
        batched=True (forward path only): the same results with the H frames of a stage coded as ONE batch per coder call -- motion
        estimation / coding and the lifting stay pair by pair (the MV codec chains through `dpb`), but pWave.forward is independent per
        frame, so the 8 / 4 / 2 / 1 luma planes (and twice as many chroma planes) of a stage share every launch."""
        import os
        G = len(frames_y)
        stages = G.bit_length() - 1
        coded = [[frames_y[i], frames_c[i], None] for i in range(G)]
        bits = [0.0] * G
        n = G
        if batched and bin_folder is not None:
            raise RuntimeError("batched=True is the forward (rate-estimate) path; the bitstream path writes one file per frame")
        for stage in range(stages):
            n //= 2
            dpb = {"mv_feature": None, "ref_mv_y": None}
            step = 1 << stage
            me_num = min(self.num_me_stages - 1, stage)
            code_lt = stage + 1 == stages
            if batched:
                Hy, Hc, me_bits = [], [], []
                px = frames_y[0].size(2) * frames_y[0].size(3)
                for grp in range(n):
                    i = grp * 2 * step
                    mv_hat, ref_mv, bpp_y, bpp_z = self.compute_and_code_motion(coded[i][0], coded[i + step][0], q_index, dpb, stage_idx=me_num)
                    dpb = {"mv_feature": ref_mv["mv_feature"], "ref_mv_y": ref_mv["mv_y_hat"]}
                    L_y, H_y, _, _ = self.forward_MCTF(coded[i][0], coded[i + step][0], mv_hat, me_num)
                    L_c, H_c, _, _ = self.forward_MCTF(coded[i][1], coded[i + step][1], bilineardownsacling(mv_hat) / 2, me_num)
                    coded[i] = [L_y, L_c, None]
                    coded[i + step][2] = mv_hat
                    Hy.append(H_y), Hc.append(H_c), me_bits.append((bpp_y + bpp_z) * px)
                qp = self.hp_qp_scale(me_num, q_index)
                ry = self.hp_coder.forward(torch.cat(Hy, 0), q_index, qp_scale=qp)
                rc = self.hp_coder.forward(torch.cat(Hc, 0), q_index, qp_scale=qp)
                by, bc = ry["bits"]["bits_total"], rc["bits"]["bits_total"]
                for grp in range(n):
                    j = grp * 2 * step + step
                    coded[j][0], coded[j][1] = ry["x_hat"][grp:grp + 1], rc["x_hat"][2 * grp:2 * grp + 2]
                    # the reference's per-call `bits_total` is the batch MEAN (pWave.py:308): luma + mean of the two chroma planes
                    bits[j] = float(by[grp]) + float(bc[2 * grp:2 * grp + 2].mean()) + float(me_bits[grp])
                if code_lt:
                    rl, rlc = self.lp_coder.forward(coded[0][0], q_index), self.lp_coder.forward(coded[0][1], q_index)
                    coded[0][0], coded[0][1] = rl["x_hat"], rlc["x_hat"]
                    bits[0] = float(rl["bits_total"]) + float(rlc["bits_total"])
                continue
            for grp in range(n):
                i = grp * 2 * step
                path = os.path.join(bin_folder, f"{i + step}.bin") if bin_folder else None
                r = self.encode_one_stage([coded[i][0], coded[i][1]], [coded[i + step][0], coded[i + step][1]], code_lt, dpb, output_path=path,
                                          pic_height=frames_y[0].size(2), pic_width=frames_y[0].size(3), skip_decoding=skip_decoding,
                                          stage_idx=me_num, q_index=q_index)
                coded[i] = [r["L_t"], r["L_tc"], None]
                coded[i + step] = [r["H_t"], r["H_tc"], r["mv_hat"]]
                dpb = r["dpb"]
                bits[i + step] = float(r["bit_H"]) + float(r["bit_ME"])
                if code_lt:
                    bits[i] = float(r["bit_L"])
        for stage in reversed(range(stages)):
            step = 1 << stage
            for grp in reversed(range(G // (2 * step))):
                i = grp * 2 * step
                me_num = min(self.num_me_stages - 1, stage)
                mv = coded[i + step][2]
                ry, cy = self.inverse_MCTF(coded[i][0], coded[i + step][0], mv, stage_idx=me_num)
                rc, cc = self.inverse_MCTF(coded[i][1], coded[i + step][1], mv, stage_idx=me_num, downscale=True)
                coded[i], coded[i + step] = [ry, rc, None], [cy, cc, None]
        return [c[0] for c in coded], [c[1] for c in coded], bits

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
                 postprocess=False, entropy_model=False, motion=False, **kwargs):
        """motion=True + entropy_model=True builds the reference's WHOLE module tree (pMCTF_L.py:36-112): strict state_dict parity
        with its checkpoints (3 224 entries for num_me_stages = 4), forward_one_stage without a given motion field, compress_mv /
        decompress_mv, compress_one_stage / decompress_one_stage, update()."""
        super().__init__()
        self.bitdepth = bitdepth
        self.dynamic_range = 2 ** bitdepth - 1
        self.lossy = lossy
        self.lp_coder = pWave(bitdepth, decomp_levels, lossy, postprocess=postprocess, entropy_model=entropy_model)
        self.hp_coder = pWave(bitdepth, decomp_levels, lossy, postprocess=postprocess, entropy_model=entropy_model)
        if motion:
            from ...entropy_models.entropy_models import BitEstimator
            from ...entropy_models.gaussian_model import CompressionModel
            from ...layers.video.four_part_prior import MVCoderQuad
            from ...layers.video.layers import DepthConvBlock
            from ...layers.video.mv_codec import MvDec, MvEnc, get_hyper_dec_model, get_hyper_enc_model
            from ...layers.video.video_net import ME_Spynet
            self.channel_mv = mv = 64
            self.channel_N, self.channel_M = 64, 32
            n = num_me_stages
            self.optic_flow = ME_Spynet(L=6)
            self.mv_encoder = nn.ModuleList([MvEnc(2, mv) for _ in range(n)])
            self.mv_decoder = nn.ModuleList([MvDec(2, mv) for _ in range(n)])
            self.mv_hyper_prior_encoder = nn.ModuleList([get_hyper_enc_model(self.channel_N, mv) for _ in range(n)])
            self.mv_hyper_prior_decoder = nn.ModuleList([get_hyper_dec_model(self.channel_N, mv) for _ in range(n)])
            self.mv_y_prior_fusion_adaptor_0 = nn.ModuleList([DepthConvBlock(mv, mv * 2) for _ in range(n)])
            self.mv_y_prior_fusion_adaptor_1 = nn.ModuleList([DepthConvBlock(mv * 2, mv * 2) for _ in range(n)])
            self.mv_y_prior_fusion = nn.ModuleList([nn.Sequential(DepthConvBlock(mv * 2, mv * 3), DepthConvBlock(mv * 3, mv * 3)) for _ in range(n)])
            self.mv_y_spatial_prior = nn.ModuleList([nn.Sequential(DepthConvBlock(mv * 3, mv * 3), DepthConvBlock(mv * 3, mv * 3),
                                                                   DepthConvBlock(mv * 3, mv * 2)) for _ in range(n)])
            for k in (1, 2, 3):
                setattr(self, f"mv_y_spatial_prior_adaptor_{k}", nn.ModuleList([nn.Conv2d(mv * 4, mv * 3, 1) for _ in range(n)]))
            self.mv_y_q_scale_enc = nn.ParameterList([nn.Parameter(torch.ones((2, 1, 1, 1))) for _ in range(n)])
            self.mv_y_q_scale_dec = nn.ParameterList([nn.Parameter(torch.ones((2, 1, 1, 1))) for _ in range(n)])
            self.mv_bit_est = nn.ModuleList([BitEstimator(mv) for _ in range(n)])
            self.em = CompressionModel(y_distribution="laplace")
            self.mv_coder = MVCoderQuad(enc_dec_quant=True)
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
    four-step entropy-parameter networks, the LL band's autoregressive model, PostProcess and SpyNet are replaced the same way; the
    MV codec and the ConvLSTM context stay the reference's stock torch modules."""
    from ...layers import LiftingScheme2D

    def adopt(dst: nn.Module, src: nn.Module):
        src_params = dict(src.named_parameters(remove_duplicate=False))
        for name, _ in list(dst.named_parameters(remove_duplicate=False)):
            mod, leaf = dst, name
            *path, leaf = name.split(".")
            for p in path:
                mod = getattr(mod, p)
            mod._parameters[leaf] = src_params[name]
        src_bufs = dict(src.named_buffers(remove_duplicate=False))
        for name, _ in list(dst.named_buffers(remove_duplicate=False)):   # e.g. the masks of the LL model's masked convolutions
            if name in src_bufs:
                mod = dst
                *path, leaf = name.split(".")
                for p in path:
                    mod = getattr(mod, p)
                mod._buffers[leaf] = src_bufs[name]
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
                old = cf[lvl]["ll"] if "ll" in cf[lvl] else None   # the LL band's autoregressive model (context_fusion.py:56): evaluation forward on csrc/pmctf_llar.cu
                if old is not None and all(hasattr(old, n) for n in ("maskedConv1", "residualBlocks", "maskedConv2", "convs")) \
                        and getattr(old, "num_features", 0) == 128 and not getattr(old, "context", False):
                    from ...layers.context_fusion import ContextFusionSubband
                    new = ContextFusionSubband(num_features=128, num_parameters=old.num_parameters, context=False, in_channels=1)
                    cf[lvl]["ll"] = adopt(new.to(next(old.parameters()).device), old).train(old.training)
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
