"""Flow-based backward warp and chroma motion-vector scaling
(reference: pMCTF/layers/video/video_net.py:32-71)."""
from ... import ops


def torch_warp(feature, flow):
    """grid_sample(bilinear, border, align_corners=True) driven by a pixel-unit flow (video_net.py:32-50)."""
    return flow_warp(feature, flow)


def flow_warp(im, flow):
    """video_net.py:53-55"""
    from ... import train
    if train.needs_grad(im, flow):  # autograd: forward kernel + its adjoint (pmctf_flow_warp_bwd)
        if flow.size(0) != im.size(0) and im.size(0) % flow.size(0) == 0 and flow.size(0) != 1:
            flow = flow.repeat_interleave(im.size(0) // flow.size(0), 0)
        return train.flow_warp(im, flow)
    return ops.flow_warp(im, flow)


def bilineardownsacling(inputfeature, factor=2):
    """2x2 box mean (video_net.py:66-71).  The reference's only use divides the result by 2
    (pMCTF_L.py:317,336,401); the kernel fuses that, so it is undone here for the bare call."""
    if factor != 2 or inputfeature.size(1) != 2:
        raise NotImplementedError("only the factor-2 motion-field case of the hot path is implemented")
    from ... import train
    if train.needs_grad(inputfeature):
        return train.chroma_mv(inputfeature) * 2.0
    return ops.chroma_mv_down(inputfeature) * 2.0


def chroma_mv(mv):
    """bilineardownsacling(mv) / 2 as one kernel."""
    return ops.chroma_mv_down(mv)
