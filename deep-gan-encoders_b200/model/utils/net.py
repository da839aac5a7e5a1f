"""Hot-path pieces of the reference's `model/utils/net.py` used by the encoders:
FromRGB (:231-240), downscale2d (:42-43).  (The StyleGAN1 generator lives in model/stylegan1/net.py.)"""
import torch
from torch import nn

import model.utils.lreq as ln
from dge_b200 import ops


class FromRGB(nn.Module):
    """1x1 ln.Conv2d (3 -> outputs, bias) + leaky_relu(0.2); reference model/utils/net.py:231-240."""

    def __init__(self, channels, outputs):
        super().__init__()
        self.from_rgb = ln.Conv2d(channels, outputs, 1, 1, 0)

    def run(self, x):
        """NCHW image -> F32B feature map (one fused kernel)."""
        c = self.from_rgb
        scale = 1.0 if c.implicit_lreq else c.std
        w = c.weight.detach() if scale == 1.0 else c.weight.detach() * scale
        return ops.from_rgb(x.float(), w, c.scaled_bias(), slope=0.2)

    def run_with_stats(self, x, eps=1e-8):
        """-> (F32B feature map, style [N, 2C] = mean||std, mean_rstd): the first block's instance statistics come out
        of the same pass when the fused kernel covers the width (C in 16, 32), else from dge_instance_stats."""
        c = self.from_rgb
        if c.weight.shape[0] not in (16, 32):
            f = self.run(x)
            return (f,) + ops.instance_stats(f, eps)
        scale = 1.0 if c.implicit_lreq else c.std
        w = c.weight.detach() if scale == 1.0 else c.weight.detach() * scale
        return ops.from_rgb_stats(x.float(), w, c.scaled_bias(), slope=0.2, eps=eps)

    def forward(self, x):
        ln._guard('FromRGB', x, self.from_rgb.weight)
        return self.run(x).to_nchw()


def downscale2d(x, factor=2):
    """2x2 average pool of an NCHW tensor (reference :42-43)."""
    ln._guard('downscale2d', x)
    if factor != 2:
        raise NotImplementedError('downscale2d: factor 2 only')
    f = ops.nchw_to_f32b(x.float())
    return ops.blend(f, f, 0.5, 0.5, pool=True).to_nchw()
