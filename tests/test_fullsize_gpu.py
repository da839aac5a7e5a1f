"""Parity at BASELINE.json's FULL size (configs[2]: StyleGAN2-FFHQ-1024 synthesis + BE(16, 9) at 1024x1024):
the CUDA path against the CPU oracle on the same seeded weights / inputs / noise stream, plus size-independent
properties (instance-norm moments, conv linearity) that exercise every tile of the 1024^2 launches.

Tolerance: 1e-3 of the output scale (the north star's bar for fp32 parity); observed ~2e-5.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3


def _rel(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max() / b.detach().cpu().double().abs().max())


def test_e_plus_g_forward_full_resolution_matches_oracle():
    import bench
    dev = torch.device("cuda")
    G, E = bench.build_ours(dev)                      # seeded random-init weights, zero-init params perturbed
    E.set_noise_mode("reference")                     # draw the per-stage noise on the CPU like the reference
    gsd, esd = bench.oracle_state()                   # the same weights as plain CPU state dicts
    g = torch.Generator().manual_seed(11)
    z = torch.randn(2, 512, generator=g)
    with torch.no_grad():
        imgs1 = G(z.to(dev), trunc_psi=0.7, trunc_layers=8, randomize_noise=False)["image"]
        torch.manual_seed(123)
        const2, w2 = E(imgs1)
        imgs2 = G.synthesis(w2)["image"]
        torch.cuda.synchronize()
        torch.manual_seed(123)
        ref_img2, ref_const2, ref_w2 = bench.oracle_step(gsd, esd, imgs1.cpu())
    assert imgs2.shape == (2, 3, 1024, 1024) and const2.shape == (2, 512, 4, 4) and w2.shape == (2, 18, 512)
    assert _rel(w2, ref_w2) < TOL
    assert _rel(const2, ref_const2) < TOL
    assert _rel(imgs2, ref_img2) < TOL


def test_full_size_kernel_properties():
    from dge_b200 import ops
    n, c, h = 8, 16, 1024
    g = torch.Generator(device="cuda").manual_seed(5)
    # instance norm at 1024^2: every (sample, channel) plane comes out with mean 0 and variance 1 (E.py:58)
    x = ops.F32B(n, c, h, h)
    x.t.copy_(torch.randn(x.t.shape, device="cuda", generator=g) * 3.0 + 1.5)
    style, mr = ops.instance_stats(x)
    xn, _ = ops.instance_norm(x, mr)
    y = xn.to_nchw().double()
    assert float(y.mean(dim=(2, 3)).abs().max()) < 1e-4
    assert float((y.var(dim=(2, 3), unbiased=False) - 1).abs().max()) < 1e-3
    ref_mean = x.to_nchw().double().mean(dim=(2, 3))
    assert float((style[:, :c].double().cpu() - ref_mean.cpu()).abs().max()) < 1e-5
    # the conv is linear in its input over the whole 1024^2 grid (every tile, both M blocks orders, borders)
    wt = torch.randn(c, c, 3, 3, device="cuda", generator=g) * 0.1
    wpk = ops.pack_conv_weight(wt)
    a = torch.randn(2, c, h, h, device="cuda", generator=g)
    b = torch.randn(2, c, h, h, device="cuda", generator=g)

    def conv(t):
        return ops.conv(ops.nchw_to_act(t), wpk, c, ops.CONV_3X3, out_f32b=True)["f32b"].to_nchw()

    lhs = conv(0.75 * a - 1.25 * b)
    rhs = 0.75 * conv(a) - 1.25 * conv(b)
    assert _rel(lhs, rhs) < 2e-4
    # ... and agrees with the fp32 library conv on a border strip and an interior window of the full map
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(a, wt, padding=1)
    out = conv(a)
    for sl in ((slice(0, 40), slice(0, 1024)), (slice(1000, 1024), slice(900, 1024)), (slice(500, 560), slice(480, 560))):
        assert _rel(out[:, :, sl[0], sl[1]], ref[:, :, sl[0], sl[1]]) < 2e-4
