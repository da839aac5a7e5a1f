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


BIGGAN_DEEP_256 = {"attention_layer_position": 8, "channel_width": 128, "class_embed_dim": 128, "eps": 0.0001,
                   "layers": [[False, 16, 16], [True, 16, 16], [False, 16, 16], [True, 16, 8], [False, 8, 8], [True, 8, 8],
                              [False, 8, 8], [True, 8, 4], [False, 4, 4], [True, 4, 2], [False, 2, 2], [True, 2, 1]],
                   "n_stats": 51, "num_classes": 1000, "output_dim": 256, "z_dim": 128}


def test_biggan_deep_256_full_width_matches_oracle():
    """BASELINE configs[3] generator at its real size (channel_width 128, 12 GenBlocks + attention at 64^2, 256^2 output;
    biggan_generator.py:232-304) against the CPU oracle on the same seeded weights, batch 2, truncation 0.4."""
    from model.biggan_generator import BigGAN
    from model.utils.biggan_config import BigGANConfig
    from oracle import biggan as obg
    torch.manual_seed(0)
    G = BigGAN(BigGANConfig.from_dict(BIGGAN_DEEP_256)).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        named = dict(list(G.named_parameters()) + list(G.named_buffers()))
        for k, p in named.items():
            if k.endswith(("running_means", ".bias")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.2)
            elif k.endswith("running_vars"):
                p.copy_(torch.rand(p.shape, generator=g) + 0.5)
            elif k.endswith("bn.weight"):
                p.copy_(1 + 0.2 * torch.randn(p.shape, generator=g))     # uninitialised memory upstream (:124-125)
            elif k.endswith("gamma"):
                p.fill_(0.7)
        # spectral norm (eval mode divides by u^T W v with the STORED u, v): converge them by power iteration so that the
        # network is well conditioned, as a trained checkpoint's are
        for k, w in named.items():
            if k.endswith("weight_orig"):
                m = w.detach().reshape(w.shape[0], -1)
                u = torch.nn.functional.normalize(torch.randn(m.shape[0], generator=g), dim=0)
                for _ in range(8):
                    v = torch.nn.functional.normalize(m.t() @ u, dim=0)
                    u = torch.nn.functional.normalize(m @ v, dim=0)
                named[k[:-len("orig")] + "u"].copy_(u)
                named[k[:-len("orig")] + "v"].copy_(v)
    sd = {k: v.detach().clone() for k, v in G.state_dict().items()}
    z = torch.randn(2, 128, generator=g).clamp_(-2, 2) * 0.4
    label = torch.zeros(2, 1000)
    label[0, 30] = 1
    label[1, 207] = 1
    with torch.no_grad():
        ref_img, ref_cond = obg.biggan(sd, BIGGAN_DEEP_256, z, label, 0.4)
        Gc = G.cuda()
        img, cond = Gc(z.cuda(), label.cuda(), 0.4)
    assert img.shape == (2, 3, 256, 256) and ref_img.shape == img.shape
    assert _rel(cond, ref_cond) < TOL
    assert _rel(img, ref_img) < TOL


def test_e_blur_16_9_at_1024_matches_oracle():
    """BASELINE configs[4] encoder variant at its real size: E_Blur.BE(16, 512, 9) on a 1024^2 image (E_Blur.py:99-134:
    blur + stride-2 `transform_kernel` convs for the resolutions >= 128) against the CPU oracle, same noise stream."""
    from model.E.E_Blur import BE
    from oracle import encoder as oenc
    torch.manual_seed(0)
    E = BE(16, 512, 9, 512, 3).eval()
    g = torch.Generator().manual_seed(2)
    with torch.no_grad():
        for k, p in E.named_parameters():
            if k.endswith(("bias", "noise_weight_1", "noise_weight_2", "bias_1", "bias_2")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    sd = {k: v.detach().clone() for k, v in E.state_dict().items()}
    img = torch.randn(1, 3, 1024, 1024, generator=g).clamp_(-1, 1)
    with torch.no_grad():
        torch.manual_seed(77)
        ref_const, ref_w = oenc.be_blur_forward(sd, img, 9)
        Ec = E.cuda()
        torch.manual_seed(77)
        const, w = Ec(img.cuda())
    assert const.shape == ref_const.shape and w.shape == (1, 18, 512)
    assert _rel(w, ref_w) < TOL
    assert _rel(const, ref_const) < TOL
