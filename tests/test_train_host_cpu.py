"""Host logic of the training path on CPU: the graphs `BE._forward_autograd` and `SynthesisModule._forward_autograd`
record, with the one CUDA-only node (dge_b200.autograd.conv2d -> tcgen05 kernels) swapped for ATen's conv, against the
golden fixtures of the unmodified reference and autograd through the oracle.  The kernels themselves are covered on
the GPU (tests/test_conv_gpu.py, tests/test_train_gpu.py); the public entry points still refuse CPU tensors."""
import os

import pytest
import torch
import torch.nn.functional as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def _check_pin(g, pin, tol, name=""):
    """Compare a gradient with the compact pin tests/golden/make_golden.py::training_grads stored for it."""
    if "full" in pin:
        assert rel(g, pin["full"]) < tol, name
        return
    assert tuple(g.shape) == pin["shape"], name
    assert abs(g.norm().double().item() - pin["l2"]) <= tol * pin["l2"], name
    scale = pin["l2"] / g.numel() ** 0.5
    assert abs(g.double().sum().item() - pin["sum"]) <= tol * max(abs(pin["sum"]), scale * g.numel() ** 0.5), name
    assert (g.flatten()[:16] - pin["head"]).abs().max().item() <= tol * max(pin["head"].abs().max().item(), scale), name


@pytest.fixture(scope="module")
def ref_grads():
    """Gradients produced by `loss.backward()` through the unmodified reference (train_grads.pt)."""
    return torch.load(os.path.join(GOLD, "train_grads.pt"))


@pytest.fixture()
def aten_conv(monkeypatch):
    import model.E.E as EM
    import model.stylegan2_generator as SG

    def conv(x, w, planes=2):
        return F.conv2d(x, w, padding=w.shape[-1] // 2)

    import model.stylegan1.net as S1
    import model.E.E_Blur as EB
    import model.E.E_BIG as EG
    import lpips as LP
    monkeypatch.setattr(LP.tc, "conv2d", conv)
    import model.biggan_generator as BG
    monkeypatch.setattr(EG.tc, "conv2d", conv)
    monkeypatch.setattr(BG.tc, "conv2d", conv)
    # these tests hold the graphs of separate torch nodes (the cross-check of the fused nodes, tests/test_train_fused_cpu.py)
    monkeypatch.setattr(EG, "FUSED_TRAIN", False)
    monkeypatch.setattr(BG, "FUSED_TRAIN", False)
    monkeypatch.setattr(EB.tc, "conv2d", conv)
    monkeypatch.setattr(EM.tc, "conv2d", conv)
    monkeypatch.setattr(SG.tc, "conv2d", conv)
    monkeypatch.setattr(S1.tc, "conv2d", conv)


def test_encoder_graph_matches_reference_gradients(aten_conv):
    from model.E.E import BE
    fx = torch.load(os.path.join(GOLD, "be_s16_l4.pt"))
    gx = torch.load(os.path.join(GOLD, "be_s16_l4_grads.pt"))
    E = BE(**fx["config"])
    E.load_state_dict(fx["state_dict"], strict=True)
    torch.manual_seed(fx["noise_seed"])
    const, w = E._forward_autograd(fx["img"], 9)
    assert rel(const, fx["const"]) < 2e-5 and rel(w, fx["w"]) < 2e-5
    (((const - gx["t_const"]) ** 2).mean() + ((w - gx["t_w"]) ** 2).mean()).backward()
    got = {k: p.grad for k, p in E.named_parameters() if p.grad is not None}
    assert set(got) == set(gx["grads"])
    for k, g in gx["grads"].items():
        assert rel(got[k], g) < 1e-4, k


def test_synthesis_graph_matches_oracle_gradient(aten_conv, ref_grads):
    from model.stylegan2_generator import StyleGAN2Generator
    from oracle import stylegan2 as osg2
    fx = torch.load(os.path.join(GOLD, "sg2_res32.pt"))
    G = StyleGAN2Generator(**fx["config"])
    G.load_state_dict(fx["state_dict"], strict=True)
    wp = fx["wp"].clone().requires_grad_(True)
    out = G.synthesis._forward_autograd(wp)
    assert rel(out["image"], fx["image"]) < 2e-5
    target = torch.randn(out["image"].shape, generator=torch.Generator().manual_seed(1))
    ((out["image"] - target) ** 2).mean().backward()
    assert all(p.grad is None for p in G.parameters())       # frozen generator: constants
    wp_r = fx["wp"].clone().requires_grad_(True)
    ref = osg2.synthesis(fx["state_dict"], wp_r, fx["config"]["resolution"])
    ((ref["image"] - target) ** 2).mean().backward()
    assert set(out) == set(ref)
    assert rel(wp.grad, wp_r.grad) < 1e-5
    assert rel(wp.grad, ref_grads["sg2_dwp"]) < 1e-4          # ... and the reference's own backward
    torch.manual_seed(77)
    out_rn = G.synthesis._forward_autograd(fx["wp"].clone().requires_grad_(True), randomize_noise=True)
    assert rel(out_rn["image"], fx["image_randnoise_seed77"]) < 2e-5


def test_stylegan1_graph_matches_fixture_and_oracle_gradient(aten_conv, ref_grads):
    from model.stylegan1.net import Generator
    from oracle import stylegan1 as osg1
    fx = torch.load(os.path.join(GOLD, "sg1_l6.pt"))
    Gs = Generator(**fx["config"])
    Gs.load_state_dict(fx["state_dict"], strict=True)
    for lod, img in fx["images"].items():
        styles = fx["styles"].clone().requires_grad_(True)
        torch.manual_seed(60 + lod)
        out = Gs._decode_autograd(styles, lod)
        assert rel(out, img) < 2e-5, lod
        target = torch.randn(out.shape, generator=torch.Generator().manual_seed(2))
        ((out - target) ** 2).mean().backward()
        styles_r = fx["styles"].clone().requires_grad_(True)
        torch.manual_seed(60 + lod)
        ((osg1.decode(fx["state_dict"], styles_r, lod) - target) ** 2).mean().backward()
        assert rel(styles.grad, styles_r.grad) < 1e-5, lod
        assert rel(styles.grad, ref_grads["sg1_dstyles"][lod]) < 1e-4, lod
    assert all(p.grad is None for p in Gs.parameters())


def test_e_blur_graph_matches_fixture_and_oracle_gradients(aten_conv, ref_grads):
    """Case-2 encoder (strided transform_kernel convs + blur).  Two fp32 evaluations of this chain differ by ~6e-5 from
    the fp64 value on the 4x4 blocks, hence 3e-4 here."""
    from model.E.E_Blur import BE
    from oracle import encoder as oenc
    fx = torch.load(os.path.join(GOLD, "e_blur_s16_l6.pt"))
    E = BE(**fx["config"])
    E.load_state_dict(fx["state_dict"], strict=True)
    torch.manual_seed(fx["noise_seed"])
    const, w = E._forward_autograd(fx["img"], 9)
    assert rel(const, fx["const"]) < 2e-5 and rel(w, fx["w"]) < 2e-5
    (const.sum() + (w ** 2).mean()).backward()
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fx["state_dict"].items()}
    torch.manual_seed(fx["noise_seed"])
    const_r, w_r = oenc.be_blur_forward(sd, fx["img"], fx["config"]["layer_count"])
    (const_r.sum() + (w_r ** 2).mean()).backward()
    checked = 0
    for k, p in E.named_parameters():
        if sd[k].grad is not None:
            assert rel(p.grad, sd[k].grad) < 3e-4, k
            _check_pin(p.grad, ref_grads["e_blur"][k], 5e-4, k)
            checked += 1
    assert checked > 50 and checked == len(ref_grads["e_blur"])


def test_biggan_graph_matches_fixture_and_oracle_gradient(aten_conv, ref_grads):
    from model.biggan_generator import BigGAN
    from model.utils.biggan_config import BigGANConfig
    from oracle import biggan as obg
    fx = torch.load(os.path.join(GOLD, "biggan_small.pt"))
    G = BigGAN(BigGANConfig.from_dict(fx["config"]))
    G.load_state_dict(fx["state_dict"], strict=True)
    G.eval()
    for trunc, img in fx["images"].items():
        z = fx["z"].clone().requires_grad_(True)
        cond = torch.cat((z, F.linear(fx["label"], G.embeddings.weight.detach())), dim=1)
        out = G.generator._forward_autograd(cond, trunc)
        assert rel(out, img) < 2e-5, trunc
        target = torch.randn(out.shape, generator=torch.Generator().manual_seed(4))
        ((out - target) ** 2).mean().backward()
        z_r = fx["z"].clone().requires_grad_(True)
        ref, _ = obg.biggan(fx["state_dict"], fx["config"], z_r, fx["label"], trunc)
        ((ref - target) ** 2).mean().backward()
        assert rel(z.grad, z_r.grad) < 1e-5, trunc
        assert rel(z.grad, ref_grads["biggan_dz"][trunc]) < 1e-4, trunc
    assert all(p.grad is None for p in G.parameters())


def test_e_big_graph_matches_fixture_and_oracle_gradients(aten_conv, ref_grads):
    """BigGAN encoder: the conditional-BN scale / offset layers are spectral-norm wrapped AND trainable here."""
    from model.E.E_BIG import BE
    from oracle import biggan as obg
    fx = torch.load(os.path.join(GOLD, "e_big_s16_l4.pt"))
    E = BE(**fx["config"])
    E.load_state_dict(fx["state_dict"], strict=True)
    E.eval()
    torch.manual_seed(13)
    f = E._features_autograd(fx["img"], fx["cond"])
    assert rel(f, fx["features_seed13"]) < 2e-5
    (f ** 2).mean().backward()
    frozen = ("_u", "_v", "running_means", "running_vars")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(frozen))
          for k, v in fx["state_dict"].items()}
    torch.manual_seed(13)
    (obg.e_big_features(sd, fx["img"], fx["cond"], fx["config"]["layer_count"]) ** 2).mean().backward()
    checked = 0
    for k, p in E.named_parameters():
        if p.grad is not None:
            assert rel(p.grad, sd[k].grad) < 1e-4, k
            _check_pin(p.grad, ref_grads["e_big"][k], 5e-4, k)
            checked += 1
    assert checked >= 40 and checked == len(ref_grads["e_big"])


def test_lpips_structure_matches_oracle_and_torchvision(aten_conv):
    """LPIPS-VGG16 stand-in (third-party package absent, parity unpinned): distance and image gradient against the
    oracle restatement of the published algorithm, feature stack against torchvision's VGG16 with the same weights,
    state_dict keys of the package."""
    import lpips
    from oracle import lpips as olp
    from torchvision.models import vgg16
    torch.manual_seed(0)
    m = lpips.LPIPS(net="vgg", pretrained=False, pnet_rand=True, verbose=False)
    with torch.no_grad():
        for k in range(5):
            getattr(m, f"lin{k}").model[1].weight.abs_()
    sd = m.state_dict()
    assert {"scaling_layer.shift", "net.slice1.0.weight", "net.slice5.28.bias", "lin4.model.1.weight",
            "lins.0.model.1.weight"} <= set(sd)
    assert all(not p.requires_grad for p in m.net.parameters())
    g = torch.Generator().manual_seed(1)
    a = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).requires_grad_(True)
    b = torch.rand(2, 3, 64, 64, generator=g) * 2 - 1
    d = m._distance(a, b)
    assert d.shape == (2, 1, 1, 1)
    (ga,) = torch.autograd.grad(d.mean(), a)
    a_r = a.detach().clone().requires_grad_(True)
    d_r = olp.lpips_vgg(sd, a_r, b)
    (ga_r,) = torch.autograd.grad(d_r.mean(), a_r)
    assert rel(d, d_r) < 1e-5 and rel(ga, ga_r) < 1e-4
    assert float(m._distance(b, b).abs().max()) == 0.0
    v = vgg16(weights=None).features
    v.load_state_dict({f"{i}.{n}": sd[f"net.slice{k + 1}.{i}.{n}"] for k, idxs in enumerate(olp.TAPS) for i in idxs
                       for n in ("weight", "bias")})
    x = torch.rand(1, 3, 32, 32, generator=g)
    with torch.no_grad():
        assert rel(m.net(x)[4], v[:30](x)) < 1e-5
    with pytest.raises(Exception):
        m(a, b)                                     # public entry point: CUDA only


def test_differentiable_ssim_matches_oracle():
    import metric.pytorch_ssim as ps
    from oracle import losses as ol
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, 3, 40, 40, generator=g).requires_grad_(True)
    b = torch.rand(2, 3, 40, 40, generator=g)
    s1 = ps._ssim_mean_autograd(a, b)
    (g1,) = torch.autograd.grad(s1, a)
    s2 = ol.ssim(a, b)
    (g2,) = torch.autograd.grad(s2, a)
    assert abs(float(s1.detach()) - float(s2.detach())) < 1e-6 and rel(g1, g2) < 1e-5


def test_public_entry_points_still_refuse_cpu_tensors():
    from dge_b200 import ops
    from model.E.E import BE
    E = BE(startf=16, maxf=32, layer_count=4)
    with pytest.raises(ops.DgeError):
        E(torch.zeros(1, 3, 32, 32))
    from dge_b200 import autograd as tc
    with pytest.raises(ops.DgeError):
        tc.conv2d(torch.zeros(1, 16, 8, 8), torch.zeros(16, 16, 3, 3))


def test_e_blur_margin_fixture_on_the_host_graph_and_oracle(aten_conv):
    """The margin fixture (tests/golden/make_margin_fixtures.py) pins the oracle and the recorded graph like the others."""
    from model.E.E_Blur import BE
    from oracle import encoder as oenc
    fx = torch.load(os.path.join(GOLD, "e_blur_margin.pt"))
    assert fx["min_preactivation_over_std"] > 2e-5
    E = BE(**fx["config"])
    E.load_state_dict(fx["state_dict"], strict=True)
    torch.manual_seed(fx["noise_seed"])
    const, w = E._forward_autograd(fx["img"], 9)
    assert rel(const, fx["const"]) < 2e-5 and rel(w, fx["w"]) < 2e-5
    (const.sum() + (w ** 2).mean()).backward()
    for k, g in fx["grads"].items():
        assert rel(dict(E.named_parameters())[k].grad, g) < 3e-4, k
    torch.manual_seed(fx["noise_seed"])
    with torch.no_grad():
        c_o, w_o = oenc.be_blur_forward(fx["state_dict"], fx["img"], fx["config"]["layer_count"])
    assert rel(c_o, fx["const"]) < 2e-5 and rel(w_o, fx["w"]) < 2e-5
