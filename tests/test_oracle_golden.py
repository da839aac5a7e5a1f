"""Pin oracle/ (the CPU restatement) against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only; fp32; tolerance 2e-5 of the output scale."""
import os

import pytest
import torch

from oracle import encoder as oenc
from oracle import stylegan2 as osg2

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-5


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.fixture(scope="module")
def sg2():
    return torch.load(os.path.join(GOLD, "sg2_res32.pt"))


@pytest.fixture(scope="module")
def be():
    return torch.load(os.path.join(GOLD, "be_s16_l4.pt"))


def test_sg2_mapping_truncation(sg2):
    sd = sg2["state_dict"]
    w = osg2.mapping(sd, sg2["z"])
    assert rel(w, sg2["w"]) < TOL
    wp = osg2.truncation(w, sd["truncation.w_avg"], 8, sg2["trunc_psi"], sg2["trunc_layers"])
    assert rel(wp, sg2["wp"]) < TOL


def test_sg2_blocks(sg2):
    sd = sg2["state_dict"]
    for name, b in sg2["blocks"].items():
        if name.startswith("layer"):
            idx = int(name[5:])
            y, style = osg2.modulate_conv_block(sd, f"synthesis.{name}.", b["x"], b["w"], up=(idx % 2 == 1))
        else:
            y, style = osg2.modulate_conv_block(sd, f"synthesis.{name}.", b["x"], b["w"], ksize=1, demodulate=False,
                                                add_noise=False, lrelu=False)
        assert rel(style, b["style"]) < TOL, name
        assert rel(y, b["y"]) < TOL, name


def test_sg2_synthesis_and_generator(sg2):
    sd = sg2["state_dict"]
    out = osg2.synthesis(sd, sg2["wp"], 32)
    assert rel(out["image"], sg2["image"]) < TOL
    for k, v in sg2["styles"].items():
        assert rel(out[k], v) < TOL, k
    full = osg2.generator(sd, sg2["z"], 32, trunc_psi=sg2["trunc_psi"], trunc_layers=sg2["trunc_layers"])
    assert rel(full["image"], sg2["image"]) < TOL


def test_sg2_randomize_noise_stream(sg2):
    """randomize_noise=True: torch.randn(N,1,res,res) per layer in layer order (stylegan2_generator.py:912-913)."""
    sd = sg2["state_dict"]
    torch.manual_seed(77)
    res_of = lambda idx: 4 * 2 ** ((idx + 1) // 2)
    noises = {idx: torch.randn(2, 1, res_of(idx), res_of(idx)) for idx in range(7)}
    out = osg2.synthesis(sd, sg2["wp"], 32, noises=noises)
    assert rel(out["image"], sg2["image_randnoise_seed77"]) < TOL


def test_skip_upsample_is_bilinear_like():
    """the skip upsample equals depthwise conv_transpose2d(fir, stride 2, padding 1)."""
    import torch.nn.functional as F
    x = torch.randn(2, 3, 5, 7)
    k = osg2.fir_kernel(4.0).repeat(3, 1, 1, 1)
    ref = F.conv_transpose2d(x, k, stride=2, padding=1, groups=3)
    assert rel(osg2.upsample_skip(x), ref) < 1e-6


def test_be_from_rgb_and_blocks(be):
    sd = be["state_dict"]
    x = oenc.from_rgb(sd, be["img"])
    assert rel(x, be["from_rgb"]) < TOL
    torch.manual_seed(5)
    for i, b in be["blocks_seed5"].items():
        y, w1, w2 = oenc.be_block(sd, f"decode_block.{i}.", b["x"])
        assert rel(y, b["y"]) < TOL, i
        assert rel(w1, b["w1"]) < TOL, i
        assert rel(w2, b["w2"]) < TOL, i


def test_be_forward(be):
    torch.manual_seed(be["noise_seed"])
    const, w = oenc.be_forward(be["state_dict"], be["img"], be["config"]["layer_count"])
    assert rel(const, be["const"]) < TOL
    assert rel(w, be["w"]) < TOL


def test_be_gradients(be):
    """Training step: autograd through the oracle == loss.backward() through the unmodified reference
    (be_s16_l4_grads.pt, E_align_s2.py:205)."""
    gx = torch.load(os.path.join(GOLD, "be_s16_l4_grads.pt"))
    sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in be["state_dict"].items()}
    torch.manual_seed(be["noise_seed"])
    with torch.enable_grad():
        const, w = oenc.be_forward(sd, be["img"], be["config"]["layer_count"])
        loss = ((const - gx["t_const"]) ** 2).mean() + ((w - gx["t_w"]) ** 2).mean()
        loss.backward()
    assert abs(float(loss) - float(gx["loss"])) < 1e-5 * abs(float(gx["loss"]))
    assert set(k for k, v in sd.items() if v.grad is not None) == set(gx["grads"])
    for k, g in gx["grads"].items():
        assert rel(sd[k].grad, g) < 1e-4, k


def _check_pin(g, pin, tol, name=""):
    """Compare a gradient with the compact pin make_golden.training_grads stored for it."""
    if "full" in pin:
        assert rel(g, pin["full"]) < tol, name
        return
    assert tuple(g.shape) == pin["shape"], name
    assert abs(g.norm().double().item() - pin["l2"]) <= tol * pin["l2"], name
    scale = pin["l2"] / g.numel() ** 0.5
    assert abs(g.double().sum().item() - pin["sum"]) <= tol * max(abs(pin["sum"]), scale * g.numel() ** 0.5), name
    assert (g.flatten()[:16] - pin["head"]).abs().max().item() <= tol * max(pin["head"].abs().max().item(), scale), name


def test_training_gradients_of_the_other_families_match_the_reference():
    """Autograd through the oracle == `loss.backward()` through the unmodified reference (train_grads.pt): StyleGAN2
    synthesis (d/dwp), StyleGAN1 decode (d/dstyles), BigGAN (d/dz), E_Blur and E_BIG (every parameter)."""
    from oracle import biggan as obg
    from oracle import stylegan1 as osg1
    gx = torch.load(os.path.join(GOLD, "train_grads.pt"))
    tgt = lambda shape, seed: torch.randn(shape, generator=torch.Generator().manual_seed(seed))
    with torch.enable_grad():
        fx = torch.load(os.path.join(GOLD, "sg2_res32.pt"))
        wp = fx["wp"].clone().requires_grad_(True)
        img = osg2.synthesis(fx["state_dict"], wp, fx["config"]["resolution"])["image"]
        ((img - tgt(img.shape, 1)) ** 2).mean().backward()
        assert rel(wp.grad, gx["sg2_dwp"]) < 1e-4

        fx = torch.load(os.path.join(GOLD, "sg1_l6.pt"))
        for lod, want in gx["sg1_dstyles"].items():
            st = fx["styles"].clone().requires_grad_(True)
            torch.manual_seed(60 + lod)
            img = osg1.decode(fx["state_dict"], st, lod)
            ((img - tgt(img.shape, 2)) ** 2).mean().backward()
            assert rel(st.grad, want) < 1e-4, lod

        fx = torch.load(os.path.join(GOLD, "biggan_small.pt"))
        for trunc, want in gx["biggan_dz"].items():
            z = fx["z"].clone().requires_grad_(True)
            img, _ = obg.biggan(fx["state_dict"], fx["config"], z, fx["label"], trunc)
            ((img - tgt(img.shape, 4)) ** 2).mean().backward()
            assert rel(z.grad, want) < 1e-4, trunc

        fx = torch.load(os.path.join(GOLD, "e_blur_s16_l6.pt"))
        sd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fx["state_dict"].items()}
        torch.manual_seed(fx["noise_seed"])
        const, w = oenc.be_blur_forward(sd, fx["img"], fx["config"]["layer_count"])
        (const.sum() + (w ** 2).mean()).backward()
        assert {k for k, v in sd.items() if v.grad is not None} == set(gx["e_blur"])
        for k, pin in gx["e_blur"].items():
            _check_pin(sd[k].grad, pin, 5e-4, k)

        fx = torch.load(os.path.join(GOLD, "e_big_s16_l4.pt"))
        frozen = ("_u", "_v", "running_means", "running_vars")
        sd = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith(frozen))
              for k, v in fx["state_dict"].items()}
        torch.manual_seed(13)
        (obg.e_big_features(sd, fx["img"], fx["cond"], fx["config"]["layer_count"]) ** 2).mean().backward()
        assert {k for k, v in sd.items() if v.grad is not None} == set(gx["e_big"])
        for k, pin in gx["e_big"].items():
            _check_pin(sd[k].grad, pin, 5e-4, k)


def test_full_training_iteration_of_the_oracle_matches_the_reference():
    """E_align_s2.py:152-207 through the oracle (encoder -> synthesis -> image-space + latent-space space_loss ->
    backward) == the same chain through the unmodified reference modules and its own `space_loss`
    (train_grads.pt['e2g_iteration']).  This is the checker tests/test_train_gpu.py holds the CUDA path to."""
    from oracle import losses as oloss
    gx = torch.load(os.path.join(GOLD, "train_grads.pt"))["e2g_iteration"]
    fx = torch.load(os.path.join(GOLD, "e2g_res32.pt"))

    def lp(a, b):
        return ((a - b) ** 2).mean(dim=(1, 2, 3), keepdim=True) + 0.1 * (a - b).abs().mean(dim=(1, 2, 3), keepdim=True)

    esd = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in fx["e_state_dict"].items()}
    with torch.enable_grad():
        torch.manual_seed(fx["noise_seed"])
        const2, w2 = oenc.be_forward(esd, fx["imgs1"], fx["e_config"]["layer_count"])
        imgs2 = osg2.synthesis(fx["g_state_dict"], w2, fx["g_config"]["resolution"])["image"]
        l_img, info_img = oloss.space_loss(fx["imgs1"], imgs2, lpips_model=lp)
        l_w, info_w = oloss.space_loss(fx["wp1"], w2, image_space=False)
        (l_img + 0.01 * l_w).backward()
    assert abs(float(l_img.detach()) - gx["l_img"]) < 1e-5 * abs(gx["l_img"])
    assert abs(float(l_w.detach()) - gx["l_w"]) < 1e-5 * abs(gx["l_w"])
    flat = lambda i: list(i[0]) + list(i[1:])
    for got, want in ((info_img, gx["info_img"]), (info_w, gx["info_w"])):
        for u, v in zip(flat(got), flat(want)):
            assert abs(u - v) <= 1e-4 * abs(v) + 1e-7
    assert {k for k, v in esd.items() if v.grad is not None} == set(gx["grads"])
    for k, pin in gx["grads"].items():
        _check_pin(esd[k].grad, pin, 5e-4, k)


def test_e2g_roundtrip():
    fx = torch.load(os.path.join(GOLD, "e2g_res32.pt"))
    gsd, esd = fx["g_state_dict"], fx["e_state_dict"]
    full = osg2.generator(gsd, fx["z"], 32, trunc_psi=0.7, trunc_layers=8) if False else None
    w = osg2.mapping(gsd, fx["z"], num_layers=1)
    wp = osg2.truncation(w, gsd["truncation.w_avg"], 8, 0.7, 8)
    assert rel(wp, fx["wp1"]) < TOL
    imgs1 = osg2.synthesis(gsd, wp, 32)["image"]
    assert rel(imgs1, fx["imgs1"]) < TOL
    torch.manual_seed(fx["noise_seed"])
    c2, w2 = oenc.be_forward(esd, fx["imgs1"], 4)
    assert rel(c2, fx["const2"]) < TOL and rel(w2, fx["w2"]) < TOL
    imgs2 = osg2.synthesis(gsd, fx["w2"], 32)["image"]
    assert rel(imgs2, fx["imgs2"]) < TOL


def _lpips_stand_in(a, b):
    return ((a - b) ** 2).mean(dim=(1, 2, 3))


def _loss_inputs(case):
    g = torch.Generator().manual_seed(case["seed"])
    a = torch.randn(case["shape"], generator=g)
    b = a + 0.3 * torch.randn(case["shape"], generator=g)
    return a, b


def test_space_loss_oracle():
    import warnings
    from oracle import losses as olosses
    cases = torch.load(os.path.join(GOLD, "space_loss.pt"))
    for name, c in cases.items():
        a, b = _loss_inputs(c)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            loss, info = olosses.space_loss(a, b, image_space=c["image_space"], lpips_model=_lpips_stand_in)
        assert abs(float(loss) - c["loss"]) <= 1e-6 * max(1.0, abs(c["loss"])), name
        flat = lambda i: list(i[0]) + list(i[1:])
        for x, y in zip(flat(info), flat(c["info"])):
            assert abs(x - y) <= 1e-6 * max(1.0, abs(y)), name


def test_lreq_adam_oracle():
    from oracle import optim as ooptim
    fx = torch.load(os.path.join(GOLD, "lreq_adam.pt"))
    params = [p.clone() for p in fx["init"]]
    v = [torch.zeros_like(p) for p in params]
    steps = [0] * len(params)
    for gs in fx["grads"]:
        for i, g in enumerate(gs):
            if g is not None:
                steps[i] += 1
        ooptim.lreq_adam_step(params, gs, v, steps, fx["coefs"], fx["lr"], fx["beta2"])
    assert steps == fx["steps"]
    for p, q in zip(params, fx["final"]):
        assert rel(p, q) < 1e-6
    for a, b in zip(v, fx["exp_avg_sq"]):
        assert rel(a, b) < 1e-6


def test_pggan_oracle():
    from oracle import pggan as opg
    fx = torch.load(os.path.join(GOLD, "pggan_res32.pt"))
    sd = fx["state_dict"]
    for lod, img in fx["images"].items():
        assert rel(opg.generator(sd, fx["z"], 32, lod=lod), img) < TOL, lod
    assert rel(opg.conv_block(sd, "layer4", fx["block_up"]["x"], upsample=True), fx["block_up"]["y"]) < TOL
    assert rel(opg.conv_block(sd, "layer3", fx["block_plain"]["x"]), fx["block_plain"]["y"]) < TOL
    assert rel(opg.conv_block(sd, "output1", fx["block_out"]["x"], ksize=1, padding=0, gain=1.0, lrelu=False),
               fx["block_out"]["y"]) < TOL


def test_e_pg_oracle():
    from oracle import pggan as opg
    fx = torch.load(os.path.join(GOLD, "e_pg_s16_l4.pt"))
    sd = fx["state_dict"]
    torch.manual_seed(8)
    for i, b in fx["blocks_seed8"].items():
        assert rel(opg.e_pg_block(sd, f"decode_block.{i}.", b["x"]), b["y"]) < TOL, i
    torch.manual_seed(8)
    assert rel(opg.e_pg_features(sd, fx["img"], 4), fx["features_seed8"]) < TOL


def test_sg1_oracle():
    from oracle import stylegan1 as osg1
    fx = torch.load(os.path.join(GOLD, "sg1_l6.pt"))
    sd = fx["state_dict"]
    styles = osg1.mapping(fx["map_state_dict"], fx["z"], 12, mapping_layers=3, buffer1=fx["buffer1"], coefs=fx["coefs"])
    assert rel(styles, fx["styles"]) < TOL
    for lod, img in fx["images"].items():
        torch.manual_seed(60 + lod)
        assert rel(osg1.decode(sd, fx["styles"], lod), img) < TOL, lod
    torch.manual_seed(9)
    x = sd["const"]
    for i in range(6):
        y = osg1.decode_block(sd, f"decode_block.{i}.", x, fx["styles"][:, 2 * i], fx["styles"][:, 2 * i + 1])
        if i in fx["blocks_seed9"]:
            assert rel(y, fx["blocks_seed9"][i]["y"]) < TOL, i
        x = y


def test_sg1_fused_scale_is_box_sum_of_plain_transposed_conv():
    """The decomposition the CUDA path uses: conv_transpose2d(x, 4-shift-sum(W), stride 2, padding 1) equals the 2x2
    box sum of the raw stride-2/padding-0 transposed conv with the 3x3 weights (lreq.py:127-140)."""
    import torch.nn.functional as F
    from oracle import stylegan1 as osg1
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 6, 7, generator=g)
    w = torch.randn(5, 4, 3, 3, generator=g)
    ref = osg1.conv_transpose_fused(x, w)
    raw = F.conv_transpose2d(x, w, stride=2, padding=0)            # (2H+1) x (2W+1)
    box = raw[:, :, :-1, :-1] + raw[:, :, 1:, :-1] + raw[:, :, :-1, 1:] + raw[:, :, 1:, 1:]
    assert rel(box, ref) < 1e-5


def test_e_blur_oracle():
    fx = torch.load(os.path.join(GOLD, "e_blur_s16_l6.pt"))
    sd = fx["state_dict"]
    torch.manual_seed(fx["noise_seed"])
    const, w = oenc.be_blur_forward(sd, fx["img"], 6)
    assert rel(const, fx["const"]) < TOL and rel(w, fx["w"]) < TOL
    b = fx["block0_seed71"]
    torch.manual_seed(71)
    y, w1, w2 = oenc.be_blur_block(sd, "decode_block.0.", b["x"], True)
    assert rel(y, b["y"]) < TOL and rel(w1, b["w1"]) < TOL and rel(w2, b["w2"]) < TOL
    b = fx["block4_seed72"]
    torch.manual_seed(72)
    y, w1, w2 = oenc.be_blur_block(sd, "decode_block.4.", b["x"], False)
    assert rel(y, b["y"]) < TOL and rel(w1, b["w1"]) < TOL and rel(w2, b["w2"]) < TOL


def test_biggan_oracle():
    from oracle import biggan as obg
    fx = torch.load(os.path.join(GOLD, "biggan_small.pt"))
    sd, cfg = fx["state_dict"], fx["config"]
    for trunc, img in fx["images"].items():
        out, cond = obg.biggan(sd, cfg, fx["z"], fx["label"], trunc)
        assert rel(out, img) < 5e-5, trunc
        assert rel(cond, fx["cond"]) < TOL
    assert rel(obg.self_attn(sd, "generator.layers.2.", fx["attn"]["x"]), fx["attn"]["y"]) < 5e-5
    assert rel(obg.gen_block(sd, "generator.layers.3.", fx["block_up_drop"]["x"], fx["cond"], 0.4, True, cfg["eps"]),
               fx["block_up_drop"]["y"]) < 5e-5


def test_e_big_oracle():
    from oracle import biggan as obg
    fx = torch.load(os.path.join(GOLD, "e_big_s16_l4.pt"))
    sd = fx["state_dict"]
    torch.manual_seed(13)
    for i, b in fx["blocks_seed13"].items():
        assert rel(obg.e_big_block(sd, f"decode_block.{i}.", b["x"], fx["cond"]), b["y"]) < 5e-5, i
    torch.manual_seed(13)
    assert rel(obg.e_big_features(sd, fx["img"], fx["cond"], 4), fx["features_seed13"]) < 5e-5


def test_gradcam_oracle():
    from oracle import gradcam as ogc
    fx = torch.load(os.path.join(GOLD, "gradcam_tiny.pt"))
    idx, mode = ogc.class_index(fx["logits"])
    assert idx.tolist() == fx["index"].tolist() and mode == fx["index_max"]
    pp = ogc.gradcam_pp(fx["pp"]["feature"], fx["pp"]["gradient"], (40, 24))
    assert pp.dtype == torch.float64 and rel(pp, fx["pp"]["out"]) < 1e-9
    base = ogc.gradcam(fx["base"]["feature"], fx["base"]["gradient"], (40, 24))
    assert rel(base, fx["base"]["out"]) < 1e-7
    heat, cam = ogc.mask2cam(fx["pp"]["out"], fx["imgs"])
    assert rel(heat, fx["mask2cam"]["heat"]) < 1e-7 and rel(cam, fx["mask2cam"]["cam"]) < 1e-6
