"""Generate the golden fixtures in this directory by importing the UNMODIFIED reference.

Run once in the dev container (needs /root/reference or $DGE_REF; neither exists on the GPU box):
    python tests/golden/make_golden.py
The fixtures pin oracle/ (tests/test_oracle_golden.py) and, on the GPU, the CUDA path
(tests/test_golden_gpu.py).  Nothing else reads the reference.

Small channel counts keep the files small; every zero-initialised parameter (bias, noise strength,
noise weights, w_avg) is perturbed so that each term of the forward is exercised.
"""
import os
import sys
import types

import torch

REF = os.environ.get("DGE_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    for n in ["matplotlib", "matplotlib.pyplot", "boto3", "botocore", "botocore.exceptions", "lpips", "tensorboardX"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    # the repo's own mirrored `model` package must not shadow the reference
    sys.path[:] = [p for p in sys.path if "deep-gan-encoders_b200" not in p]
    sys.path.insert(0, REF)


def perturb(module, names, gen, scale=0.1):
    with torch.no_grad():
        for k, p in list(module.named_parameters()) + list(module.named_buffers()):
            if any(k.endswith(s) for s in names):
                p.copy_(torch.randn(p.shape, generator=gen) * scale)


def clone_sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def main():
    import_reference()
    import model.stylegan2_generator as sg2
    import model.E.E as E

    torch.set_grad_enabled(False)
    gen = torch.Generator().manual_seed(4321)

    # ---- StyleGAN2, resolution 32, <=64 channels, 64-d latent spaces ---------------------------
    torch.manual_seed(1234)
    cfg = dict(resolution=32, z_space_dim=64, w_space_dim=64, mapping_fmaps=64, fmaps_base=1024, fmaps_max=64)
    G = sg2.StyleGAN2Generator(**cfg).eval()
    perturb(G, ["bias", "noise_strength", "w_avg"], gen)
    z = torch.randn(2, 64, generator=gen)
    out = G(z, trunc_psi=0.7, trunc_layers=4, randomize_noise=False)
    fx = {"config": cfg, "state_dict": clone_sd(G), "z": z, "trunc_psi": 0.7, "trunc_layers": 4,
          "w": out["w"], "wp": out["wp"], "image": out["image"],
          "styles": {k: v for k, v in out.items() if k.startswith("style") or k.startswith("output_style")}}
    # per-block vectors (standalone ModulateConvBlock.forward): plain, up, ToRGB
    blocks = {}
    x = G.synthesis.early_layer(out["wp"][:, 0])
    for idx in range(G.synthesis.num_layers - 1):
        layer = getattr(G.synthesis, f"layer{idx}")
        y, style = layer(x, out["wp"][:, idx])
        blocks[f"layer{idx}"] = {"x": x.clone(), "w": out["wp"][:, idx].clone(), "y": y.clone(), "style": style.clone()}
        x = y
        if idx % 2 == 0:
            o = getattr(G.synthesis, f"output{idx // 2}")
            rgb, style = o(x, out["wp"][:, idx + 1])
            blocks[f"output{idx // 2}"] = {"x": x.clone(), "w": out["wp"][:, idx + 1].clone(), "y": rgb.clone(),
                                           "style": style.clone()}
    fx["blocks"] = blocks
    # randomize_noise=True restated with explicit noise: seed the global generator, record the draws
    torch.manual_seed(77)
    out_rn = G.synthesis(out["wp"], randomize_noise=True)
    fx["image_randnoise_seed77"] = out_rn["image"]
    torch.save(fx, os.path.join(HERE, "sg2_res32.pt"))
    print("sg2_res32.pt: image", tuple(out["image"].shape), float(out["image"].mean()), float(out["image"].std()))

    # ---- encoder BE(startf=16, maxf=64, layer_count=4) on 32x32 images ---------------------------
    torch.manual_seed(4242)
    ecfg = dict(startf=16, maxf=64, layer_count=4, latent_size=512, channels=3)
    Enc = E.BE(**ecfg).eval()
    perturb(Enc, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
    img = out["image"].clone()
    torch.manual_seed(99)
    const, w = Enc(img)
    efx = {"config": ecfg, "state_dict": clone_sd(Enc), "img": img, "noise_seed": 99, "const": const, "w": w}
    # per-block vectors
    x = Enc.FromRGB(img)
    efx["from_rgb"] = x.clone()
    eb = {}
    torch.manual_seed(5)
    for i, blk in enumerate(Enc.decode_block):
        y, w1, w2 = blk(x)
        eb[i] = {"x": x.clone(), "y": y.clone(), "w1": w1.clone(), "w2": w2.clone()}
        x = y
    efx["blocks_seed5"] = eb
    torch.save(efx, os.path.join(HERE, "be_s16_l4.pt"))
    print("be_s16_l4.pt: const", tuple(const.shape), float(const.mean()), float(const.std()), "w", tuple(w.shape))

    # ---- E -> G round trip (the benchmark's data flow, E_align_s2.py:153,160), small ----------------
    # encoder with 64-d latent is impossible (view(...,512) is hard-coded, E.py:131), so use a 512-d G
    torch.manual_seed(2024)
    cfg2 = dict(resolution=32, fmaps_base=512, fmaps_max=32, mapping_layers=1)   # w/z 512-d, channels 32,32,32,16
    G2 = sg2.StyleGAN2Generator(**cfg2).eval()
    perturb(G2, ["bias", "noise_strength", "w_avg"], gen)
    torch.manual_seed(4243)
    E2 = E.BE(startf=16, maxf=32, layer_count=4, latent_size=512, channels=3).eval()
    perturb(E2, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
    z2 = torch.randn(2, 512, generator=gen)
    r1 = G2(z2, trunc_psi=0.7, trunc_layers=8, randomize_noise=False)
    torch.manual_seed(123)
    c2, w2_ = E2(r1["image"])
    img2 = G2.synthesis(w2_)["image"]
    torch.save({"g_config": cfg2, "e_config": dict(startf=16, maxf=32, layer_count=4, latent_size=512, channels=3),
                "g_state_dict": clone_sd(G2), "e_state_dict": clone_sd(E2), "z": z2, "imgs1": r1["image"],
                "wp1": r1["wp"], "noise_seed": 123, "const2": c2, "w2": w2_, "imgs2": img2},
               os.path.join(HERE, "e2g_res32.pt"))
    print("e2g_res32.pt: imgs2", tuple(img2.shape), float(img2.mean()), float(img2.std()))


def stand_in_lpips(a, b):
    """`lpips` (PyPI, unpinned) is absent offline; the fixtures use this deterministic stand-in for the callable."""
    return ((a - b) ** 2).mean(dim=(1, 2, 3))


def losses_and_optimizer():
    import warnings
    warnings.filterwarnings("ignore")
    import training_utils as tu
    from model.utils.custom_adam import LREQAdam
    torch.set_grad_enabled(False)
    cases = {}
    specs = {"img64": ((2, 3, 64, 64), True), "img512x384": ((1, 3, 512, 384), True), "w": ((2, 8, 512), False),
             "const": ((2, 64, 4, 4), False)}
    for i, (name, (shape, image_space)) in enumerate(specs.items()):
        g = torch.Generator().manual_seed(900 + i)
        a = torch.randn(shape, generator=g)
        b = a + 0.3 * torch.randn(shape, generator=g)
        loss, info = tu.space_loss(a, b, image_space=image_space, lpips_model=stand_in_lpips)
        cases[name] = {"shape": shape, "seed": 900 + i, "image_space": image_space, "loss": float(loss), "info": info}
    torch.save(cases, os.path.join(HERE, "space_loss.pt"))
    print("space_loss.pt:", {k: round(v["loss"], 6) for k, v in cases.items()})

    # LREQAdam: 4 tensors (two with lr_equalization_coef), 3 steps, one tensor skipped (grad None) at step 2
    torch.set_grad_enabled(True)
    g = torch.Generator().manual_seed(77)
    shapes = [(16, 8, 3, 3), (64, 32), (1, 32, 1, 1), (70001,)]
    coefs = [0.0589, 0.125, None, None]
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    for p, c in zip(params, coefs):
        if c is not None:
            p.lr_equalization_coef = c
    init = [p.detach().clone() for p in params]
    opt = LREQAdam(params, lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
    grads = []
    for step in range(3):
        gs = []
        for i, p in enumerate(params):
            if step == 1 and i == 2:
                p.grad = None
                gs.append(None)
            else:
                p.grad = torch.randn(p.shape, generator=g)
                gs.append(p.grad.clone())
        grads.append(gs)
        opt.step()
    torch.save({"shapes": shapes, "coefs": coefs, "init": init, "grads": grads, "lr": 0.0015, "beta2": 0.99,
                "final": [p.detach().clone() for p in params],
                "exp_avg_sq": [opt.state[p]["exp_avg_sq"].clone() for p in params],
                "steps": [opt.state[p]["step"] for p in params]}, os.path.join(HERE, "lreq_adam.pt"))
    print("lreq_adam.pt: steps", [opt.state[p]["step"] for p in params])


def pggan():
    import contextlib
    import io
    import model.pggan.pggan_generator as pg
    import model.E.E_PG as EPG
    torch.set_grad_enabled(False)
    gen = torch.Generator().manual_seed(555)
    torch.manual_seed(31)
    cfg = dict(resolution=32, z_space_dim=64, fmaps_base=1024, fmaps_max=64)
    G = pg.PGGANGenerator(**cfg).eval()
    perturb(G, ["bias"], gen)
    z = torch.randn(2, 64, generator=gen)
    fx = {"config": cfg, "state_dict": clone_sd(G), "z": z, "images": {}}
    with contextlib.redirect_stdout(io.StringIO()):
        for lod in (0, 1.0, 0.5, 2):
            fx["images"][lod] = G(z, lod=lod)["image"].clone()
        fx["z_out"] = G(z)["z"].clone()
        x = torch.randn(2, 64, 8, 8, generator=gen)
        fx["block_up"] = {"x": x, "y": G.layer4(x).clone()}
        fx["block_plain"] = {"x": x, "y": G.layer3(x).clone()}
        fx["block_out"] = {"x": x, "y": G.output1(x).clone()}
    torch.save(fx, os.path.join(HERE, "pggan_res32.pt"))
    print("pggan_res32.pt:", {k: tuple(v.shape) for k, v in fx["images"].items()})

    torch.manual_seed(32)
    ecfg = dict(startf=16, maxf=64, layer_count=4, latent_size=512, channels=3, pggan=False)
    E = EPG.BE(**ecfg).eval()
    perturb(E, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias", "instance_norm_3.weight"], gen)
    img = fx["images"][0]
    efx = {"config": ecfg, "state_dict": clone_sd(E), "img": img, "blocks_seed8": {}}
    x = E.FromRGB(img)
    torch.manual_seed(8)
    for i, blk in enumerate(E.decode_block):
        y, _, _ = blk(x)
        efx["blocks_seed8"][i] = {"x": x.clone(), "y": y.clone()}
        x = y
    efx["features_seed8"] = x.clone()
    out = E(img)
    efx["forward_returns"] = [o.clone() for o in out]
    torch.save(efx, os.path.join(HERE, "e_pg_s16_l4.pt"))
    print("e_pg_s16_l4.pt: features", tuple(x.shape), "forward returns", out)


def stylegan1():
    import model.stylegan1.net as sg1
    torch.set_grad_enabled(False)
    gen = torch.Generator().manual_seed(808)
    # 6 blocks: 4..128; the last block (res 128) is a fused-scale (transposed conv) block, the others nearest-up + conv
    torch.manual_seed(41)
    cfg = dict(startf=16, maxf=32, layer_count=6, latent_size=64, channels=3)
    Gs = sg1.Generator(**cfg).eval()
    perturb(Gs, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
    with torch.no_grad():
        Gs.const.copy_(torch.randn(Gs.const.shape, generator=gen))
    torch.manual_seed(42)
    Gm = sg1.Mapping(num_layers=12, mapping_layers=3, latent_size=64, dlatent_size=64, mapping_fmaps=64).eval()
    Gm.buffer1 = torch.randn(12, 64, generator=gen) * 0.1
    coefs = torch.ones(1, 12, 1)
    coefs[:, :6] = 0.7
    z = torch.randn(2, 64, generator=gen)
    styles = Gm(z, coefs)
    fx = {"config": cfg, "state_dict": clone_sd(Gs), "map_state_dict": clone_sd(Gm), "buffer1": Gm.buffer1.clone(),
          "coefs": coefs, "z": z, "styles": styles.clone(), "images": {}}
    for lod in (5, 3, 0):
        torch.manual_seed(60 + lod)
        fx["images"][lod] = Gs.forward(styles, lod).clone()
    # per-block vectors (seed 9): block 0 (const, batch 1), a nearest-up block, the fused-scale block
    x = Gs.const
    blocks = {}
    torch.manual_seed(9)
    for i, blk in enumerate(Gs.decode_block):
        y = blk(x, styles[:, 2 * i], styles[:, 2 * i + 1])
        if i < 3:
            blocks[i] = {"x": x.clone(), "y": y.clone()}
        x = y
    fx["blocks_seed9"] = blocks
    # the fused-scale block alone on a small input (it only depends on the `fused_scale` flag, not the size)
    xs = torch.randn(2, 32, 8, 8, generator=gen)
    torch.manual_seed(10)
    fx["fused_block_seed10"] = {"x": xs, "y": Gs.decode_block[5](xs, styles[:, 10], styles[:, 11]).clone()}
    torch.save(fx, os.path.join(HERE, "sg1_l6.pt"))
    print("sg1_l6.pt:", {k: tuple(v.shape) for k, v in fx["images"].items()}, [b.fused_scale for b in Gs.decode_block])


def e_blur():
    import model.E.E_Blur as EB
    torch.set_grad_enabled(False)
    gen = torch.Generator().manual_seed(2323)
    torch.manual_seed(51)
    # 6 blocks on 128x128 images: blocks 0..3 use the strided transform-kernel conv, block 4 blur+conv+pool, block 5 last
    cfg = dict(startf=16, maxf=32, layer_count=6, latent_size=512, channels=3)
    E = EB.BE(**cfg).eval()
    perturb(E, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
    img = torch.randn(2, 3, 128, 128, generator=gen)
    torch.manual_seed(70)
    const, w = E(img)
    fx = {"config": cfg, "state_dict": clone_sd(E), "img_seed": 2323, "img": img, "noise_seed": 70, "const": const,
          "w": w, "fused": [b.fused_scale for b in E.decode_block]}
    xs = torch.randn(2, 16, 12, 20, generator=gen)      # ragged size through one strided block
    torch.manual_seed(71)
    y, w1, w2 = E.decode_block[0](xs)
    fx["block0_seed71"] = {"x": xs, "y": y.clone(), "w1": w1.clone(), "w2": w2.clone()}
    xs4 = torch.randn(2, 32, 8, 8, generator=gen)
    torch.manual_seed(72)
    y, w1, w2 = E.decode_block[4](xs4)
    fx["block4_seed72"] = {"x": xs4, "y": y.clone(), "w1": w1.clone(), "w2": w2.clone()}
    torch.save(fx, os.path.join(HERE, "e_blur_s16_l6.pt"))
    print("e_blur_s16_l6.pt: const", tuple(const.shape), "w", tuple(w.shape), fx["fused"])


def biggan():
    import model.biggan_generator as bg
    from model.utils.biggan_config import BigGANConfig
    import model.E.E_BIG as EBIG
    torch.set_grad_enabled(False)
    gen = torch.Generator().manual_seed(9090)
    torch.manual_seed(61)
    cfg = dict(output_dim=64, z_dim=16, class_embed_dim=16, channel_width=32, num_classes=10,
               layers=[[False, 16, 16], [True, 16, 8], [True, 8, 4], [True, 4, 2], [True, 2, 1]],
               attention_layer_position=2, eps=1e-4, n_stats=11)
    G = bg.BigGAN(BigGANConfig.from_dict(cfg)).eval()
    for k, p in list(G.named_parameters()) + list(G.named_buffers()):
        if k.endswith("running_means") or k.endswith(".bias"):
            p.copy_(torch.randn(p.shape, generator=gen) * 0.2)
        elif k.endswith("running_vars"):
            p.copy_(torch.rand(p.shape, generator=gen) + 0.5)
        elif k.endswith("bn.weight"):
            p.copy_(1 + 0.2 * torch.randn(p.shape, generator=gen))      # uninitialised memory upstream (:124-125)
        elif k.endswith("gamma"):
            p.fill_(0.7)
    z = torch.randn(2, 16, generator=gen) * 0.4
    label = torch.zeros(2, 10)
    label[0, 3] = 1
    label[1, 7] = 1
    # spectral-norm u/v start random: converge them with train-mode forwards (one power iteration each), then freeze
    G.train()
    for _ in range(30):
        G(z, label, 0.4)
    G.eval()
    fx = {"config": cfg, "state_dict": clone_sd(G), "z": z, "label": label, "images": {}}
    for trunc in (0.4, 0.37):
        img, cond = G(z, label, trunc)
        fx["images"][trunc] = img.clone()
    fx["cond"] = cond.clone()
    x8 = torch.randn(2, 256, 8, 8, generator=gen)
    fx["attn"] = {"x": x8, "y": G.generator.layers[2](x8).clone()}
    fx["block_up_drop"] = {"x": x8, "y": G.generator.layers[3](x8, cond, 0.4).clone()}
    torch.save(fx, os.path.join(HERE, "biggan_small.pt"))
    print("biggan_small.pt: image", tuple(img.shape), float(img.mean()), float(img.std()), "keys", len(fx["state_dict"]))

    torch.manual_seed(62)
    ecfg = dict(startf=16, maxf=64, layer_count=4, latent_size=512, channels=3, biggan=False)
    E = EBIG.BE(**ecfg).eval()
    perturb(E, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
    img32 = torch.randn(2, 3, 32, 32, generator=gen)
    cond256 = torch.randn(2, 256, generator=gen) * 0.3
    E.train()
    for _ in range(30):
        E.features_unused = None
        xx = E.FromRGB(img32)
        for blk in E.decode_block:
            xx, _, _ = blk(xx, cond256, truncation=0.4)
    E.eval()
    efx = {"config": ecfg, "state_dict": clone_sd(E), "img": img32, "cond": cond256, "blocks_seed13": {}}
    x = E.FromRGB(img32)
    torch.manual_seed(13)
    for i, blk in enumerate(E.decode_block):
        y, _, _ = blk(x, cond256, truncation=0.4)
        efx["blocks_seed13"][i] = {"x": x.clone(), "y": y.clone()}
        x = y
    efx["features_seed13"] = x.clone()
    torch.save(efx, os.path.join(HERE, "e_big_s16_l4.pt"))
    print("e_big_s16_l4.pt: features", tuple(x.shape), "keys", len(efx["state_dict"]))


def tiny_vgg():
    """VGG-like stand-in (torchvision VGG16 weights are unavailable offline): the hooked conv is followed by an
    in-place ReLU, as `features.28` is in VGG16."""
    import torch.nn as nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.features = nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(2),
                                          nn.Conv2d(8, 16, 3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(2))
            self.pool = nn.AdaptiveAvgPool2d(2)
            self.classifier = nn.Linear(64, 10)

        def forward(self, x):
            return self.classifier(torch.flatten(self.pool(self.features(x)), 1))

    return Net()


def gradcam():
    import contextlib
    import io
    import warnings
    warnings.filterwarnings("ignore")
    import metric.grad_cam as gc
    torch.set_grad_enabled(True)
    gen = torch.Generator().manual_seed(4040)
    torch.manual_seed(71)
    net = tiny_vgg()
    imgs = torch.randn(3, 3, 40, 24, generator=gen)
    fx = {"net_state": clone_sd(net), "imgs": imgs}
    with contextlib.redirect_stdout(io.StringIO()):
        for name, cls in (("pp", gc.GradCamPlusPlus), ("base", gc.GradCAM)):
            cam = cls(net, "features.3")
            x = imgs.clone().requires_grad_(True)
            out = cam(x, None)
            fx[name] = {"out": out.clone(), "feature": cam.feature.detach().clone(),
                        "gradient": cam.gradient.detach().clone()}
            cam.remove_handlers()
        logits = net(imgs).detach()
    fx["logits"] = logits
    fx["index"] = torch.tensor(np.argmax(logits.numpy(), axis=1))
    fx["index_max"] = int(np.argmax(np.bincount(np.argmax(logits.numpy(), axis=1))))
    mask = fx["pp"]["out"]
    heat, camimg = gc.mask2cam(mask, imgs)
    fx["mask2cam"] = {"heat": heat, "cam": camimg}
    torch.save(fx, os.path.join(HERE, "gradcam_tiny.pt"))
    print("gradcam_tiny.pt: out", tuple(mask.shape), mask.dtype, "index", fx["index"].tolist(), fx["index_max"])


def encoder_grads():
    """Training-step fixture: gradients of a fixed scalar loss w.r.t. every parameter of the reference encoder
    (loss.backward() as in E_align_s2.py:205), on the weights / image of be_s16_l4.pt with the noise seed 99."""
    import model.E.E as E
    fx = torch.load(os.path.join(HERE, "be_s16_l4.pt"))
    Enc = E.BE(**fx["config"])
    Enc.load_state_dict(fx["state_dict"], strict=True)
    torch.set_grad_enabled(True)
    g = torch.Generator().manual_seed(1234)
    t_const = torch.randn(fx["const"].shape, generator=g)
    t_w = torch.randn(fx["w"].shape, generator=g)
    torch.manual_seed(fx["noise_seed"])
    const, w = Enc(fx["img"])
    loss = ((const - t_const) ** 2).mean() + ((w - t_w) ** 2).mean()     # MSE terms of space_loss (training_utils.py:82-88)
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in Enc.named_parameters() if p.grad is not None}
    unused = [k for k, p in Enc.named_parameters() if p.grad is None]    # parameters the forward never touches
    torch.save({"t_const": t_const, "t_w": t_w, "loss": loss.detach(), "grads": grads, "unused": unused},
               os.path.join(HERE, "be_s16_l4_grads.pt"))
    print("be_s16_l4_grads.pt: loss", float(loss), "params", len(grads), "unused", unused,
          "max|g|", max(float(v.abs().max()) for v in grads.values()))


def _summ(g):
    """Compact pin of a gradient tensor: full values when small, else (l2 norm, sum, first 16 values)."""
    g = g.detach().clone()
    if g.numel() <= 4096:
        return {"full": g}
    return {"l2": g.norm().double().item(), "sum": g.double().sum().item(), "head": g.flatten()[:16].clone(),
            "shape": tuple(g.shape)}


def training_grads():
    """Training-path fixtures for the other families: gradients `loss.backward()` produces through the UNMODIFIED
    reference modules, on the weights / inputs of the existing fixtures (sg2_res32, sg1_l6, biggan_small, e_blur_s16_l6,
    e_big_s16_l4).  Latent gradients are stored in full, parameter gradients as compact pins (`_summ`)."""
    import model.stylegan2_generator as sg2
    import model.stylegan1.net as sg1
    import model.biggan_generator as bg
    from model.utils.biggan_config import BigGANConfig
    import model.E.E_Blur as EB
    import model.E.E_BIG as EBIG
    torch.set_grad_enabled(True)
    out = {}
    tgt = lambda shape, seed: torch.randn(shape, generator=torch.Generator().manual_seed(seed))

    fx = torch.load(os.path.join(HERE, "sg2_res32.pt"))
    G = sg2.StyleGAN2Generator(**fx["config"]).eval()
    G.load_state_dict(fx["state_dict"], strict=True)
    wp = fx["wp"].clone().requires_grad_(True)
    img = G.synthesis(wp)["image"]
    ((img - tgt(img.shape, 1)) ** 2).mean().backward()
    out["sg2_dwp"] = wp.grad.clone()

    fx = torch.load(os.path.join(HERE, "sg1_l6.pt"))
    Gs = sg1.Generator(**fx["config"]).eval()
    Gs.load_state_dict(fx["state_dict"], strict=True)
    out["sg1_dstyles"] = {}
    for lod in fx["images"]:
        st = fx["styles"].clone().requires_grad_(True)
        torch.manual_seed(60 + lod)
        img = Gs.forward(st, lod)
        ((img - tgt(img.shape, 2)) ** 2).mean().backward()
        out["sg1_dstyles"][lod] = st.grad.clone()

    fx = torch.load(os.path.join(HERE, "biggan_small.pt"))
    Gb = bg.BigGAN(BigGANConfig.from_dict(fx["config"])).eval()
    Gb.load_state_dict(fx["state_dict"], strict=True)
    out["biggan_dz"] = {}
    for trunc in fx["images"]:
        z = fx["z"].clone().requires_grad_(True)
        img, _ = Gb(z, fx["label"], trunc)
        ((img - tgt(img.shape, 4)) ** 2).mean().backward()
        out["biggan_dz"][trunc] = z.grad.clone()

    fx = torch.load(os.path.join(HERE, "e_blur_s16_l6.pt"))
    E = EB.BE(**fx["config"])
    E.load_state_dict(fx["state_dict"], strict=True)
    torch.manual_seed(fx["noise_seed"])
    const, w = E(fx["img"])
    (const.sum() + (w ** 2).mean()).backward()
    out["e_blur"] = {k: _summ(p.grad) for k, p in E.named_parameters() if p.grad is not None}

    fx = torch.load(os.path.join(HERE, "e_big_s16_l4.pt"))
    E = EBIG.BE(**fx["config"]).eval()
    E.load_state_dict(fx["state_dict"], strict=True)
    x = E.FromRGB(fx["img"])
    torch.manual_seed(13)
    for blk in E.decode_block:
        x = blk(x, fx["cond"], truncation=0.4)
        x = x[0] if isinstance(x, (tuple, list)) else x
    (x ** 2).mean().backward()
    out["e_big"] = {k: _summ(p.grad) for k, p in E.named_parameters() if p.grad is not None}
    # the whole iteration of E_align_s2.py:152-207 on the small E/G pair of e2g_res32.pt: imgs1 -> E -> G.synthesis ->
    # image-space + latent-space space_loss (stand-in LPIPS, the package is absent) -> backward into E
    import training_utils as tu
    fx = torch.load(os.path.join(HERE, "e2g_res32.pt"))
    G = sg2.StyleGAN2Generator(**fx["g_config"]).eval()
    G.load_state_dict(fx["g_state_dict"], strict=True)
    import model.E.E as E1
    E = E1.BE(**fx["e_config"])
    E.load_state_dict(fx["e_state_dict"], strict=True)

    def lp(a, b):
        return ((a - b) ** 2).mean(dim=(1, 2, 3), keepdim=True) + 0.1 * (a - b).abs().mean(dim=(1, 2, 3), keepdim=True)

    imgs1, w1 = fx["imgs1"], fx["wp1"]
    torch.manual_seed(fx["noise_seed"])
    const2, w2 = E(imgs1)
    imgs2 = G.synthesis(w2)["image"]
    l_img, info_img = tu.space_loss(imgs1, imgs2, lpips_model=lp)
    l_w, info_w = tu.space_loss(w1, w2, image_space=False)
    (l_img + 0.01 * l_w).backward()
    out["e2g_iteration"] = {"l_img": float(l_img), "l_w": float(l_w), "info_img": info_img, "info_w": info_w,
                            "grads": {k: _summ(p.grad) for k, p in E.named_parameters() if p.grad is not None}}
    torch.save(out, os.path.join(HERE, "train_grads.pt"))
    print("train_grads.pt:", {k: (len(v) if isinstance(v, dict) else tuple(v.shape)) for k, v in out.items()})


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "training_grads":
        import_reference()
        training_grads()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "encoder_grads":
        import_reference()
        encoder_grads()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "gradcam":
        import numpy as np
        globals()["np"] = np
        import_reference()
        gradcam()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "biggan":
        import_reference()
        biggan()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "eblur":
        import_reference()
        e_blur()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sg1":
        import_reference()
        stylegan1()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pggan":
        import_reference()
        pggan()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "losses":
        import_reference()
        losses_and_optimizer()
        sys.exit(0)
    main()
