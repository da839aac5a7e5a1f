"""Context measurements for the other BASELINE.json configs (not the driver's bench line): E + G forward images/s of
the StyleGAN1-256 (configs[1], batch 16), PGGAN-256 (configs[0] shapes, batch 2 and 16) and BigGAN-deep-256
(configs[3], batch 32) pairs on one B200, random-init weights, CUDA events, eager launches.
  python tools/bench_families.py ours        # this repository's kernels
  python tools/bench_families.py reference   # the UNMODIFIED reference modules (needs baseline/_ref or $DGE_REF)
Prints one JSON object."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = sys.argv[1] if len(sys.argv) > 1 else "ours"
if which == "reference":
    REF = os.environ.get("DGE_REF", os.path.join(ROOT, "baseline", "_ref"))
    for n in ["matplotlib", "matplotlib.pyplot", "boto3", "botocore", "botocore.exceptions", "lpips", "tensorboardX"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    sys.path.insert(0, REF)
else:
    sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")

BIGGAN_CFG = {"attention_layer_position": 8, "channel_width": 128, "class_embed_dim": 128, "eps": 0.0001,
              "layers": [[False, 16, 16], [True, 16, 16], [False, 16, 16], [True, 16, 8], [False, 8, 8], [True, 8, 8],
                         [False, 8, 8], [True, 8, 4], [False, 4, 4], [True, 4, 2], [False, 2, 2], [True, 2, 1]],
              "n_stats": 51, "num_classes": 1000, "output_dim": 256, "z_dim": 128}


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def perturb(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in list(m.named_parameters()):
            if p.abs().max() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


out = {"impl": which, "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
with torch.no_grad():
    # ---- StyleGAN1-256 + BE(64, 7), batch 16 (configs[1]) ------------------------------------------------
    from model.stylegan1.net import Generator, Mapping
    from model.E.E import BE
    torch.manual_seed(0)
    Gs = Generator(64, 512, 7, 512, 3).eval()
    Gm = Mapping(14, 8, 512, 512, 512).eval()
    Gm.buffer1 = torch.zeros(14, 512)
    E1 = BE(64, 512, 7, 512, 3).eval()
    for m in (Gs, E1):
        perturb(m, 1)
    Gs, Gm, E1 = Gs.to(dev), Gm.to(dev), E1.to(dev)
    Gm.buffer1 = Gm.buffer1.to(dev)
    for m in (E1, Gs):          # ours: draw the per-stage noise on the device (the reference draws it on the CPU)
        if hasattr(m, "set_noise_mode"):
            m.set_noise_mode("device")
    coefs = torch.ones(14, 1, device=dev)
    coefs[:7] = 0.7
    z = torch.randn(16, 512, device=dev)
    styles = Gm(z, coefs)
    imgs = Gs.forward(styles, 6)

    def sg1_step():
        c, w = E1(imgs)
        return Gs.forward(w, 6)
    ms = timeit(sg1_step)
    out["stylegan1_256_bs16"] = {"ms_per_step": ms, "images_per_s": 16 / (ms / 1e3)}

    # ---- PGGAN-256 + E_PG(64, 7) --------------------------------------------------------------------------
    from model.pggan.pggan_generator import PGGANGenerator
    from model.E.E_PG import BE as BE_PG
    import contextlib
    import io
    torch.manual_seed(0)
    Gp = PGGANGenerator(256).eval().to(dev)
    Ep = BE_PG(64, 512, 7, 512, 3, pggan=True).eval()
    perturb(Ep, 2)
    Ep = Ep.to(dev)
    if hasattr(Ep, "set_noise_mode"):
        Ep.set_noise_mode("device")
    for bs in (2, 16):
        zp = torch.randn(bs, 512, device=dev)

        def pg_step():
            with contextlib.redirect_stdout(io.StringIO()):      # the reference prints x.shape per block
                img = Gp(zp)["image"]
                return Ep(img)
        ms = timeit(pg_step)
        out[f"pggan_256_bs{bs}"] = {"ms_per_step": ms, "images_per_s": bs / (ms / 1e3)}

    # ---- BigGAN-deep-256 + E_BIG(64, 7), batch 32 (configs[3]) ---------------------------------------------
    from model.biggan_generator import BigGAN
    from model.utils.biggan_config import BigGANConfig
    from model.E.E_BIG import BE as BE_BIG
    torch.manual_seed(0)
    Gb = BigGAN(BigGANConfig.from_dict(BIGGAN_CFG)).eval().to(dev)
    Eb = BE_BIG(64, 512, 7, 512, 3, biggan=True).eval()
    perturb(Eb, 3)
    Eb = Eb.to(dev)
    if hasattr(Eb, "set_noise_mode"):
        Eb.set_noise_mode("device")
    zb = torch.randn(32, 128, device=dev).clamp_(-0.8, 0.8)
    label = torch.zeros(32, 1000, device=dev)
    label[:, 30] = 1

    def big_step():
        img, cond = Gb(zb, label, 0.4)
        return Eb(img, cond)
    ms = timeit(big_step, warm=2, it=5)
    out["biggan_deep_256_bs32"] = {"ms_per_step": ms, "images_per_s": 32 / (ms / 1e3)}
print(json.dumps(out))
