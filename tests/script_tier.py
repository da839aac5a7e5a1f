"""Script tier: run an UNMODIFIED entry script of the reference (`E_align_s2.py`, `embedding_img.py`) against the drop-in
package.  The script's source is read from a copy of the reference ($DGE_REF, baseline/_ref, /root/reference -- never
edited, never part of this repository); only its environment is prepared:
  * the drop-in package directory comes first on sys.path, so `import model.E.E`, `training_utils`, `lpips`, ... resolve to
    this repository; `tensorboardX` (absent here) is a null writer;
  * the checkpoints the script loads (`./checkpoint/...`, not shipped, no network) are synthetic: seeded random-init weights
    saved in the file formats the script expects;
  * the globals the script defines under `if __name__ == "__main__"` and reads inside `train()` (`device`, `resultPath*`,
    `writer_path`) are set on the loaded module.
Runs in its own process (tests/test_scripts_*.py).  `--cpu-plumbing`: no GPU -- `.cuda()` / `.to('cuda')` become no-ops and the
run must reach the first kernel call and stop there with DgeError (there is no CPU fallback).
Prints one JSON line."""
import argparse
import importlib.util
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deep-gan-encoders_b200")


def find_script(name):
    for d in (os.environ.get("DGE_REF"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if d and os.path.isfile(os.path.join(d, name)):
            return os.path.join(d, name)
    return None


class NullWriter:
    def __getattr__(self, _):
        return lambda *a, **k: None


def load_script(path, modname):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # defines train(); the argparse / main block is guarded by __name__
    return mod


def perturb(module, seed):
    import torch
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in list(module.named_parameters()):
            if k.endswith(("bias", "noise_strength", "noise_weight_1", "noise_weight_2", "bias_1", "bias_2")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script", choices=["E_align_s2.py", "embedding_img.py", "E_mis_align_cropping_s1.py", "E_align_cropping_s1.py"])
    ap.add_argument("--cpu-plumbing", action="store_true")
    ap.add_argument("--img-size", type=int, default=64)
    ap.add_argument("--iterations", type=int, default=2)
    ap.add_argument("--mtype", type=int, default=2, choices=[1, 2, 4],
                    help="E_align_s2.py: 1 = StyleGAN1 (mapping network left on the CPU as upstream), 2 = StyleGAN2, 4 = BigGAN")
    a = ap.parse_args()
    path = find_script(a.script)
    if path is None:
        print(json.dumps({"skipped": f"no copy of the reference's {a.script} ($DGE_REF, baseline/_ref, /root/reference)"}))
        return
    os.environ["DGE_LPIPS_ALLOW_RANDOM"] = "1"       # no pretrained LPIPS weights offline (explicit opt-in, see lpips/__init__.py)
    sys.path.insert(0, PKG)
    tb = types.ModuleType("tensorboardX")
    tb.SummaryWriter = lambda *x, **k: NullWriter()
    sys.modules["tensorboardX"] = tb
    import torch
    if a.cpu_plumbing:
        torch.Tensor.cuda = lambda self, *x, **k: self
        torch.nn.Module.cuda = lambda self, *x, **k: self
        _to = torch.nn.Module.to
        torch.nn.Module.to = lambda self, *x, **k: self if (x and str(x[0]).startswith("cuda")) else _to(self, *x, **k)
        _tto = torch.Tensor.to
        torch.Tensor.to = lambda self, *x, **k: self if (x and isinstance(x[0], (str, torch.device)) and
                                                         str(x[0]).startswith("cuda")) else _tto(self, *x, **k)
    from dge_b200 import ops
    from dge_b200._lib import DgeError
    work = tempfile.mkdtemp(prefix="dge_script_")
    os.chdir(work)
    res = os.path.join(work, "result")
    for d in (res, res + "/imgs", res + "/models", res + "/summaries", res + "/grad_cam"):
        os.makedirs(d)
    layers = int(__import__("math").log2(a.img_size)) - 1
    # start_features must reach maxf = 512 by the last block (the reference's last block blends `inputs`-channel features
    # with its `outputs`-channel residual: E.py:84 only works when they are equal): 16 -> 1024, 32 -> 512, 64 -> 256 upstream
    startf = max(16, 512 >> (layers - 1))
    torch.manual_seed(0)
    mod = load_script(path, "ref_script_" + a.script[:-3])
    assert os.path.abspath(sys.modules["model.E.E"].__file__).startswith(PKG), "the script must import the drop-in package"
    mod.device = torch.device("cpu" if a.cpu_plumbing else "cuda")
    mod.resultPath, mod.resultPath1_1, mod.resultPath1_2 = res, res + "/imgs", res + "/models"
    mod.writer_path = res + "/summaries"
    mod.resultPath_grad_cam = res + "/grad_cam"
    out = {"script": path, "iterations": a.iterations, "img_size": a.img_size}
    if a.script == "E_mis_align_cropping_s1.py":
        # the Grad-CAM variant (:99-106, 158-171): torchvision's VGG16 classifier (its published weights are not available
        # offline: `pretrained=True` is answered with seeded random weights), Grad-CAM++ masks, guided back-propagation and
        # mask2cam on both images of every iteration, StyleGAN2 generator
        import torchvision
        _vgg16 = torchvision.models.vgg16

        def vgg16_offline(*_a, **_k):
            torch.manual_seed(7)
            return _vgg16(weights=None)
        torchvision.models.vgg16 = vgg16_offline
        from model.stylegan2_generator import StyleGAN2Generator
        G = StyleGAN2Generator(resolution=a.img_size)
        perturb(G, 1)
        ck = os.path.join(work, "stylegan2_synth.pth")
        torch.save({"generator_smooth": G.state_dict()}, ck)
        args = argparse.Namespace(mtype=2, checkpoint_dir_GAN=ck, config_dir=None, checkpoint_dir_E=None,
                                  img_size=a.img_size, img_channels=3, z_dim=512, start_features=startf, batch_size=2,
                                  iterations=a.iterations, lr=0.0015, beta_1=0.0, experiment_dir=res)
        call = lambda: mod.train(tensor_writer=NullWriter(), args=args)
        expect = [res + "/models/E_model_ep0_iter0.pth", res + "/Loss.txt", res + "/imgs/ep0_iter0.png",
                  res + "/grad_cam/heatmap_0.png", res + "/grad_cam/cam_0.png", res + "/grad_cam/gb_0.png"]
    elif a.script in ("E_align_s2.py", "E_align_cropping_s1.py"):       # (the cropping variant: same loop, detached losses)
        cfg_path, z_dim = None, 512
        if a.mtype == 2:
            from model.stylegan2_generator import StyleGAN2Generator
            G = StyleGAN2Generator(resolution=a.img_size)
            perturb(G, 1)
            ck = os.path.join(work, "stylegan2_synth.pth")
            torch.save({"generator_smooth": G.state_dict()}, ck)
        elif a.mtype == 1:
            from model.stylegan1.net import Generator, Mapping
            ck = os.path.join(work, "sg1") + "/"
            os.makedirs(ck)
            Gs = Generator(startf=startf, maxf=512, layer_count=layers, latent_size=512, channels=3)
            Gm = Mapping(num_layers=2 * layers, mapping_layers=8, latent_size=512, dlatent_size=512, mapping_fmaps=512)
            perturb(Gs, 2)
            torch.save(Gs.state_dict(), ck + "Gs_dict.pth")
            torch.save(Gm.state_dict(), ck + "Gm_dict.pth")
            torch.save(torch.zeros(2 * layers, 512), ck + "center_tensor.pt")
        else:
            import json as _json
            from model.biggan_generator import BigGAN
            from model.utils.biggan_config import BigGANConfig
            a.img_size, layers, startf, z_dim = 128, 6, 16, 128       # a 128-px BigGAN-deep: 5 up blocks from 4x4
            cfg = {"attention_layer_position": 8, "channel_width": 64, "class_embed_dim": 128, "eps": 0.0001,
                   "layers": [[False, 16, 16], [True, 16, 16], [False, 16, 16], [True, 16, 8], [False, 8, 8], [True, 8, 4],
                              [False, 4, 4], [True, 4, 2], [False, 2, 2], [True, 2, 1]],
                   "n_stats": 51, "num_classes": 1000, "output_dim": 128, "z_dim": 128}
            cfg_path = os.path.join(work, "biggan_synth.json")
            with open(cfg_path, "w") as f:
                _json.dump(cfg, f)
            G = BigGAN(BigGANConfig.from_dict(cfg))
            with torch.no_grad():
                G.generator.bn.weight.fill_(1.0)
                G.generator.bn.bias.zero_()
            ck = os.path.join(work, "biggan_synth.pt")
            torch.save(G.state_dict(), ck)
        args = argparse.Namespace(mtype=a.mtype, checkpoint_dir_GAN=ck, config_dir=cfg_path, checkpoint_dir_E=None,
                                  img_size=a.img_size, img_channels=3, z_dim=z_dim, start_features=startf, batch_size=2,
                                  iterations=a.iterations, lr=0.0015, beta_1=0.0, experiment_dir=res)
        call = lambda: mod.train(tensor_writer=NullWriter(), args=args)
        expect = [res + "/models/E_model_ep0_iter0.pth", res + "/Loss.txt", res + "/imgs/ep0_iter0.jpg"]
    else:
        import model.E.E_Blur as EB
        from model.stylegan1.net import Generator, Mapping
        gdir = os.path.join(work, "sg1") + "/"
        os.makedirs(gdir)
        Gs = Generator(startf=startf, maxf=512, layer_count=layers, latent_size=512, channels=3)
        Gm = Mapping(num_layers=2 * layers, mapping_layers=8, latent_size=512, dlatent_size=512, mapping_fmaps=512)
        perturb(Gs, 2)
        torch.save(Gs.state_dict(), gdir + "Gs_dict.pth")
        torch.save(Gm.state_dict(), gdir + "Gm_dict.pth")
        torch.save(torch.zeros(2 * layers, 512), gdir + "center_tensor.pt")
        E = EB.BE(startf=startf, maxf=512, layer_count=layers, latent_size=512, channels=3)
        perturb(E, 3)
        eck = os.path.join(work, "E_blur_synth.pth")
        torch.save(E.state_dict(), eck)
        args = argparse.Namespace(mtype=1, checkpoint_dir_GAN=gdir, config_dir=None, checkpoint_dir_E=eck,
                                  img_size=a.img_size, img_channels=3, z_dim=512, start_features=startf, batch_size=1,
                                  iterations=a.iterations, lr=0.01, beta_1=0.0, experiment_dir=res, optimizeE=True,
                                  img_dir=None)
        imgs = (torch.rand(2, 3, a.img_size, a.img_size, generator=torch.Generator().manual_seed(5)) * 2 - 1).to(mod.device)
        call = lambda: mod.train(tensor_writer=NullWriter(), args=args, imgs_tensor=imgs)
        expect = [res + "/models/w_all_1.pt", res + "/models/img_all_1.pt", res + "/Loss.txt", res + "/summaries/00001_rec.png"]
    try:
        if not a.cpu_plumbing:
            ops.launch_count_reset()
        call()
        out["completed"] = True
    except DgeError as exc:
        out["completed"] = False
        out["dge_error"] = str(exc)[:200]
    if not a.cpu_plumbing:
        out["dge_launches"] = ops.launch_count()
        out["missing_outputs"] = [p for p in expect if not os.path.exists(p)]
        e_ck = res + "/models/E_model_ep0_iter0.pth"
        if os.path.exists(e_ck):
            sd = torch.load(e_ck, map_location="cpu")
            out["e_checkpoint_keys"] = len(sd)
            out["e_checkpoint_finite"] = bool(all(torch.isfinite(v).all() for v in sd.values()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
