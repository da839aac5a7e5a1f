"""The UNMODIFIED reference on the same B200 ("the fair bar"): its own modules through PyTorch / cuDNN, fp32
with TF32 off (its numerics), on the workloads bench.py times:
  * forward:  const2, w2 = E(imgs1); imgs2 = G.synthesis(w2)['image']            (E_align_s2.py:153,160), batch 8 @ 1024^2
  * training: one whole E_align_s2.py iteration (:140-221) -- G(z) under no_grad, E, G.synthesis, three image-space
    `space_loss` calls + the latent one, two backward / LREQAdam.step pairs -- with the reference's own `space_loss`,
    `LREQAdam`, SSIM; LPIPS = the published LPIPS-VGG16 algorithm in plain torch (oracle/lpips.py, random weights: the
    third-party package and its weights are not available offline).
Runs in its own process (the reference's module names `model`, `metric`, `training_utils` clash with the drop-in's); bench.py
launches it as a subprocess on rank 0 and copies the JSON into its line (`reference_gpu`).  Needs the copy of the reference
under baseline/_ref (git-ignored, travels with gpurun) or $DGE_REF.
usage: python tools/ref_gpu_timing.py [--batch 8] [--iters 5] [--train-iters 3] [--no-train] [--tf32]"""
import argparse
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DGE_REF", os.path.join(ROOT, "baseline", "_ref"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--startf", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--train-iters", type=int, default=3)
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--tf32", action="store_true", help="also time the forward with TF32 on / bf16 autocast (context)")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(REF, "model")):
        print(json.dumps({"unavailable": f"no copy of the reference under {REF}"}))
        return
    for n in ["matplotlib", "matplotlib.pyplot", "boto3", "botocore", "botocore.exceptions", "lpips", "tensorboardX"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["botocore.exceptions"].ClientError = Exception
    sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
    sys.path.insert(0, ROOT)          # oracle/ (LPIPS restatement only)
    sys.path.insert(0, REF)
    import math
    import torch
    import model.E.E as E
    import model.stylegan2_generator as sg2
    import training_utils as tu
    from model.utils.custom_adam import LREQAdam
    assert os.path.abspath(sg2.__file__).startswith(os.path.abspath(REF)), "reference modules must come from baseline/_ref"

    torch.manual_seed(0)
    dev = torch.device("cuda")
    layers = int(math.log2(args.res)) - 1
    G = sg2.StyleGAN2Generator(args.res).eval().to(dev)
    Enc = E.BE(args.startf, 512, layers, 512, 3).to(dev)
    out = {"what": "unmodified reference modules (baseline/_ref) on this GPU, PyTorch eager fp32, TF32 off",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "batch": args.batch, "res": args.res}

    def timeit(fn, warm, it):
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

    z = torch.randn(args.batch, 512, device=dev)
    modes = [("fp32_tf32off", False, False)] + ([("fp32_tf32on", True, False), ("bf16_autocast", True, True)] if args.tf32 else [])
    with torch.no_grad():
        imgs1 = G(z, trunc_psi=0.7, trunc_layers=8, randomize_noise=False)["image"]
        for name, tf32, ac in modes:
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if ac else torch.autocast("cuda", enabled=False)
            with ctx:
                def both():
                    c2, w2 = Enc(imgs1)
                    return G.synthesis(w2)["image"]
                tb = timeit(both, 2, args.iters)
            out[name] = {"E_plus_G_ms": tb, "images_per_s": args.batch / (tb / 1e3)}
    out["fwd_ms_per_step"] = out["fp32_tf32off"]["E_plus_G_ms"]
    out["fwd_images_per_s"] = out["fp32_tf32off"]["images_per_s"]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False

    if not args.no_train:
        from oracle import lpips as olp
        from torchvision.models import vgg16
        torch.manual_seed(1)
        vgg = vgg16(weights=None).features.to(dev).eval()
        sd = {f"net.slice{k + 1}.{i}.{s}": getattr(vgg[i], s).detach()
              for k, idxs in enumerate(([0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28])) for i in idxs
              for s in ("weight", "bias")}
        for k, c in enumerate((64, 128, 256, 512, 512)):
            sd[f"lin{k}.model.1.weight"] = torch.rand(1, c, 1, 1, device=dev)
        sd["scaling_layer.shift"] = torch.tensor([-.030, -.088, -.188], device=dev)[None, :, None, None]
        sd["scaling_layer.scale"] = torch.tensor([.458, .448, .450], device=dev)[None, :, None, None]
        lp = lambda a, b: olp.lpips_vgg(sd, a, b)
        opt = LREQAdam([{"params": Enc.parameters()}], lr=0.0015, betas=(0.0, 0.99), weight_decay=0)

        def iteration():
            with torch.no_grad():
                r = G(z, trunc_psi=0.7, trunc_layers=8, randomize_noise=False)
                i1, w1 = r["image"], r["wp"]
            const2, w2 = Enc(i1)
            i2 = G.synthesis(w2)["image"]
            opt.zero_grad()
            l0, _ = tu.space_loss(i1, i2, lpips_model=lp)
            m = i1.shape[3] // 8
            l1, _ = tu.space_loss(i1[:, :, :, m:-m], i2[:, :, :, m:-m], lpips_model=lp)
            m2 = m + i1.shape[2] // 32
            l2, _ = tu.space_loss(i1[:, :, m2:-m2, m2:-m2], i2[:, :, m2:-m2, m2:-m2], lpips_model=lp)
            opt.zero_grad()
            (l0 + l1 * 5 + l2 * 9).backward(retain_graph=True)
            opt.step()
            lw, _ = tu.space_loss(w1, w2, image_space=False)
            opt.zero_grad()
            (lw * 0.01).backward()
            opt.step()

        try:
            torch.cuda.reset_peak_memory_stats()
            tt = timeit(iteration, 1, args.train_iters)
            out["train_iter_ms"] = tt
            out["train_images_per_s"] = args.batch / (tt / 1e3)
            out["train_peak_gib"] = torch.cuda.max_memory_allocated() / 2 ** 30
            out["train_what"] = ("E_align_s2.py:140-221 iteration with the reference's modules, space_loss, SSIM, LREQAdam; "
                                 "LPIPS-VGG16 in plain torch with random weights")
        except Exception as exc:  # noqa: BLE001 -- e.g. out of memory next to the bench process
            out["train_error"] = repr(exc)[:300]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
