"""BASELINE.json configs[4]: the real-image inversion loop of `embedding_img.py:74-128` with StyleGAN2-1024 synthesis in
place of the StyleGAN1 decoder (the upstream StyleGAN2 variant of the script has unresolvable imports), N = 1:

    per image g:   reload E's weights, fresh LREQAdam state                                     (:82-83)
      per iteration: const2, w1 = E(imgs1); imgs2 = G.synthesis(w1)['image']; const3, w2 = E(imgs2)   (:86-88)
                     3 x space_loss (full image; AT1 / AT2 crops, detached as upstream :96-106)
                     loss_msiv = l_img + (l_medium + l_small) * 0.125 -> backward(retain_graph) -> step   (:109-112)
                     loss_msLv = (space_loss(w1, w2) + space_loss(const2, const3)) * 0.01 -> backward -> step   (:117-128)

Images are synthetic: imgs1 = G(z_i) for seeds 30000 + i (the repo's validation-seed convention, synthesized_IMG.py:97-98).
Reports seconds per image for the upstream 1500 iterations (measured ms / iteration x 1500), the measured wall time of the
run, and the final reconstruction MSE.  `--impl reference` runs the UNMODIFIED reference modules (baseline/_ref) through the
same loop on the same GPU with the same seeds (encoder noise is drawn from the CPU generator in both, so the two runs see
identical noise); bench.py launches both and reports them side by side (`inversion`).
usage: python tools/bench_invert.py [--impl ours|reference] [--images 4] [--iterations 10] [--res 1024]"""
import argparse
import json
import math
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DGE_REF", os.path.join(ROOT, "baseline", "_ref"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--images", type=int, default=4)
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--res", type=int, default=1024)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--encoder", default="e", choices=["e", "blur"],
                    help="e: model/E/E.py (case 1, what E_align_s2.py trains); blur: model/E/E_Blur.py (case 2, the class "
                         "embedding_img.py:9 imports)")
    ap.add_argument("--noise", default="device", choices=["device", "reference"],
                    help="ours: where the encoder's per-block noise is drawn.  'reference' = on the CPU generator then copied, "
                         "as upstream (E.py:60,73): the same stream as the reference arm, used for the MSE comparison; "
                         "'device' = on the GPU (no 11 MB of CPU randn + H2D per encoder pass), used for the timing")
    ap.add_argument("--graphs", action="store_true",
                    help="ours: CUDA-graph replay of the frozen nodes: synthesis pass + LPIPS (dge_b200.graphs.GRAPHS, the DGE_TRAIN_GRAPHS=1 switch)")
    a = ap.parse_args()
    sys.path.insert(0, ROOT)
    if a.impl == "reference":
        if not os.path.isdir(os.path.join(REF, "model")):
            print(json.dumps({"impl": "reference", "unavailable": f"no copy of the reference under {REF}"}))
            return
        for n in ["matplotlib", "matplotlib.pyplot", "boto3", "botocore", "botocore.exceptions", "lpips", "tensorboardX"]:
            sys.modules.setdefault(n, types.ModuleType(n))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["botocore.exceptions"].ClientError = Exception
        sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
        sys.path.insert(0, REF)
    else:
        sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
    import torch
    if a.encoder == "blur":
        import model.E.E_Blur as EM
    else:
        import model.E.E as EM
    import model.stylegan2_generator as SG
    import training_utils as tu
    from model.utils.custom_adam import LREQAdam
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda")
    layers = int(math.log2(a.res)) - 1
    startf = max(16, 512 >> (layers - 1))
    torch.manual_seed(0)
    G = SG.StyleGAN2Generator(a.res).eval()
    E = EM.BE(startf, 512, layers, 512, 3)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for m in (G, E):
            for k, p in m.named_parameters():
                if k.endswith(("bias", "noise_strength", "noise_weight_1", "noise_weight_2", "bias_1", "bias_2")):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    G, E = G.to(dev), E.to(dev)
    if a.impl == "ours":
        E.set_noise_mode(a.noise)
        if a.graphs:
            from dge_b200 import graphs
            graphs.GRAPHS = True
    e_state = {k: v.detach().clone() for k, v in E.state_dict().items()}
    if a.impl == "ours":
        import lpips
        lp = lpips.LPIPS(net="vgg", pretrained=False, pnet_rand=True, verbose=False).to(dev)
        lp_sd = None
    else:
        lp = None
    # the same LPIPS weights in both arms: seeded random VGG16 + non-negative linear layers (published weights unavailable)
    from torchvision.models import vgg16
    torch.manual_seed(2)
    vgg = vgg16(weights=None).features
    sd = {f"net.slice{k + 1}.{i}.{s}": getattr(vgg[i], s).detach().clone()
          for k, idxs in enumerate(([0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28])) for i in idxs
          for s in ("weight", "bias")}
    for k, c in enumerate((64, 128, 256, 512, 512)):
        sd[f"lin{k}.model.1.weight"] = torch.rand(1, c, 1, 1)
    if a.impl == "ours":
        missing = lp.load_state_dict(sd, strict=False).missing_keys
        assert all(k.startswith(("scaling_layer", "lins.")) for k in missing), missing
        lp = lp.to(dev)
    else:
        from oracle import lpips as olp
        sd = {k: v.to(dev) for k, v in sd.items()}
        lp = lambda x, y: olp.lpips_vgg(sd, x, y)

    def crop(x, m):
        return x[:, :, m:-m, m:-m]

    def one_image(i, iters):
        torch.manual_seed(30000 + i)
        z = torch.randn(1, 512)
        with torch.no_grad():
            imgs1 = G(z.to(dev), trunc_psi=0.7, trunc_layers=8, randomize_noise=False)["image"]
        E.load_state_dict(e_state)                                                     # :82
        opt = LREQAdam([{"params": E.parameters()}], lr=a.lr, betas=(0.0, 0.99), weight_decay=0)   # :83 (fresh state)
        mse = None
        for _ in range(iters):
            const2, w1 = E(imgs1)
            imgs2 = G.synthesis(w1)["image"]
            const3, w2 = E(imgs2)
            l_img, info = tu.space_loss(imgs1, imgs2, lpips_model=lp)
            m = imgs1.shape[3] // 8
            l_med, _ = tu.space_loss(imgs1[:, :, :, m:-m].detach().clone(), imgs2[:, :, :, m:-m].detach().clone(),
                                     lpips_model=lp)
            m2 = m + imgs1.shape[2] // 32
            l_small, _ = tu.space_loss(crop(imgs1, m2).detach().clone(), crop(imgs2, m2).detach().clone(), lpips_model=lp)
            opt.zero_grad()
            (l_img + (l_med + l_small) * 0.125).backward(retain_graph=True)
            opt.step()
            l_w, _ = tu.space_loss(w1, w2, image_space=False)
            l_c1, _ = tu.space_loss(const2, const3, image_space=False)
            opt.zero_grad()
            ((l_w + l_c1) * 0.01).backward()
            opt.step()
            mse = info[0][0]
        return mse

    one_image(0, 5 if a.graphs else 2)                    # warm-up (allocator, caches, cuDNN heuristics, graph capture)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    prof = None
    if os.environ.get("DGE_HOSTPROF"):          # where the HOST time of the loop goes (cProfile, top entries to stderr)
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    kprof = None
    if os.environ.get("DGE_GPUTIME"):           # how much of the wall time the GPU is busy (sum of kernel times, to stderr)
        kprof = torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA])
        kprof.__enter__()
    t0 = time.perf_counter()
    mses = [one_image(i, a.iterations) for i in range(a.images)]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if kprof is not None:
        kprof.__exit__(None, None, None)
        busy = sum(e.self_device_time_total for e in kprof.key_averages()) / 1e3
        print(f"GPU busy {busy / (a.images * a.iterations):.2f} ms / iteration of {wall / (a.images * a.iterations) * 1e3:.2f} ms wall "
              f"(under the profiler)", file=sys.stderr)
        print(kprof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=70),
              file=sys.stderr)
    if prof is not None:
        import io
        import pstats
        prof.disable()
        buf = io.StringIO()
        pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(40)
        print(buf.getvalue(), file=sys.stderr)
    ms_it = wall / (a.images * a.iterations) * 1e3
    out = {"impl": a.impl, "encoder": "E_Blur.BE" if a.encoder == "blur" else "E.BE", "encoder_noise": a.noise if a.impl == "ours" else "reference", "graphs": bool(a.graphs and a.impl == "ours"), "workload": f"configs[4]: embedding_img.py:74-128 loop, StyleGAN2-{a.res} synthesis + "
                                       f"BE({startf},{layers}), batch 1, synthetic images G(z_i), seeds 30000+i",
           "images": a.images, "iterations_per_image": a.iterations, "ms_per_iteration": ms_it,
           "s_per_image_measured": wall / a.images, "s_per_image_at_1500_iterations": ms_it * 1.5,
           "recon_mse_after_last_iteration": mses, "recon_mse_mean": sum(mses) / len(mses),
           "peak_gib": torch.cuda.max_memory_allocated() / 2 ** 30, "timing": "wall clock around the whole loop with a "
           "device synchronize on both sides (the loop reads losses back to the host every iteration, as upstream)"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
