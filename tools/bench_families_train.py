"""Context measurement for BASELINE.json configs[1] and configs[3] (not the driver's bench line): one E_align_s2.py TRAINING
iteration (:140-221) of the StyleGAN1-256 pair (mtype 1: `BE(64, 7)` + `Generator(64, 7)`, batch 16) and of the
BigGAN-deep-256 pair (mtype 4: `E_BIG(64, 7)` + BigGAN, batch 32) on one B200, random-init weights.
  python tools/bench_families_train.py ours | reference      (reference = the UNMODIFIED modules under baseline/_ref)
ours: the case-1 encoder of mtype 1 and all losses run the fused nodes; the StyleGAN1 / BigGAN generators and E_BIG record
graphs of separate torch nodes with their convs on the tensor-core kernels.  Prints one JSON object."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = sys.argv[1] if len(sys.argv) > 1 else "ours"
only = sys.argv[2] if len(sys.argv) > 2 else "both"
sys.path.insert(0, ROOT)
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

import training_utils as tu
from model.utils.custom_adam import LREQAdam

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
# under torchrun (BASELINE configs[3] names 4 x B200): one process per GPU, each with its own batch; LREQAdam then owns the
# bucketed NCCL all-reduce of the encoder gradients (dge_b200/dist.py), nothing else changes in the loop
WORLD, RANK = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if WORLD > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
dev = torch.device("cuda")
BIGGAN_CFG = {"attention_layer_position": 8, "channel_width": 128, "class_embed_dim": 128, "eps": 0.0001,
              "layers": [[False, 16, 16], [True, 16, 16], [False, 16, 16], [True, 16, 8], [False, 8, 8], [True, 8, 8],
                         [False, 8, 8], [True, 8, 4], [False, 4, 4], [True, 4, 2], [False, 2, 2], [True, 2, 1]],
              "n_stats": 51, "num_classes": 1000, "output_dim": 256, "z_dim": 128}


def make_lpips():
    from torchvision.models import vgg16
    torch.manual_seed(2)
    vgg = vgg16(weights=None).features
    sd = {f"net.slice{k + 1}.{i}.{s}": getattr(vgg[i], s).detach().clone()
          for k, idxs in enumerate(([0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28])) for i in idxs
          for s in ("weight", "bias")}
    for k, c in enumerate((64, 128, 256, 512, 512)):
        sd[f"lin{k}.model.1.weight"] = torch.rand(1, c, 1, 1)
    if which == "ours":
        import lpips
        lp = lpips.LPIPS(net="vgg", pretrained=False, pnet_rand=True, verbose=False)
        lp.load_state_dict(sd, strict=False)
        return lp.to(dev)
    from oracle import lpips as olp
    sd = {k: v.to(dev) for k, v in sd.items()}
    return lambda x, y: olp.lpips_vgg(sd, x, y)


def perturb(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in list(m.named_parameters()):
            if p.abs().max() == 0:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def timeit(fn, warm, it):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if WORLD > 1:
        dist.barrier()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    if WORLD > 1:                      # device time, max over ranks
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def losses_and_steps(opt, lp, imgs1, imgs2, w1, w2):
    l0, _ = tu.space_loss(imgs1, imgs2, lpips_model=lp)
    m = imgs1.shape[3] // 8
    l1, _ = tu.space_loss(imgs1[:, :, :, m:-m], imgs2[:, :, :, m:-m], lpips_model=lp)
    m2 = m + imgs1.shape[2] // 32
    l2, _ = tu.space_loss(imgs1[:, :, m2:-m2, m2:-m2], imgs2[:, :, m2:-m2, m2:-m2], lpips_model=lp)
    opt.zero_grad()
    (l0 + l1 * 5 + l2 * 9).backward(retain_graph=True)
    opt.step()
    lw, _ = tu.space_loss(w1, w2, image_space=False)
    opt.zero_grad()
    (lw * 0.01).backward()
    opt.step()


out = {"impl": which, "gpu": torch.cuda.get_device_name(0)}
lp = make_lpips()
if only in ("both", "sg1"):
    from model.E.E import BE
    from model.stylegan1.net import Generator, Mapping
    torch.manual_seed(0)
    Gs = Generator(64, 512, 7, 512, 3)
    Gm = Mapping(14, 8, 512, 512, 512).eval()
    Gm.buffer1 = torch.zeros(14, 512)
    E1 = BE(64, 512, 7, 512, 3)
    for mod in (Gs, E1):
        perturb(mod, 1)
    Gs, Gm, E1 = Gs.to(dev), Gm.to(dev), E1.to(dev)
    Gm.buffer1 = Gm.buffer1.to(dev)
    for mod in (E1, Gs):
        if hasattr(mod, "set_noise_mode"):
            mod.set_noise_mode("device")
    coefs = torch.ones(1, 14, 1, device=dev)
    coefs[:, :7] = 0.7
    opt = LREQAdam([{"params": E1.parameters()}], lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
    z = torch.randn(16, 512, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + RANK))

    def sg1_iter():
        with torch.no_grad():
            w1 = Gm(z, coefs_m=coefs)
            imgs1 = Gs.forward(w1, 6)
        const2, w2 = E1(imgs1)
        imgs2 = Gs.forward(w2, 6)
        losses_and_steps(opt, lp, imgs1, imgs2, w1, w2)
    ms = timeit(sg1_iter, 2, 5)
    out["stylegan1_256_bs16_train"] = {"ms_per_iteration": ms, "images_per_s": 16 * WORLD / (ms / 1e3),
                                       "peak_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
    if only == "both":
        del Gs, Gm, E1, opt
        torch.cuda.empty_cache()
if only in ("both", "biggan"):
    from model.biggan_generator import BigGAN
    from model.E.E_BIG import BE as BE_BIG
    from model.utils.biggan_config import BigGANConfig
    torch.manual_seed(0)
    Gb = BigGAN(BigGANConfig.from_dict(BIGGAN_CFG)).eval()
    with torch.no_grad():
        Gb.generator.bn.weight.fill_(1.0)
        Gb.generator.bn.bias.zero_()
        # random spectral-norm vectors give huge effective weights (and NaN images): converge them as a trained checkpoint's are
        for m in Gb.modules():
            if hasattr(m, "weight_u"):
                m.train()
                for _ in range(20):
                    for hook in m._forward_pre_hooks.values():
                        hook(m, None)
        Gb.eval()
    Gb = Gb.to(dev)
    Eb = BE_BIG(64, 512, 7, 512, 3, biggan=True)
    perturb(Eb, 3)
    Eb = Eb.to(dev)
    if hasattr(Eb, "set_noise_mode"):
        Eb.set_noise_mode("device")
    opt = LREQAdam([{"params": Eb.parameters()}], lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
    bs = 32
    zb = torch.randn(bs, 128, device=dev, generator=torch.Generator(device=dev).manual_seed(200 + RANK)).clamp_(-2, 2) * 0.4
    label = torch.zeros(bs, 1000, device=dev)
    label[:, 30] = 1

    def big_iter():
        with torch.no_grad():
            imgs1, const1 = Gb(zb, label, 0.4)
        const2, w2 = Eb(imgs1, const1)
        imgs2, _ = Gb(w2, label, 0.4)
        losses_and_steps(opt, lp, imgs1, imgs2, zb, w2)
    if os.environ.get("DGE_SYNCDBG") and WORLD > 1:
        def rep(tag):
            nm = [k for k, _ in Eb.named_parameters()]
            for what, ts in (("param", [p.detach() for p in Eb.parameters()]),
                             ("grad", [p.grad.detach() if p.grad is not None else torch.zeros(1, device=dev) for p in Eb.parameters()])):
                t = torch.stack([x.double().sum() for x in ts])
                lo, hi = t.clone(), t.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                bad = [nm[i] for i in torch.nonzero(hi != lo).flatten().tolist()]
                if RANK == 0:
                    print(tag, what, "differ", len(bad), bad[:3], file=sys.stderr, flush=True)
        rep("init")
        for it in range(3):
            big_iter()
            rep(f"after iteration {it}")
    ms = timeit(big_iter, 2, 3)
    if os.environ.get("DGE_KPROF"):             # where the iteration's GPU time goes (kernel table to stderr)
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU,
                                                torch.profiler.ProfilerActivity.CUDA]) as kp:
            big_iter()
            torch.cuda.synchronize()
        print(kp.key_averages().table(sort_by="self_cuda_time_total", row_limit=60, max_name_column_width=90),
              file=sys.stderr)
    out["biggan_deep_256_bs32_train"] = {"ms_per_iteration": ms, "images_per_s": bs * WORLD / (ms / 1e3),
                                         "peak_gib": torch.cuda.max_memory_allocated() / 2 ** 30}
out["n_gpus"] = WORLD
if WORLD > 1:
    # the replicas must hold identical weights after their steps (same initial weights, averaged gradients)
    enc = Eb if only in ("both", "biggan") else E1
    names = [k for k, _ in enc.named_parameters()]
    chk = torch.stack([p.detach().double().sum() for p in enc.parameters()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    differ = [names[i] for i in torch.nonzero(hi != lo).flatten().tolist()]
    out["parameters_finite"] = bool(torch.isfinite(chk).all())
    out["replicas_in_sync"] = not differ
    out["parameters_out_of_sync"] = differ[:12]
    dist.barrier()
if RANK == 0:
    print(json.dumps(out))
if WORLD > 1:
    dist.destroy_process_group()
