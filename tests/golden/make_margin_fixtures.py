"""Gradient fixtures WITH A MARGIN AROUND ZERO, from the UNMODIFIED reference (run in the dev container only).

leaky_relu's derivative jumps at 0: an implementation whose conv outputs differ from the reference's by rounding puts a
unit with |pre-activation| below that rounding on the other slope, and the gradients move by ~1e-3..1e-2 of their scale
(tests/test_train_gpu.py header).  The fixtures written here are small enough, and their seed is SEARCHED, so that the
smallest |pre-activation| of every leaky_relu input is far above the kernels' 2^-17 operand rounding: gradients of the
CUDA path can then be held to the reference's own numbers at the 1e-3 bar with nothing replayed.
    python tests/golden/make_margin_fixtures.py      ->  tests/golden/e_blur_margin.pt"""
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import clone_sd, import_reference, perturb  # noqa: E402


def margin_run(fn):
    """Run fn() with F.leaky_relu observed -> (result, min over calls of min|x| / std(x))."""
    orig = F.leaky_relu
    worst = [float("inf")]

    def lrelu(x, negative_slope=0.01, inplace=False):
        worst[0] = min(worst[0], float(x.detach().abs().min() / x.detach().std()))
        return orig(x, negative_slope)

    F.leaky_relu = lrelu
    try:
        return fn(), worst[0]
    finally:
        F.leaky_relu = orig


def e_blur_margin():
    import model.E.E_Blur as EB
    cfg = dict(startf=16, maxf=32, layer_count=4, latent_size=512, channels=3)
    best = None
    for seed in range(40):
        gen = torch.Generator().manual_seed(9000 + seed)
        torch.manual_seed(100 + seed)
        E = EB.BE(**cfg).eval()
        perturb(E, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
        img = torch.randn(2, 3, 32, 32, generator=gen)

        def run():
            torch.manual_seed(500 + seed)
            return E(img)
        with torch.no_grad():
            _, m = margin_run(run)
        if best is None or m > best[0]:
            best = (m, seed)
    m, seed = best
    gen = torch.Generator().manual_seed(9000 + seed)
    torch.manual_seed(100 + seed)
    E = EB.BE(**cfg).eval()
    perturb(E, ["noise_weight_1", "noise_weight_2", "bias_1", "bias_2", "bias"], gen)
    img = torch.randn(2, 3, 32, 32, generator=gen)
    torch.manual_seed(500 + seed)
    const, w = E(img)
    (const.sum() + (w ** 2).mean()).backward()
    fx = {"config": cfg, "state_dict": clone_sd(E), "img": img, "noise_seed": 500 + seed, "const": const.detach().clone(),
          "w": w.detach().clone(), "grads": {k: p.grad.clone() for k, p in E.named_parameters() if p.grad is not None},
          "min_preactivation_over_std": m, "fused": [b.fused_scale for b in E.decode_block]}
    torch.save(fx, os.path.join(HERE, "e_blur_margin.pt"))
    print(f"e_blur_margin.pt: seed {seed}, min |pre-activation| / std = {m:.2e}, {len(fx['grads'])} gradients, fused {fx['fused']}")


if __name__ == "__main__":
    import_reference()
    e_blur_margin()
