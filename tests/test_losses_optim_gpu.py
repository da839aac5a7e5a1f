"""GPU parity of `space_loss` / SSIM (training_utils.py:54-99, metric/pytorch_ssim.py) and the fused LREQAdam step
(model/utils/custom_adam.py) against the golden values produced by the unmodified reference and the oracle.
Tolerance: 2e-5 relative on every scalar (fp32 reductions in a different order; moments are accumulated in fp64)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _lpips_stand_in(a, b):
    return ((a - b) ** 2).mean(dim=(1, 2, 3))


def _close(x, y, tol=2e-5):
    return abs(x - y) <= tol * max(1.0, abs(y))


def test_space_loss_matches_reference_golden():
    import training_utils as tu
    cases = torch.load(os.path.join(GOLD, "space_loss.pt"))
    for name, c in cases.items():
        g = torch.Generator().manual_seed(c["seed"])
        a = torch.randn(c["shape"], generator=g)
        b = a + 0.3 * torch.randn(c["shape"], generator=g)
        with torch.no_grad():
            loss, info = tu.space_loss(a.cuda(), b.cuda(), image_space=c["image_space"], lpips_model=_lpips_stand_in)
        assert loss.is_cuda and _close(float(loss), c["loss"]), (name, float(loss), c["loss"])
        flat = lambda i: list(i[0]) + list(i[1:])
        for k, (x, y) in enumerate(zip(flat(info), flat(c["info"]))):
            assert _close(x, y), (name, k, x, y)


def test_ssim_module_and_function_vs_oracle():
    import metric.pytorch_ssim as pssim
    from oracle import losses as olosses
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, 3, 70, 53, generator=g)
    b = (a + 0.1 * torch.randn(a.shape, generator=g)).clamp(0, 1)
    ref = float(olosses.ssim(a, b))
    with torch.no_grad():
        assert _close(float(pssim.ssim(a.cuda(), b.cuda())), ref)
        assert _close(float(pssim.SSIM()(a.cuda(), b.cuda())), ref)
        assert _close(float(pssim.ssim(a.cuda(), a.cuda())), 1.0)     # comparing-baseline.py:88 identity sanity


def test_space_loss_identity_sanity():
    """comparing-baseline.py:88: identical inputs => MSE 0, cosine loss ~0, SSIM 1."""
    import training_utils as tu
    a = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        loss, info = tu.space_loss(a, a.clone(), image_space=True, lpips_model=_lpips_stand_in)
    assert info[0][0] == 0.0 and abs(info[2]) < 1e-6 and abs(info[3]) < 1e-6 and abs(float(loss)) < 1e-5


def test_lreq_adam_matches_reference_golden():
    from model.utils.custom_adam import LREQAdam
    fx = torch.load(os.path.join(GOLD, "lreq_adam.pt"))
    params = [torch.nn.Parameter(p.clone().cuda()) for p in fx["init"]]
    for p, c in zip(params, fx["coefs"]):
        if c is not None:
            p.lr_equalization_coef = c
    opt = LREQAdam(params, lr=fx["lr"], betas=(0.0, fx["beta2"]))
    for gs in fx["grads"]:
        for p, g in zip(params, gs):
            p.grad = None if g is None else g.clone().cuda()
        opt.step()
    rel = lambda a, b: ((a.detach().cpu() - b).abs().max() / b.abs().max()).item()
    assert [opt.state[p]["step"] for p in params] == fx["steps"]
    for p, q in zip(params, fx["final"]):
        assert rel(p, q) < 1e-6
    for p, v in zip(params, fx["exp_avg_sq"]):
        assert rel(opt.state[p]["exp_avg_sq"], v) < 1e-6
    with pytest.raises(ValueError):
        LREQAdam(params, betas=(0.9, 0.99))


def test_lreq_adam_on_encoder_parameters():
    """All 101 parameter tensors of BE(16,9)-like encoders go through ONE kernel launch; compare with the oracle."""
    from dge_b200 import ops
    from model.E.E import BE
    from model.utils.custom_adam import LREQAdam
    from oracle import optim as ooptim
    torch.manual_seed(0)
    E = BE(16, 64, 4, 512, 3).cuda()
    params = list(E.parameters())
    g = torch.Generator().manual_seed(5)
    ref_p = [p.detach().cpu().clone() for p in params]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    coefs = [getattr(p, "lr_equalization_coef", None) for p in params]
    opt = LREQAdam(E.parameters(), lr=0.0015, betas=(0.0, 0.99))
    for step in range(1, 3):
        grads = [torch.randn(p.shape, generator=g) for p in ref_p]
        for p, gr in zip(params, grads):
            p.grad = gr.cuda()
        ops.launch_count_reset()
        opt.step()
        assert ops.launch_count() == 1
        ooptim.lreq_adam_step(ref_p, grads, ref_v, [step] * len(ref_p), coefs, 0.0015, 0.99)
    for p, q in zip(params, ref_p):
        assert ((p.detach().cpu() - q).abs().max() / q.abs().max().clamp_min(1e-20)).item() < 1e-6


def test_space_loss_of_detached_images_still_reaches_the_lpips_module():
    """E_mis_align_cropping_s1.py:171-193 calls `backward()` on losses whose images are all `.detach().clone()`: upstream
    that works only because the LPIPS module's `lin` weights require grad.  The drop-in keeps that graph: the loss has a
    grad_fn, backward() fills the `lin` gradients, and the value is the one the no-grad evaluation returns."""
    import os
    os.environ["DGE_LPIPS_ALLOW_RANDOM"] = "1"
    import lpips
    import training_utils as tu
    lp = lpips.LPIPS(net="vgg", pretrained=False, pnet_rand=True, verbose=False).cuda()
    g = torch.Generator().manual_seed(3)
    a = torch.rand(2, 3, 64, 64, generator=g).cuda() * 2 - 1
    b = torch.rand(2, 3, 64, 64, generator=g).cuda() * 2 - 1
    loss, info = tu.space_loss(a.detach().clone(), b.detach().clone(), lpips_model=lp)
    assert loss.requires_grad
    loss.backward()
    grads = [p.grad for p in lp.lins.parameters()]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads) and any(gr.abs().max() > 0 for gr in grads)
    with torch.no_grad():
        loss0, info0 = tu.space_loss(a, b, lpips_model=lp)
    assert not loss0.requires_grad
    assert abs(float(loss) - float(loss0)) <= 1e-5 * abs(float(loss0)) and abs(info[4] - info0[4]) <= 1e-5 * abs(info0[4]) + 1e-9
    # the `lin` gradients of the fused node (dge_b200/train_lpips.py::_LpipsLinOnlyFn) = those of the graph of torch nodes
    fused = [gr.clone() for gr in grads]
    for p in lp.lins.parameters():
        p.grad = None
    lpips.FUSED = False
    try:
        loss_u, _ = tu.space_loss(a.detach().clone(), b.detach().clone(), lpips_model=lp)
        loss_u.backward()
    finally:
        lpips.FUSED = True
    for gf, p in zip(fused, lp.lins.parameters()):
        # (mean of squared DIFFERENCES of unit-normalised features of two random-VGG maps: the two evaluation orders differ
        #  by ~2e-3 of these 1e-9-sized values)
        assert ((gf - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-20)).item() < 1e-2
