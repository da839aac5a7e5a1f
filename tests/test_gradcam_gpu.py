"""GPU parity of the Grad-CAM path (a14, metric/grad_cam.py): device CAM maps / class index / JET overlay against the
golden fixture produced by the unmodified reference classes on a seeded VGG-like stand-in, and an end-to-end run of the
mirrored classes (hooks + backward on the caller's network, post-processing on the device).
Index / argmax / bincount results are bit-exact; float maps within 1e-5 (float64 path) / 1e-4 (float32 path)."""
import contextlib
import io
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.fixture(scope="module")
def fx():
    return torch.load(os.path.join(GOLD, "gradcam_tiny.pt"))


def test_cam_maps_from_hooked_tensors(fx):
    import metric.grad_cam as gc
    pp = gc.cam_maps(fx["pp"]["feature"].cuda(), fx["pp"]["gradient"].cuda(), (40, 24), plus=True)
    assert pp.dtype == torch.float64 and pp.shape == (3, 1, 40, 24)
    assert rel(pp, fx["pp"]["out"]) < 1e-5
    base = gc.cam_maps(fx["base"]["feature"].cuda(), fx["base"]["gradient"].cuda(), (40, 24), plus=False)
    assert rel(base, fx["base"]["out"]) < 1e-4


def test_class_index_bit_exact(fx):
    import metric.grad_cam as gc
    idx, mode = gc._class_index(fx["logits"].cuda(), None)
    assert idx.cpu().tolist() == fx["index"].tolist() and mode == fx["index_max"]
    # ties: first maximum / smallest class wins, as np.argmax / np.bincount do
    logits = torch.tensor([[1., 3., 3., 0.], [5., 5., 1., 0.], [0., 1., 2., 2.], [9., 0., 0., 9.]]).cuda()
    idx, mode = gc._class_index(logits, None)
    assert idx.cpu().tolist() == [1, 0, 2, 0] and mode == 0


def test_mask2cam(fx):
    import metric.grad_cam as gc
    heat, cam = gc.mask2cam(fx["pp"]["out"].cuda(), fx["imgs"].cuda())
    assert rel(heat, fx["mask2cam"]["heat"]) < 1e-6
    assert rel(cam, fx["mask2cam"]["cam"]) < 1e-5


def test_gradcam_classes_end_to_end(fx):
    """Mirrored classes on the caller's network (hooks + autograd), device post-processing."""
    import metric.grad_cam as gc
    from make_golden import tiny_vgg
    net = tiny_vgg()
    net.load_state_dict(fx["net_state"])
    net = net.cuda()
    with contextlib.redirect_stdout(io.StringIO()) as buf:
        for name, cls, tol in (("pp", gc.GradCamPlusPlus, 1e-4), ("base", gc.GradCAM, 1e-3)):
            cam = cls(net, "features.3")
            x = fx["imgs"].clone().cuda().requires_grad_(True)
            out = cam(x, None)
            cam.remove_handlers()
            assert out.dtype == torch.float64 and out.shape == (3, 1, 40, 24)
            assert rel(out, fx[name]["out"]) < tol, name
        gbp = gc.GuidedBackPropagation(net)
        x = fx["imgs"].clone().cuda().requires_grad_(True)
        g = gbp(x)
        assert g.shape == x.shape and torch.isfinite(g).all()
    assert "feature shape:" in buf.getvalue() and "gradient shape:" in buf.getvalue()


def test_fused_vgg16_matches_the_hooked_network():
    """torchvision VGG16 (random weights: the pretrained ones are not available offline): the conv stack on this library's
    kernels (dge_b200/vgg_fused.py) against the same network run through its own modules with the reference's hooks --
    logits, the hooked feature / gradient of `features.28`, the Grad-CAM++ mask and the guided-backprop image gradient,
    with and without GuidedBackPropagation's ReLU hooks in force (E_mis_align_cropping_s1.py:99-106 registers both on one net).
    Gradients in the L2 norm: a ReLU / max-pool decision within rounding of a tie may differ between the two arithmetic
    orders and moves isolated entries only."""
    import numpy as np
    import metric.grad_cam as gc
    from torchvision.models import vgg16
    # the comparison network must compute in true fp32: cuDNN's default TF32 convs (10-bit operands) flip ReLU / max-pool
    # decisions against an fp32-equivalent run by themselves
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    net = vgg16(weights=None).cuda().eval()
    assert gc.vgg_fused.supported(net)
    g = torch.Generator().manual_seed(3)
    imgs = (torch.rand(2, 3, 64, 64, generator=g) * 2 - 1).cuda()
    index = np.array([17, 17])
    rel2 = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()

    def run(fused, with_gbp):
        gc.FUSED_VGG = fused
        net.__dict__.pop('_dge_fused_vgg', None)
        for m in net.modules():
            m._backward_hooks.clear()
            m._forward_hooks.clear()
        try:
            with contextlib.redirect_stdout(io.StringIO()) as buf:
                cam = gc.GradCamPlusPlus(net, "features.28")
                gbp = gc.GuidedBackPropagation(net) if with_gbp else None
                mask = cam(imgs, index)
                feat, grad = cam.feature.detach().clone(), cam.gradient.detach().clone()
                gi = None
                if gbp is not None:
                    x = imgs.clone().requires_grad_(True)
                    gi = gbp(x, index).detach().clone()
                cam.remove_handlers()
            assert "feature shape:" in buf.getvalue() and "gradient shape:" in buf.getvalue()
            return mask, feat, grad, gi
        finally:
            gc.FUSED_VGG = True

    with torch.no_grad():
        fused_logits = gc.vgg_fused.FusedVGG(net).forward(imgs)
        assert rel(fused_logits, net(imgs)) < 1e-3
    for with_gbp in (False, True):
        mf, ff, gf, xf = run(True, with_gbp)
        mh, fh, gh, xh = run(False, with_gbp)
        assert mf.dtype == torch.float64 and mf.shape == (2, 1, 64, 64)
        assert rel(ff, fh) < 1e-3                                  # the post-ReLU map the forward hook sees
        assert rel2(gf, gh) < 3e-3, with_gbp                       # the masked (and, with gbp, clamped) gradient
        assert rel2(mf, mh) < 3e-3, with_gbp
        if with_gbp:
            assert float(gf.min()) >= 0 and float(gh.min()) >= 0   # GuidedBackPropagation's hooks clamp Grad-CAM too
            assert xf.shape == imgs.shape and rel2(xf, xh) < 3e-3
