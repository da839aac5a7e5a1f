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
