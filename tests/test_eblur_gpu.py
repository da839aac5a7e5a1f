"""GPU parity of the case-2 (blur / strided) encoder `model/E/E_Blur.py` (a7) and of the stride-2 4x4 tensor-core conv
(DGE_CONV_DOWN4X4S2) against the reference golden fixture, the oracle and a plain PyTorch fp32 conv."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-4


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.mark.parametrize("checker", [True, False], ids=["checker", "tcgen05"])
@pytest.mark.parametrize("case", [(2, 16, 32, 16, 24), (1, 32, 64, 34, 18), (2, 64, 128, 8, 8), (2, 128, 256, 16, 16)],
                         ids=lambda c: "n%d_ci%d_co%d_%dx%d" % c)
def test_strided_4x4_conv_vs_torch(case, checker):
    from dge_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, cin, cout, h, w = case
    g = torch.Generator(device="cuda").manual_seed(17 + cin)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g)
    w4 = torch.randn(cout, cin, 4, 4, device="cuda", generator=g) / (4.0 * cin ** 0.5)
    bias = torch.randn(cout, device="cuda", generator=g)
    # space-to-depth operand: channel block 2*py+px holds x[2y+py, 2x+px]
    s2d = torch.cat([x[:, :, py::2, px::2] for py in (0, 1) for px in (0, 1)], dim=1).contiguous()
    xa = ops.nchw_to_act(s2d)
    wpk = ops.pack_conv_weight(w4)
    out = ops.conv(xa, wpk, cout, ops.CONV_DOWN4X4S2, bias=bias, slope=0.2, out_nchw=True, checker=checker)["nchw"]
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.conv2d(x, w4, bias, stride=2, padding=1), 0.2)
    assert out.shape == ref.shape
    assert rel(out, ref) < TOL


def test_e_blur_golden():
    from model.E.E_Blur import BE
    fx = torch.load(os.path.join(GOLD, "e_blur_s16_l6.pt"))
    E = BE(**fx["config"])
    assert list(E.state_dict().keys()) == list(fx["state_dict"].keys())
    assert [b.fused_scale for b in E.decode_block] == fx["fused"]
    E.load_state_dict(fx["state_dict"], strict=True)
    E = E.cuda().eval()
    with torch.no_grad():
        for name, idx, seed in (("block0_seed71", 0, 71), ("block4_seed72", 4, 72)):
            b = fx[name]
            torch.manual_seed(seed)
            y, w1, w2 = E.decode_block[idx](b["x"].cuda())
            assert rel(y, b["y"]) < TOL and rel(w1, b["w1"]) < TOL and rel(w2, b["w2"]) < TOL, name
        torch.manual_seed(fx["noise_seed"])
        const, w = E(fx["img"].cuda())
    assert rel(const, fx["const"]) < TOL
    assert rel(w, fx["w"]) < TOL


def test_e_blur_256_vs_oracle():
    """E_Blur.BE(64, 512, 7) on 256x256 (embedding_img.py's encoder at the Cat-256 size), batch 2, vs the CPU oracle."""
    from model.E.E_Blur import BE
    from oracle import encoder as oenc
    torch.manual_seed(21)
    E = BE(64, 512, 7, 512, 3).eval()
    gen = torch.Generator().manual_seed(22)
    with torch.no_grad():
        for k, p in E.named_parameters():
            if k.endswith(("bias", "bias_1", "bias_2", "noise_weight_1", "noise_weight_2")):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        sd = {k: v.clone() for k, v in E.state_dict().items()}
        img = torch.randn(2, 3, 256, 256, generator=gen)
        torch.manual_seed(4)
        rc, rw = oenc.be_blur_forward(sd, img, 7)
        Ec = E.cuda()
        torch.manual_seed(4)
        c, w = Ec(img.cuda())
    assert rel(c, rc) < 1e-3 and rel(w, rw) < 1e-3
