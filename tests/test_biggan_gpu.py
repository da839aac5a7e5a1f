"""GPU parity of the BigGAN-deep generator (a10) and the BigGAN encoder E_BIG (a12) mirrors against golden fixtures from
the unmodified reference (eval mode, spectral-norm u/v converged) and the CPU oracle."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 3e-4


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.fixture(scope="module")
def big():
    from model.biggan_generator import BigGAN
    from model.utils.biggan_config import BigGANConfig
    fx = torch.load(os.path.join(GOLD, "biggan_small.pt"))
    G = BigGAN(BigGANConfig.from_dict(fx["config"]))
    assert set(G.state_dict().keys()) == set(fx["state_dict"].keys())
    G.load_state_dict(fx["state_dict"], strict=True)
    return fx, G.cuda().eval()


def test_biggan_forward(big):
    fx, G = big
    with torch.no_grad():
        for trunc, img in fx["images"].items():
            out, cond = G(fx["z"].cuda(), fx["label"].cuda(), trunc)
            assert out.shape == img.shape and rel(cond, fx["cond"]) < TOL
            assert rel(out, img) < TOL, trunc
        # the scripts pass the truncation as a CUDA tensor (E_align_s2.py:148)
        out, _ = G(fx["z"].cuda(), fx["label"].cuda(), torch.tensor(0.4).cuda())
        assert rel(out, fx["images"][0.4]) < TOL
        with pytest.raises(AssertionError):
            G(fx["z"].cuda(), fx["label"].cuda(), 0.0)


def test_biggan_blocks(big):
    fx, G = big
    cond = fx["cond"].cuda()
    with torch.no_grad():
        assert rel(G.generator.layers[2](fx["attn"]["x"].cuda()), fx["attn"]["y"]) < TOL
        assert rel(G.generator.layers[3](fx["block_up_drop"]["x"].cuda(), cond, 0.4), fx["block_up_drop"]["y"]) < TOL


def test_e_big_golden():
    from model.E.E_BIG import BE
    fx = torch.load(os.path.join(GOLD, "e_big_s16_l4.pt"))
    E = BE(**fx["config"])
    assert set(E.state_dict().keys()) == set(fx["state_dict"].keys())
    E.load_state_dict(fx["state_dict"], strict=True)
    E = E.cuda().eval()
    cond = fx["cond"].cuda()
    with torch.no_grad():
        torch.manual_seed(13)
        for i, b in fx["blocks_seed13"].items():
            y, _, _ = E.decode_block[i](b["x"].cuda(), cond, truncation=0.4)
            assert rel(y, b["y"]) < TOL, i
        torch.manual_seed(13)
        assert rel(E.features(fx["img"].cuda(), cond), fx["features_seed13"]) < TOL


def test_e_big_heads_vs_oracle():
    """E_BIG.BE(64, 512, 4?) is too small for the 8192-wide head, so run the real C4 encoder E_BIG(64,512,7) at 256, N=2."""
    from model.E.E_BIG import BE
    from oracle import biggan as obg
    torch.manual_seed(7)
    E = BE(64, 512, 7, 512, 3, biggan=True).eval()
    gen = torch.Generator().manual_seed(8)
    with torch.no_grad():
        for k, p in E.named_parameters():
            if k.endswith(("bias", "bias_1", "bias_2", "noise_weight_1", "noise_weight_2")):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        img = torch.randn(2, 3, 256, 256, generator=gen)
        cond = torch.randn(2, 256, generator=gen) * 0.3
        E.train()                       # converge the spectral-norm vectors of the CBN linears, then freeze
        Ec = E.cuda()
        Ec.eval()
        sd = {k: v.detach().cpu().clone() for k, v in Ec.state_dict().items()}
        torch.manual_seed(2)
        rc, rz = obg.e_big_forward(sd, img, cond, 7)
        torch.manual_seed(2)
        c_v, z = Ec(img.cuda(), cond.cuda())
    assert c_v.shape == (2, 256) and z.shape == (2, 128)
    assert rel(c_v, rc) < 1e-3 and rel(z, rz) < 1e-3
