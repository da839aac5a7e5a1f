"""GPU parity of the StyleGAN1 generator mirror (a8) against golden fixtures from the unmodified reference and the
CPU oracle: Mapping, Generator.forward at several lods, per-block vectors incl. the const/batch-1 first block and the
fused-scale (transposed conv + 4-tap transform + blur) block."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-4


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


@pytest.fixture(scope="module")
def sg1():
    from model.stylegan1.net import Generator, Mapping
    fx = torch.load(os.path.join(GOLD, "sg1_l6.pt"))
    Gs = Generator(**fx["config"])
    assert set(Gs.state_dict().keys()) == set(fx["state_dict"].keys())
    Gs.load_state_dict(fx["state_dict"], strict=True)
    Gm = Mapping(num_layers=12, mapping_layers=3, latent_size=64, dlatent_size=64, mapping_fmaps=64)
    Gm.buffer1 = torch.zeros(12, 64)     # the scripts assign buffer1 after construction (E_align_s2.py:35-41)
    Gm.load_state_dict(fx["map_state_dict"], strict=True)
    return fx, Gs.cuda().eval(), Gm.cuda().eval()


def test_sg1_mapping(sg1):
    fx, Gs, Gm = sg1
    with torch.no_grad():
        styles = Gm(fx["z"].cuda(), fx["coefs"].cuda())
    assert rel(styles, fx["styles"]) < TOL


def test_sg1_generator_forward(sg1):
    fx, Gs, Gm = sg1
    with torch.no_grad():
        for lod, img in fx["images"].items():
            torch.manual_seed(60 + lod)
            out = Gs.forward(fx["styles"].cuda(), lod)
            assert out.shape == img.shape
            assert rel(out, img) < TOL, lod
    with pytest.raises(NotImplementedError):
        Gs.forward(fx["styles"].cuda(), 3, blend=0.5)


def test_sg1_blocks(sg1):
    fx, Gs, Gm = sg1
    st = fx["styles"].cuda()
    with torch.no_grad():
        torch.manual_seed(9)
        for i in range(3):
            b = fx["blocks_seed9"][i]
            y = Gs.decode_block[i](b["x"].cuda(), st[:, 2 * i], st[:, 2 * i + 1])
            assert rel(y, b["y"]) < TOL, i
        # keep the RNG stream aligned with the fixture: blocks 3..5 were also run under seed 9 upstream (unused here)
        fb = fx["fused_block_seed10"]
        torch.manual_seed(10)
        y = Gs.decode_block[5](fb["x"].cuda(), st[:, 10], st[:, 11])
        assert rel(y, fb["y"]) < TOL


def test_config1_stylegan1_256_vs_oracle():
    """BASELINE configs[1] shapes: Generator(64,512,7) at 256 (two fused-scale blocks), batch 4, vs the CPU oracle."""
    from model.stylegan1.net import Generator
    from oracle import stylegan1 as osg1
    torch.manual_seed(5)
    Gs = Generator(64, 512, 7, 512, 3).eval()
    gen = torch.Generator().manual_seed(6)
    with torch.no_grad():
        for k, p in Gs.named_parameters():
            if k.endswith(("bias", "bias_1", "bias_2", "noise_weight_1", "noise_weight_2")):
                p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        Gs.const.copy_(torch.randn(Gs.const.shape, generator=gen))
        sd = {k: v.clone() for k, v in Gs.state_dict().items()}
        styles = torch.randn(4, 14, 512, generator=gen)
        torch.manual_seed(1)
        ref = osg1.decode(sd, styles, 6)
        Gc = Gs.cuda()
        torch.manual_seed(1)
        out = Gc.forward(styles.cuda(), 6)
    assert out.shape == (4, 3, 256, 256)
    assert rel(out, ref) < 1e-3
