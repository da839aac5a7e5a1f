"""GPU parity of the PGGAN generator (a9) and PGGAN encoder (a11) mirrors against golden fixtures from the
unmodified reference, plus BASELINE config[0] (PGGAN-256 + E_PG(64,7), batch 2) against the CPU oracle."""
import contextlib
import io
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-4


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


def test_pggan_generator_golden():
    from model.pggan.pggan_generator import PGGANGenerator
    fx = torch.load(os.path.join(GOLD, "pggan_res32.pt"))
    G = PGGANGenerator(**fx["config"])
    assert set(G.state_dict().keys()) == set(fx["state_dict"].keys())
    G.load_state_dict(fx["state_dict"], strict=True)
    G = G.cuda().eval()
    buf = io.StringIO()
    with torch.no_grad(), contextlib.redirect_stdout(buf):
        for lod, img in fx["images"].items():
            out = G(fx["z"].cuda(), lod=lod)
            assert rel(out["image"], img) < TOL, lod
        out = G(fx["z"].cuda())
        assert rel(out["z"], fx["z_out"]) < TOL
        for name, layer in (("block_up", G.layer4), ("block_plain", G.layer3), ("block_out", G.output1)):
            assert rel(layer(fx[name]["x"].cuda()), fx[name]["y"]) < TOL, name
    assert "torch.Size([2, 32, 32, 32])" in buf.getvalue()      # the reference's print(x.shape) side effect (:196)
    with pytest.raises(ValueError):
        G(torch.zeros(2, 3).cuda())
    with pytest.raises(ValueError):
        G(fx["z"].cuda(), lod=9)


def test_e_pg_golden():
    from model.E.E_PG import BE
    fx = torch.load(os.path.join(GOLD, "e_pg_s16_l4.pt"))
    E = BE(**fx["config"])
    assert list(E.state_dict().keys()) == list(fx["state_dict"].keys())
    E.load_state_dict(fx["state_dict"], strict=True)
    E = E.cuda().eval()
    with torch.no_grad():
        torch.manual_seed(8)
        for i, b in fx["blocks_seed8"].items():
            y, w1, w2 = E.decode_block[i](b["x"].cuda())
            assert (w1, w2) == (0, 0)
            assert rel(y, b["y"]) < TOL, i
        torch.manual_seed(8)
        assert rel(E.features(fx["img"].cuda()), fx["features_seed8"]) < TOL
        a, b = E(fx["img"].cuda())
    # the reference returns two 0-dim int64 zeros (E_PG.py:164)
    assert a.shape == b.shape == torch.Size([]) and a.dtype == torch.int64 and int(a) == 0 and int(b) == 0


def test_config0_pggan256_plumbing_vs_oracle():
    """BASELINE configs[0]: PGGAN-256 random-init G + E_PG(64,7,pggan=True), batch 2."""
    from model.E.E_PG import BE
    from model.pggan.pggan_generator import PGGANGenerator
    from oracle import pggan as opg
    torch.manual_seed(11)
    G = PGGANGenerator(256).eval()
    E = BE(64, 512, 7, 512, 3, pggan=True).eval()
    gen = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for m in (G, E):
            for k, p in m.named_parameters():
                if k.endswith(("bias", "bias_1", "bias_2", "noise_weight_1", "noise_weight_2")):
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.1)
        gsd = {k: v.clone() for k, v in G.state_dict().items()}
        esd = {k: v.clone() for k, v in E.state_dict().items()}
        z = torch.randn(2, 512, generator=gen)
        ref_img = opg.generator(gsd, z, 256)
        torch.manual_seed(3)
        ref_feat = opg.e_pg_features(esd, ref_img, 7)
        Gc, Ec = G.cuda(), E.cuda()
        with contextlib.redirect_stdout(io.StringIO()):
            img = Gc(z.cuda())["image"]
        assert img.shape == (2, 3, 256, 256)
        assert rel(img, ref_img) < TOL
        torch.manual_seed(3)
        feat = Ec.features(ref_img.cuda())
        assert feat.shape == (2, 512)
        assert rel(feat, ref_feat) < 1e-3
