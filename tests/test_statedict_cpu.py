"""Host-side (no GPU) checks: the mirrored modules keep the reference's state_dict layout, attributes and error
behaviour; the C-ABI library loads and exports every symbol include/dge_b200.h declares."""
import os
import re

import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from dge_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dge_b200.h")).read()
    declared = set(re.findall(r"\b(dge_[a-z0-9_]+)\s*\(", header))
    declared -= {"dge_conv_args"}
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/dge_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.dge_version() >= 100


def test_conv_args_struct_matches_header():
    import ctypes
    from dge_b200._lib import ConvArgs
    assert ctypes.sizeof(ConvArgs) == 232
    assert ConvArgs.noise_bstride.offset == 64 and ConvArgs.out_raw_up.offset == 184


def test_sg2_state_dict_layout_matches_reference():
    from model.stylegan2_generator import StyleGAN2Generator
    fx = torch.load(os.path.join(GOLD, "sg2_res32.pt"))
    G = StyleGAN2Generator(**fx["config"])
    ref = fx["state_dict"]
    mine = G.state_dict()
    assert set(mine.keys()) == set(ref.keys())
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k
        assert mine[k].dtype == ref[k].dtype, k
    G.load_state_dict(ref, strict=True)
    assert "synthesis.layer1.weight" in G.pth_to_tf_var_mapping


def test_sg2_1024_key_count_and_shapes():
    """StyleGAN2Generator(1024) has 165 state_dict entries."""
    from model.stylegan2_generator import StyleGAN2Generator
    with torch.device("meta"):
        G = StyleGAN2Generator(1024)
    sd = G.state_dict()
    assert len(sd) == 165
    assert tuple(sd["synthesis.layer16.weight"].shape) == (32, 32, 3, 3)
    assert tuple(sd["synthesis.layer15.filter.kernel"].shape) == (1, 1, 4, 4)
    assert tuple(sd["synthesis.output8.weight"].shape) == (3, 32, 1, 1)
    assert tuple(sd["synthesis.layer9.style.weight"].shape) == (512, 512)
    assert tuple(sd["synthesis.upsample.kernel"].shape) == (1, 1, 4, 4)


def test_sg2_errors():
    from model.stylegan2_generator import StyleGAN2Generator
    with pytest.raises(ValueError):
        StyleGAN2Generator(100)
    with pytest.raises(ValueError):
        StyleGAN2Generator(32, architecture="bogus")
    G = StyleGAN2Generator(8, z_space_dim=16, w_space_dim=16, mapping_fmaps=16, fmaps_base=128, fmaps_max=16)
    with pytest.raises(ValueError):
        G.mapping(torch.zeros(2, 3))
    with pytest.raises(ValueError):
        G.synthesis(torch.zeros(2, 3, 16))


def test_be_state_dict_layout_and_lreq_coefs():
    from model.E.E import BE
    fx = torch.load(os.path.join(GOLD, "be_s16_l4.pt"))
    E = BE(**fx["config"])
    ref = fx["state_dict"]
    mine = E.state_dict()
    assert list(mine.keys()) == list(ref.keys())        # same keys, same (registration) order
    for k in ref:
        assert tuple(mine[k].shape) == tuple(ref[k].shape), k
    E.load_state_dict(ref, strict=True)
    # lr_equalization_coef contract read by LREQAdam (lreq.py:60-62,118-120)
    c = E.decode_block[0].conv_1
    assert abs(c.weight.lr_equalization_coef - (2 ** 0.5) / (9 * 16) ** 0.5) < 1e-7
    lin = E.decode_block[0].inver_mod1
    assert abs(lin.weight.lr_equalization_coef - 1.0 / (32 ** 0.5)) < 1e-7
    assert E.decode_block[0].conv_3.bias.lr_equalization_coef == 1.0
    assert not hasattr(E.decode_block[0].bias_1, "lr_equalization_coef")


def test_product_path_fails_loudly_without_gpu():
    """No CPU fallback: CPU tensors raise instead of silently computing something else."""
    from dge_b200 import DgeError
    from model.E.E import BE
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    E = BE(16, 32, 2, 512, 3)
    with torch.no_grad(), pytest.raises(DgeError):
        E(torch.zeros(1, 3, 8, 8))


def test_header_is_plain_c_and_structs_match_ctypes_field_by_field(tmp_path):
    """include/dge_b200.h must compile as C99 (the boundary is a C ABI, not C++), and every field of the two structs
    that cross it must sit at the offset the ctypes mirror assumes."""
    import ctypes
    import shutil
    import subprocess
    from dge_b200._lib import ConvArgs, Sg2PrepItem
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    pairs = [("dge_conv_args", ConvArgs), ("dge_sg2_prep_item", Sg2PrepItem)]
    lines = ['#include "dge_b200.h"', "#include <stdio.h>", "int main(void) {"]
    for cname, st in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe)], check=True, capture_output=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, st in pairs:
        assert int(got[cname]) == ctypes.sizeof(st), cname
        for fname, _ in st._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(st, fname).offset, f"{cname}.{fname}"


def test_ctypes_signatures_have_the_declared_argument_counts():
    """Every prototype in the header against the ctypes table: same number of parameters."""
    from dge_b200 import _lib
    header = open(os.path.join(ROOT, "include", "dge_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    protos = re.findall(r"\b(dge_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", header)
    assert len(protos) >= 50
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        restype, argtypes = _lib.SIGNATURES[name]
        assert len(argtypes) == n, f"{name}: header has {n} parameters, ctypes table {len(argtypes)}"


def test_wgrad_entry_point_validates_arguments_before_touching_the_device():
    """Error behaviour of the C ABI (no GPU needed): bad arguments come back as a negative code + message."""
    import ctypes
    from dge_b200 import _lib
    lib = _lib.load()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    cases = [
        ((None, p, p, 1, 16, 16, 8, 8, 3, 2, 0, None), "null pointer"),
        ((p, p, p, 1, 16, 16, 8, 8, 2, 2, 0, None), "ksize"),
        ((p, p, p, 1, 16, 16, 8, 8, 3, 3, 0, None), "planes"),
        ((p, p, p, 0, 16, 16, 8, 8, 3, 2, 0, None), "empty input"),
        ((p, p, p, 1, 12, 16, 8, 8, 3, 2, 0, None), "multiples of 8"),
    ]
    for args, needle in cases:
        rc = lib.dge_conv_wgrad(*args)
        assert rc < 0
        assert needle in lib.dge_last_error().decode(), (needle, lib.dge_last_error())
