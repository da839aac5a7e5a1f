"""Script tier on the CPU: the unmodified `E_align_s2.py` / `embedding_img.py` import and configure the drop-in package,
load their (synthetic) checkpoints into it, build the optimiser -- and stop at the first kernel call with DgeError, because
there is no CPU fallback.  The GPU run of the same harness is tests/test_scripts_gpu.py."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def run_tier(script, *flags):
    r = subprocess.run([sys.executable, os.path.join(HERE, "script_tier.py"), script, *flags], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


@pytest.mark.parametrize("script,extra", [("E_align_s2.py", ()), ("embedding_img.py", ()),
                                          ("E_align_s2.py", ("--mtype", "1")), ("E_align_s2.py", ("--mtype", "4")),
                                          ("E_mis_align_cropping_s1.py", ())])
def test_unmodified_script_reaches_the_first_kernel_and_refuses_the_cpu(script, extra):
    out = run_tier(script, "--cpu-plumbing", "--img-size", "32", *extra)
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["completed"] is False and "no CPU fallback" in out["dge_error"]


def test_training_utils_public_names_match_the_reference():
    """`from training_utils import *` is how every script gets its helpers AND torchvision / Image / truncnorm / F."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "deep-gan-encoders_b200"))
    import training_utils as tu
    reference_names = {"F", "Image", "get_para_GByte", "get_parameter_number", "imgPath2loader", "loader", "np", "one_hot",
                       "pytorch_ssim", "set_seed", "space_loss", "torch", "torchvision", "truncated_noise_sample",
                       "truncnorm"}          # dir(reference training_utils) minus dunders, recorded from the reference
    assert reference_names <= {n for n in dir(tu) if not n.startswith("_")}
    assert tu.get_para_GByte({"Total": 2 ** 27}) == {"Total_GB": 1.0, "Trainable_BG": 1.0}
