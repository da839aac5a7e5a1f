"""Script tier on the B200 (north star: "E_align/embedding_img.py drop in unchanged"): the UNMODIFIED
reference scripts' `train()` run end to end against the drop-in package -- synthetic checkpoints, two iterations, every
forward / backward / optimiser step on dge_b200 kernels -- and write the files they are meant to write."""
import pytest

from test_scripts_cpu import run_tier

pytestmark = pytest.mark.gpu


def test_e_align_s2_trains_two_iterations_unmodified():
    out = run_tier("E_align_s2.py", "--img-size", "64", "--iterations", "2")
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["completed"] is True, out
    assert out["missing_outputs"] == [], out
    assert out["dge_launches"] > 200, out                     # the work ran on this library's kernels
    assert out["e_checkpoint_keys"] > 50 and out["e_checkpoint_finite"], out


def test_embedding_img_inverts_two_images_unmodified():
    out = run_tier("embedding_img.py", "--img-size", "64", "--iterations", "2")
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["completed"] is True, out
    assert out["missing_outputs"] == [], out
    assert out["dge_launches"] > 200, out


@pytest.mark.parametrize("mtype", [1, 4])
def test_e_align_s2_other_generator_families_unmodified(mtype):
    """The same script with --mtype 1 (StyleGAN1: `Gm` stays on the CPU as upstream, E_align_s2.py:33-44,108) and --mtype 4
    (BigGAN-deep + the class-conditional encoder E_BIG)."""
    out = run_tier("E_align_s2.py", "--img-size", "64", "--iterations", "2", "--mtype", str(mtype))
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["completed"] is True, out
    assert out["missing_outputs"] == [], out
    assert out["dge_launches"] > 200, out
    assert out["e_checkpoint_keys"] > 50 and out["e_checkpoint_finite"], out


def test_e_mis_align_cropping_trains_two_iterations_unmodified():
    """The Grad-CAM variant of the training script (E_mis_align_cropping_s1.py: Grad-CAM++ masks, guided back-propagation and
    mask2cam on a torchvision VGG16 every iteration) -- unmodified, on the drop-in package."""
    out = run_tier("E_mis_align_cropping_s1.py", "--img-size", "64", "--iterations", "2")
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["completed"] is True, out
    assert out["missing_outputs"] == [], out
    assert out["dge_launches"] > 300, out
    assert out["e_checkpoint_keys"] > 50 and out["e_checkpoint_finite"], out


def test_e_align_cropping_s1_trains_two_iterations_unmodified():
    """E_align_cropping_s1.py: the E_align loop whose three image losses are all computed on detached clones."""
    out = run_tier("E_align_cropping_s1.py", "--img-size", "64", "--iterations", "2")
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["completed"] is True, out
    assert out["missing_outputs"] == [], out
    assert out["dge_launches"] > 200, out
    assert out["e_checkpoint_keys"] > 50 and out["e_checkpoint_finite"], out
