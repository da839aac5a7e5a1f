"""Oracle restatement of the LPIPS-VGG16 distance.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED.  The algorithm lives in a third-party dependency that is absent from /root/reference and from this
image: `lpips` (PyPI), listed without a version in the reference's requirements.txt:12 and called at E_align_s2.py:98,
embedding_img.py:61, comparing-baseline.py:15 (`lpips.LPIPS(net='vgg')`), consumed at training_utils.py:93.  Neither the
package nor its weights can be fetched (no network), and the reference holds no test vector for it.  This file restates
the published algorithm (Zhang, Isola, Efros, Shechtman, Wang: "The Unreasonable Effectiveness of Deep Features as a
Perceptual Metric", CVPR 2018; package v0.1, net='vgg', lpips=True, spatial=False) on torchvision's VGG16 layer stack,
as plain fp32 PyTorch over a state_dict with the package's key names; it pins the STRUCTURE of
deep-gan-encoders_b200/lpips (same weights -> same number), not the published weights."""
import torch
import torch.nn.functional as F

SHIFT = (-.030, -.088, -.188)
SCALE = (.458, .448, .450)
# conv indices of torchvision.models.vgg16().features per tap; a 2x2 max-pool precedes every tap but the first
TAPS = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21), (24, 26, 28))


def vgg16_taps(sd, x, prefix="net."):
    """relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 of VGG16."""
    taps, h = [], x
    for k, idxs in enumerate(TAPS):
        if k > 0:
            h = F.max_pool2d(h, 2, 2)
        for i in idxs:
            h = F.relu(F.conv2d(h, sd[f"{prefix}slice{k + 1}.{i}.weight"], sd[f"{prefix}slice{k + 1}.{i}.bias"],
                                padding=1))
        taps.append(h)
    return taps


def unit_normalize(x, eps=1e-10):
    return x / (x.pow(2).sum(dim=1, keepdim=True).sqrt() + eps)


def lpips_vgg(sd, in0, in1, normalize=False):
    """-> [N, 1, 1, 1]."""
    if normalize:
        in0, in1 = 2 * in0 - 1, 2 * in1 - 1
    shift = torch.tensor(SHIFT).view(1, 3, 1, 1).to(in0)
    scale = torch.tensor(SCALE).view(1, 3, 1, 1).to(in0)
    f0 = vgg16_taps(sd, (in0 - shift) / scale)
    f1 = vgg16_taps(sd, (in1 - shift) / scale)
    total = 0
    for k in range(5):
        d = (unit_normalize(f0[k]) - unit_normalize(f1[k])) ** 2
        total = total + F.conv2d(d, sd[f"lin{k}.model.1.weight"]).mean(dim=(2, 3), keepdim=True)
    return total
