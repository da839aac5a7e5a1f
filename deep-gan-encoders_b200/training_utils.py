"""Drop-in for the reference's `training_utils.py`: `space_loss` (:54-99) as fused device reductions, plus the
small host helpers the scripts import (`set_seed` :46-51, `truncated_noise_sample` :32-44, `one_hot` :27-30,
`get_parameter_number` :17-20, `imgPath2loader` / `loader` :10-15, `get_para_GByte` :22-25).  The scripts do
`from training_utils import *` and rely on the names that leaks (`torchvision`, `Image`, `truncnorm`, `F`, `np`,
`torch`, `pytorch_ssim`), so the same modules are imported at module level here.

`space_loss` semantics kept: MSE / mean-MSE / std-MSE (unbiased std), implicit-dim softmax KL (logged only,
NaN -> 0, inf -> 1), cosine over the flattened WHOLE batch, avg-pool while H > 256, SSIM, LPIPS through the
caller's `lpips_model`, `loss = 5*mse + 3*cos + (1-ssim) + 2*lpips`, `loss_info` of Python floats.
One pass over the image pair produces all six moments; the whole call does ONE device->host read instead of
the reference's seven `.item()` syncs.  When an input requires grad (the training scripts back-propagate through
this loss) the loss terms are recorded for autograd instead -- see `_space_loss_autograd`.
"""
import math

import numpy as np
import torch
import torchvision
from PIL import Image
from scipy.stats import truncnorm
from torch.nn import functional as F

import metric.pytorch_ssim as pytorch_ssim
from dge_b200 import ops

# host-side image loading (embedding_img.py:214, embedding_v2_*.py): PIL file -> RGB -> size x size -> [3, size, size] in [0, 1]
loader = torchvision.transforms.Compose([torchvision.transforms.ToTensor()])


def imgPath2loader(image_name, size):
    image = Image.open(image_name).convert('RGB').resize((size, size))
    return loader(image).to(torch.float)


def get_para_GByte(parameter_number):
    # the reference reports the TOTAL count under both keys (and spells the second one 'Trainable_BG'), :22-25
    gb = parameter_number['Total'] * 8 / 1024 / 1024 / 1024
    return {'Total_GB': gb, 'Trainable_BG': gb}


def get_parameter_number(net):
    total_num = sum(p.numel() for p in net.parameters())
    trainable_num = sum(p.numel() for p in net.parameters() if p.requires_grad)
    return {'Total': total_num, 'Trainable': trainable_num}


def one_hot(x, class_count=1000):
    return torch.eye(class_count)[x, :]


def truncated_noise_sample(batch_size=1, dim_z=128, truncation=1., seed=None):
    state = None if seed is None else np.random.RandomState(seed)
    values = truncnorm.rvs(-2, 2, size=(batch_size, dim_z), random_state=state).astype(np.float32)
    return truncation * values


def set_seed(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True


def avg_pool_to_256(x):
    """`while x.shape[2] > 256: x = F.avg_pool2d(x, 2, 2)` (:81-84) as one pooling kernel."""
    f = 1
    h, w = x.shape[2], x.shape[3]
    while h > 256:
        h, w, f = h // 2, w // 2, f * 2
    if f == 1:
        return x
    out = torch.empty((x.shape[0], x.shape[1], h, w), dtype=torch.float32, device=x.device)
    ops.check(ops.lib().dge_avgpool_nchw(ops._p(x), ops._p(out), x.shape[0] * x.shape[1], h, w, f, ops._stream()))
    return out


def _pair_stats(a, b):
    """One pass of `dge_pair_moments` + one of `dge_softmax_kl_sum` over two equal-sized fp32 tensors and ONE host
    sync -> dict of the scalar statistics `space_loss` reports (training_utils.py:63-77)."""
    dev, n = a.device, a.numel()
    L = ops.lib()
    acc = torch.zeros(8, dtype=torch.float64, device=dev)   # 0..5 moments, 6 kl sum, 7 ssim sum (filled by the caller)
    ops.check(L.dge_pair_moments(ops._p(a), ops._p(b), n, ops._p(acc), ops._stream()))
    # implicit-dim softmax: dim 0 for 0/1/3-D inputs, dim 1 otherwise (torch.nn.functional._get_softmax_dim)
    dim = 0 if a.ndim in (0, 1, 3) else 1
    shape = list(a.shape)
    outer = int(np.prod(shape[:dim])) if dim > 0 else 1
    inner = int(np.prod(shape[dim + 1:])) if dim + 1 < len(shape) else 1
    ops.check(L.dge_softmax_kl_sum(ops._p(a), ops._p(b), outer, shape[dim], inner, ops._p(acc[6:7]), ops._stream()))
    return acc


def _stats_from_sums(host, n):
    sa, sb, saa, sbb, sab, sdd, kl_sum = host[:7]
    mse1 = sdd / n
    m1, m2 = sa / n, sb / n
    mse2 = (m1 - m2) ** 2
    if n > 1:
        std1 = math.sqrt(max((saa - n * m1 * m1) / (n - 1), 0.0))
        std2 = math.sqrt(max((sbb - n * m2 * m2) / (n - 1), 0.0))
    else:
        std1 = std2 = float('nan')
    mse3 = (std1 - std2) ** 2
    kl = kl_sum / n
    if math.isnan(kl):
        kl = 0.0
    if math.isinf(kl):
        kl = 1.0
    denom = math.sqrt(saa) * math.sqrt(sbb)
    cos = 1 - (sab / denom if denom > 0 else float('nan'))
    return mse1, mse2, mse3, kl, cos


class _MseCosFn(torch.autograd.Function):
    """(MSE, 1 - cosine) of two tensors from the six sums `dge_pair_moments` already produced for the log (:63, :74-76).
    Backward is linear in the inputs: d/db = alpha * a + beta * b with two device scalars (no host sync)."""

    @staticmethod
    def forward(ctx, a, b, acc):
        n = a.numel()
        sa, sb, saa, sbb, sab, sdd = acc[0], acc[1], acc[2], acc[3], acc[4], acc[5]
        ctx.save_for_backward(a, b, acc)
        mse = (sdd / n).float()
        cos = (1 - sab / (saa.sqrt() * sbb.sqrt())).float()
        return mse, cos

    @staticmethod
    def backward(ctx, g_mse, g_cos):
        a, b, acc = ctx.saved_tensors
        n = a.numel()
        saa, sbb, sab = acc[2], acc[3], acc[4]
        na, nb = saa.sqrt(), sbb.sqrt()
        gm, gc = g_mse.double(), g_cos.double()
        grads = [None, None]
        # d mse/d b = 2 (b - a) / n ;  d (1 - cos)/d b = -a / (|a||b|) + (a.b) b / (|a||b|^3)   (and symmetrically for a)
        if ctx.needs_input_grad[1]:
            alpha = (-2 * gm / n - gc / (na * nb)).float()
            beta = (2 * gm / n + gc * sab / (na * nb * sbb)).float()
            grads[1] = torch.addcmul(a * alpha, b, beta)
        if ctx.needs_input_grad[0]:
            alpha = (-2 * gm / n - gc / (na * nb)).float()
            beta = (2 * gm / n + gc * sab / (na * nb * saa)).float()
            grads[0] = torch.addcmul(b * alpha, a, beta)
        return grads[0], grads[1], None


class _AvgPoolFn(torch.autograd.Function):
    """`while x.shape[2] > 256: x = F.avg_pool2d(x, 2, 2)` (:81-84) as one f x f mean; backward = broadcast / f^2."""

    @staticmethod
    def forward(ctx, x, f):
        ctx.f = f
        return avg_pool_to_256(x.contiguous())

    @staticmethod
    def backward(ctx, g):
        f = ctx.f
        return F.interpolate(g, scale_factor=f, mode='nearest') / float(f * f), None


FUSED_TRAIN = True     # False: the loss terms as separate torch nodes (cross-check of the fused nodes)


def _space_loss_autograd(a, b, image_space, lpips_model):
    """Training form (E_align_s2.py:184-205 back-propagates through this loss).  The log statistics and the MSE / cosine
    terms come from ONE pass of the fused moments kernel (`_MseCosFn`: their gradient is alpha*a + beta*b), the pooled SSIM
    and LPIPS terms are fused nodes too (metric/pytorch_ssim.py, dge_b200/train_lpips.py); one host read-back per call."""
    n = a.numel()
    acc = _pair_stats(a.detach(), b.detach())
    if FUSED_TRAIN:
        mse, cos = _MseCosFn.apply(a, b, acc)
    else:
        fa, fb = a.reshape(-1), b.reshape(-1)
        mse = (fa - fb).square().mean()                                            # :63
        cos = 1 - fa.dot(fb) / (fa.dot(fa).sqrt() * fb.dot(fb).sqrt())             # :74-76
    if image_space:
        f = 1
        while a.shape[2] // f > 256:                                               # :81-84 (2x2 means compose)
            f *= 2
        if f == 1:
            pa, pb = a, b
        elif FUSED_TRAIN:
            pa, pb = _AvgPoolFn.apply(a, f), _AvgPoolFn.apply(b, f)
        else:
            pa, pb = F.avg_pool2d(a, f, f), F.avg_pool2d(b, f, f)
        ssim_l = 1 - pytorch_ssim.ssim(pa, pb)                                     # :87-88
        lp = lpips_model(pa, pb).mean()                                            # :93
        extra = torch.stack((cos.detach().double(), ssim_l.detach().double(), lp.detach().double()))
    else:
        ssim_l, lp = torch.tensor(0), torch.tensor(0)                              # :90, 95
        extra = cos.detach().double().view(1)
    loss = 5 * mse + 3 * cos + ssim_l + 2 * lp                                     # :97
    host = torch.cat((acc, extra)).cpu().tolist()                                  # the one host read-back of this call
    mse1, mse2, mse3, kl, _ = _stats_from_sums(host[:8], n)
    if image_space:
        return loss, [[mse1, mse2, mse3], kl, host[8], host[9], host[10]]
    return loss, [[mse1, mse2, mse3], kl, host[8], 0, 0]


def space_loss(imgs1, imgs2, image_space=True, lpips_model=None):
    if not (imgs1.is_cuda and imgs2.is_cuda):
        raise ops.DgeError('space_loss: dge_b200 runs on a B200 only; there is no CPU fallback')
    if imgs1.reshape(-1).shape[0] != imgs2.reshape(-1).shape[0]:
        print('error: vector1 dimentions are not equal to vector2 dimentions')
        return
    a = imgs1.float().contiguous()
    b = imgs2.float().contiguous()
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return _space_loss_autograd(a, b, image_space, lpips_model)
    dev = a.device
    n = a.numel()
    L = ops.lib()
    acc = _pair_stats(a, b)
    lp = None
    if image_space:
        pa, pb = avg_pool_to_256(a), avg_pool_to_256(b)
        ops.check(L.dge_ssim_sum(ops._p(pa), ops._p(pb), pa.shape[0] * pa.shape[1], pa.shape[2], pa.shape[3],
                                 ops._p(acc[7:8]), ops._stream()))
        lp = lpips_model(pa, pb).mean()      # third-party LPIPS (unpinned) stays the caller's module
        n_ssim = pa.numel()
    host = acc.cpu().tolist()               # the one sync of this call
    mse1, mse2, mse3, kl, cos = _stats_from_sums(host, n)
    if image_space:
        ssim_l = 1 - host[7] / n_ssim
        lp_v = float(lp.item())
    else:
        ssim_l, lp_v = 0, 0
    loss = 5 * mse1 + 3 * cos + ssim_l + 2 * lp_v
    if image_space and lp.requires_grad:
        # neither image carries a gradient but the caller's LPIPS module does (its `lin` weights): the reference's loss then
        # still has a grad_fn, and E_mis_align_cropping_s1.py:187-193 calls backward() on exactly such losses
        loss_t = torch.tensor(5 * mse1 + 3 * cos + ssim_l, dtype=torch.float32, device=dev) + 2 * lp
    else:
        loss_t = torch.tensor(loss, dtype=torch.float32, device=dev)
    loss_info = [[mse1, mse2, mse3], kl, cos, ssim_l, lp_v]
    return loss_t, loss_info
