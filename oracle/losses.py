"""Oracle restatement of `space_loss` (reference training_utils.py:54-99) and SSIM (metric/pytorch_ssim.py:8-38).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain PyTorch fp32."""
import math

import torch
import torch.nn.functional as F


def gaussian_window(channel, window_size=11, sigma=1.5):
    """create_window, metric/pytorch_ssim.py:8-16."""
    g = torch.Tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2d = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2d.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):
    """_ssim with size_average=True, metric/pytorch_ssim.py:18-38."""
    c = img1.shape[1]
    win = gaussian_window(c, window_size).to(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, win, padding=pad, groups=c)
    mu2 = F.conv2d(img2, win, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=c) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def space_loss(imgs1, imgs2, image_space=True, lpips_model=None):
    """training_utils.py:54-99.  Returns (loss tensor, [[mse, mse_mean, mse_std], kl, cos, ssim_loss, lpips])."""
    imgs1, imgs2 = imgs1.contiguous(), imgs2.contiguous()
    mse1 = F.mse_loss(imgs1, imgs2)
    mse2 = F.mse_loss(imgs1.mean(), imgs2.mean())
    mse3 = F.mse_loss(imgs1.std(), imgs2.std())
    dim = 0 if imgs1.ndim in (0, 1, 3) else 1                   # implicit-dim softmax (:68)
    k1, k2 = F.softmax(imgs1, dim=dim), F.softmax(imgs2, dim=dim)
    kl = F.kl_div(torch.log(k2), k1, reduction="mean")          # nn.KLDivLoss() default reduction (:56,69)
    kl = torch.where(torch.isnan(kl), torch.full_like(kl, 0), kl)
    kl = torch.where(torch.isinf(kl), torch.full_like(kl, 1), kl)
    a, b = imgs1.view(-1), imgs2.view(-1)
    cos = 1 - a.dot(b) / (torch.sqrt(a.dot(a)) * torch.sqrt(b.dot(b)))
    if image_space:
        while imgs1.shape[2] > 256:
            imgs1, imgs2 = F.avg_pool2d(imgs1, 2, 2), F.avg_pool2d(imgs2, 2, 2)
        ssim_l = 1 - ssim(imgs1, imgs2)
        lp = lpips_model(imgs1, imgs2).mean()
    else:
        ssim_l, lp = torch.tensor(0), torch.tensor(0)
    loss = 5 * mse1 + 3 * cos + ssim_l + 2 * lp
    return loss, [[mse1.item(), mse2.item(), mse3.item()], kl.item(), cos.item(), ssim_l.item(), lp.item()]
