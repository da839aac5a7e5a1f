"""Oracle restatement of the Grad-CAM post-processing (reference metric/grad_cam.py:101-126, 157-194, 234-251):
the host NumPy / cv2 arithmetic applied to the hooked feature and gradient tensors.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import cv2
import numpy as np
import torch


def class_index(output):
    """:162-165 -- per-image argmax and the bincount mode."""
    index = np.argmax(output.cpu().data.numpy(), axis=1)
    return index, int(np.argmax(np.bincount(index)))


def gradcam_pp(feature, gradient, out_hw):
    """GradCamPlusPlus.__call__ after the backward pass, :166-194.  feature/gradient: torch [N,C,h,w] fp32."""
    H, W = out_hw
    n = feature.shape[0]
    cam_all = np.zeros((n, H, W))
    for i in range(n):
        g = np.maximum(gradient[i].cpu().data.numpy(), 0.)
        indicate = np.where(g > 0, 1., 0.)
        norm_factor = np.sum(g, axis=(1, 2))
        for x in range(len(norm_factor)):
            norm_factor[x] = 1. / norm_factor[x] if norm_factor[x] > 0. else 0.
        alpha = indicate * norm_factor[:, np.newaxis, np.newaxis]
        weight = np.sum(g * alpha, axis=(1, 2))
        cam = np.sum(feature[i].cpu().data.numpy() * weight[:, np.newaxis, np.newaxis], axis=0)
        cam -= np.min(cam)
        cam /= np.max(cam)
        cam_all[i] = cv2.resize(cam, (W, H))
    return torch.tensor(cam_all.reshape(n, 1, H, W))


def gradcam(feature, gradient, out_hw):
    """GradCAM.__call__ after the backward pass, :113-127."""
    H, W = out_hw
    g = gradient.cpu().data.numpy()
    weight = np.mean(g, axis=(2, 3))
    cam = np.maximum(np.sum(feature.cpu().data.numpy() * weight[:, :, np.newaxis, np.newaxis], axis=1), 0)
    cam_all = np.zeros((cam.shape[0], H, W))
    for i, j in enumerate(cam):
        j -= np.min(j)
        j /= np.max(j)
        cam_all[i] = cv2.resize(j, (W, H))
    return torch.tensor(cam_all.reshape(cam.shape[0], 1, H, W))


def mask2cam(mask, imgs):
    """:234-251."""
    imgs = imgs.detach().clone().cpu()
    mask = mask.detach().clone().cpu()
    heatmap = np.float32(imgs).copy()
    cam = np.float32(imgs).copy()
    for i, j in enumerate(mask[:, 0]):
        h = cv2.applyColorMap(np.uint8(255 * j), cv2.COLORMAP_JET)
        h = np.float32(h) / 255
        h = np.transpose(h[..., ::-1], (2, 0, 1))
        heatmap[i] = h
        cam[i] = h + np.float32(imgs[i].numpy())
        cam[i] -= np.max(np.min(cam.copy()), 0)
        cam[i] /= np.max(cam[i])
    return torch.tensor(heatmap), torch.tensor(cam)
