"""Grad-CAM / Grad-CAM++ / guided back-propagation -- drop-in for the reference's `metric/grad_cam.py`
(GradCAM :11-127, GradCamPlusPlus :129-194, GuidedBackPropagation :196-232, mask2cam :234-251).

The classifier (`net`, a torchvision VGG16 in E_mis_align_cropping_s1.py:99-106) stays the caller's PyTorch module and the
owner of its weights.  When it has the torchvision-VGG layout its convolutional stack runs forward and backward on this
library's kernels (dge_b200/vgg_fused.py: convs with bias + ReLU epilogues, ReLU / max-pool backward with the guided clamp,
tcgen05 data-gradient convs; the fully-connected head stays the caller's modules) and reports exactly what the reference's
hooks observe; any other network runs through its own modules with the reference's hooks (`FUSED_VGG = False` forces that).
What the reference does AFTER the backward pass on the host -- per-image NumPy loops, `cv2.resize`, `cv2.applyColorMap`,
several device->host copies -- runs on the device: class-index argmax / bincount mode (bit-exact), channel weights,
the CAM sum, min/max normalisation, the bilinear resize and the JET overlay.  Outputs keep the reference's dtypes:
`__call__` -> float64 [N,1,H,W] (a CUDA tensor here; the reference returns a CPU tensor that the scripts move to the GPU).
"""
import ctypes

import numpy as np
import torch

from dge_b200 import ops
from dge_b200 import vgg_fused

FUSED_VGG = True      # False: always run the caller's network through its own modules + hooks


def _fused_runner(net, layer_name=None):
    """The shared FusedVGG of a network (cached on the module) when the fused path applies, else None."""
    if not FUSED_VGG or not vgg_fused.supported(net) or net.training:
        return None
    r = net.__dict__.get('_dge_fused_vgg')
    if r is None:
        r = vgg_fused.FusedVGG(net)
        net.__dict__['_dge_fused_vgg'] = r
    if layer_name is not None and layer_name not in r.names:
        return None
    return r


def _class_index(output, index):
    """np.argmax(output, axis=1) and np.argmax(np.bincount(index)) on the device (grad_cam.py:162-165)."""
    L = ops.lib()
    n, k = output.shape
    idx = torch.empty(n, dtype=torch.int64, device=output.device)
    mode = torch.empty(1, dtype=torch.int64, device=output.device)
    if index is None:
        ops.check(L.dge_argmax_mode(ops._p(output.detach().float().contiguous()), n, k, ops._p(idx), ops._p(mode),
                                    ops._stream()))
        return idx, int(mode.item())
    index = np.asarray(index)
    return torch.as_tensor(index, device=output.device), int(np.argmax(np.bincount(index)))


def cam_maps(feature, gradient, out_hw, plus):
    """Hooked feature / gradient [N,C,h,w] -> float64 [N,1,H,W] (dge_gradcam)."""
    f = feature.detach().float().contiguous()
    g = gradient.detach().float().contiguous()
    n, c, h, w = f.shape
    H, W = out_hw
    dev = f.device
    wbuf = torch.empty(n * c, dtype=torch.float64, device=dev)
    cbuf = torch.empty(n * h * w, dtype=torch.float64, device=dev)
    out = torch.empty((n, 1, H, W), dtype=torch.float64, device=dev)
    ops.check(ops.lib().dge_gradcam(ops._p(f), ops._p(g), int(plus), ops._p(wbuf), ops._p(cbuf), ops._p(out), n, c, h, w,
                                    H, W, ops._stream()))
    return out


class GradCAM(object):
    plus = False

    def __init__(self, net, layer_name):
        self.net = net
        self.layer_name = layer_name
        self.feature = None
        self.gradient = None
        self.net.eval()
        self.handlers = []
        self._register_hook()

    def _get_features_hook(self, module, input, output):
        self.feature = output
        print("feature shape:{}".format(output.size()))

    def _get_grads_hook(self, module, input_grad, output_grad):
        self.gradient = output_grad[0]
        print("gradient shape:{}".format(output_grad[0].size()))

    def _register_hook(self):
        for (name, module) in self.net.named_modules():
            if name == self.layer_name:
                self.handlers.append(module.register_forward_hook(self._get_features_hook))
                # legacy (non-full) hook, as upstream: the hooked conv feeds an in-place ReLU, which full hooks reject
                self.handlers.append(module.register_backward_hook(self._get_grads_hook))

    def remove_handlers(self):
        for handle in self.handlers:
            handle.remove()

    def __call__(self, inputs, index):
        if not inputs.is_cuda:
            raise ops.DgeError('GradCAM: dge_b200 runs on a B200 only; there is no CPU fallback')
        self.net.zero_grad()
        fused = _fused_runner(self.net, self.layer_name)
        if fused is not None:
            output = fused.forward(inputs)                          # conv stack on dge_b200 kernels
            self.feature = fused.feature(self.layer_name)           # what the forward hook saw (post in-place ReLU)
            print("feature shape:{}".format(self.feature.size()))
            _, index_max = _class_index(output, index)
            output[:, index_max].mean().backward(retain_graph=True)
            self.gradient = fused.backward(stop_at=self.layer_name)
            print("gradient shape:{}".format(self.gradient.size()))
            return cam_maps(self.feature, self.gradient, (inputs.size(2), inputs.size(3)), self.plus)
        output = self.net(inputs)                                   # [N, num_classes]
        _, index_max = _class_index(output, index)
        target = output[:, index_max].mean()
        target.backward(retain_graph=True)
        return cam_maps(self.feature, self.gradient, (inputs.size(2), inputs.size(3)), self.plus)


class GradCamPlusPlus(GradCAM):
    plus = True

    def __init__(self, net, layer_name):
        super().__init__(net, layer_name)


class GuidedBackPropagation(object):
    """Reference :196-232 -- ReLU backward hooks on the caller's network (they also stay in force for every later backward
    through it, Grad-CAM's included); a torchvision VGG runs its conv stack on the fused kernels with the same clamp."""

    def __init__(self, net):
        self.net = net
        for (name, module) in self.net.named_modules():
            if isinstance(module, torch.nn.ReLU):
                module.register_backward_hook(self.backward_hook)
        self.net.eval()

    @classmethod
    def backward_hook(cls, module, grad_in, grad_out):
        return torch.clamp(grad_in[0], min=0.0),

    def __call__(self, inputs, index=None):
        self.net.zero_grad()
        fused = _fused_runner(self.net) if inputs.is_cuda and inputs.requires_grad else None
        if fused is not None:
            output = fused.forward(inputs)
            _, index_max = _class_index(output, index)
            output[:, index_max].mean().backward(retain_graph=True)
            g = fused.backward(stop_at=None).to(inputs.dtype)
            inputs.grad = g if inputs.grad is None else inputs.grad + g
            return inputs.grad
        output = self.net(inputs)
        _, index_max = _class_index(output, index)
        target = output[:, index_max].mean()
        target.backward(retain_graph=True)
        return inputs.grad


_JET_RGB = None


def _jet_lut(device):
    """cv2.COLORMAP_JET as a 256x3 RGB table (a constant; queried once from OpenCV)."""
    global _JET_RGB
    if _JET_RGB is None:
        import cv2
        bgr = cv2.applyColorMap(np.arange(256, dtype=np.uint8).reshape(256, 1), cv2.COLORMAP_JET).reshape(256, 3)
        _JET_RGB = torch.tensor(np.ascontiguousarray(bgr[:, ::-1]), dtype=torch.float32)
    return _JET_RGB.to(device)


def mask2cam(mask, imgs):
    """mask [n,1,h,w] (float64), imgs [n,3,h,w] -> (heatmap, cam) fp32 [n,3,h,w] (reference :234-251)."""
    if not imgs.is_cuda:
        raise ops.DgeError('mask2cam: dge_b200 runs on a B200 only; there is no CPU fallback')
    img = imgs.detach().float().contiguous()
    m = mask.detach().to(device=img.device, dtype=torch.float64).contiguous()
    n, _, h, w = img.shape
    heat = torch.empty_like(img)
    cam = torch.empty_like(img)
    scratch = torch.empty(4, dtype=torch.float32, device=img.device)
    ops.check(ops.lib().dge_mask2cam(ops._p(m), ops._p(img), ops._p(_jet_lut(img.device)), ops._p(heat), ops._p(cam),
                                     ops._p(scratch), n, h, w, ops._stream()))
    return heat, cam
