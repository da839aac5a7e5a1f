"""The VGG16 classifier of the Grad-CAM path (metric/grad_cam.py:101-232 as driven by E_mis_align_cropping_s1.py:99-190:
`GradCamPlusPlus(vgg16, 'features.28')`, `GuidedBackPropagation(vgg16)`) with its convolutional stack on this library's
kernels -- SURVEY 8f-4.  The caller's network stays the owner of the weights; when it has the torchvision VGG shape
(`features` = 3x3 stride-1 convs each followed by ReLU, 2x2 max-pools in between; `avgpool`; `classifier`) the 13 convs run
as dge_conv_forward with bias + ReLU in the epilogue (the same chain as the LPIPS node, dge_b200/train_lpips.py), the
fully-connected head runs through the caller's own modules (torch / cuBLAS, hooks included), and the backward walks the
stack with `relu_pool_bwd` (ReLU mask, arg-max routing, optional guided clamp) + tcgen05 data-gradient convs.

What the reference's hooks observe is reproduced exactly:
  * forward hook on conv `features.k`: its output tensor, which the following `ReLU(inplace=True)` has already overwritten
    when the CAM is computed -> the post-ReLU map;
  * legacy backward hook on the same conv: the gradient w.r.t. that output = (y > 0) * incoming gradient -- the operand
    `relu_pool_bwd` produces;
  * `GuidedBackPropagation` registers backward hooks on EVERY ReLU of the shared network, so once it has been constructed
    the Grad-CAM backward is clamped too (`guided` is read off the modules' hook tables).
"""
import torch
import torch.nn as nn

from . import ops


def supported(net):
    """True when `net` has the torchvision-VGG layout this runner implements."""
    f = getattr(net, 'features', None)
    if not isinstance(f, nn.Sequential) or not hasattr(net, 'avgpool') or not hasattr(net, 'classifier'):
        return False
    mods = list(f)
    i, convs = 0, 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Conv2d):
            ok = (m.kernel_size == (3, 3) and m.stride == (1, 1) and m.padding == (1, 1) and m.dilation == (1, 1)
                  and m.groups == 1 and m.bias is not None and m.out_channels % 16 == 0
                  and (m.in_channels % 16 == 0 or (convs == 0 and m.in_channels == 3)))
            if not ok or i + 1 >= len(mods) or not isinstance(mods[i + 1], nn.ReLU):
                return False
            convs += 1
            i += 2
        elif isinstance(m, nn.MaxPool2d):
            ks = m.kernel_size if isinstance(m.kernel_size, tuple) else (m.kernel_size, m.kernel_size)
            st = m.stride if isinstance(m.stride, tuple) else (m.stride, m.stride)
            if ks != (2, 2) or st != (2, 2) or m.padding not in (0, (0, 0)) or m.ceil_mode or convs == 0:
                return False
            i += 1
        else:
            return False
    return convs > 0 and isinstance(mods[-1], nn.MaxPool2d)


class FusedVGG:
    def __init__(self, net, planes=2):
        self.net, self.planes = net, planes
        mods = list(net.features)
        self.convs, self.names, self.pool_after = [], [], []
        for i, m in enumerate(mods):
            if isinstance(m, nn.Conv2d):
                self.convs.append(m)
                self.names.append(f'features.{i}')
                self.pool_after.append(i + 2 < len(mods) and isinstance(mods[i + 2], nn.MaxPool2d))
        self.relus = [m for m in mods if isinstance(m, nn.ReLU)]
        self._ops_key, self._ops = None, None

    def _operands(self):
        key = ops.weight_key(*[c.weight for c in self.convs], *[c.bias for c in self.convs])
        if key != self._ops_key:
            fwd, bwd, bias = [], [], []
            for i, c in enumerate(self.convs):
                w = c.weight.detach().float()
                if i == 0 and w.shape[1] == 3:
                    w = torch.cat((w, w.new_zeros(w.shape[0], 13, 3, 3)), dim=1).contiguous()
                fwd.append(ops.pack_conv_weight(w, planes=self.planes))
                bwd.append(ops.pack_conv_weight_dgrad(w, planes=self.planes))
                bias.append(c.bias.detach().float().contiguous())
            self._ops_key, self._ops = key, (fwd, bwd, bias)
        return self._ops

    @property
    def guided(self):
        """GuidedBackPropagation has hooked the ReLUs of this network (grad_cam.py:199-203)."""
        return any(len(m._backward_hooks) > 0 for m in self.relus)

    def forward(self, x):
        """-> logits [N, classes] (a torch tensor whose graph ends at the pooled feature map `self.p5`)."""
        fwd, _, bias = self._operands()
        x = x.detach().float().contiguous()
        if x.shape[1] == 3:
            act = ops.lpips_input(x, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0), self.planes)
        else:
            act = ops.nchw_to_act(x, planes=self.planes)
        self.saved = []
        last = len(self.convs) - 1
        for i, c in enumerate(self.convs):
            pool = self.pool_after[i]
            r = ops.conv(act, fwd[i], c.out_channels, ops.CONV_3X3, bias=bias[i], slope=0.0, out_act=not pool,
                         out_f32b=True)
            y = r['f32b']                      # every activation in fp32: any conv may be the hooked layer
            self.saved.append(y)
            if pool and i != last:
                act = ops.maxpool_to_act(y, self.planes)
            elif not pool:
                act = r['act']
        self.p5 = ops.maxpool2(self.saved[last]).to_nchw().requires_grad_(True)
        self.in_hw = x.shape[2:]
        out = self.net.avgpool(self.p5)
        return self.net.classifier(torch.flatten(out, 1))

    def feature(self, layer_name):
        return self.saved[self.names.index(layer_name)].to_nchw()

    def backward(self, stop_at=None):
        """After `target.backward()` filled `self.p5.grad`: walk the conv stack down.  stop_at = a layer name -> the gradient
        the legacy backward hook of that conv reports (NCHW);  None -> the gradient w.r.t. the input image [N, 3, H, W]."""
        _, bwd, _ = self._operands()
        guided = self.guided
        g = ops.nchw_to_f32b(self.p5.grad.contiguous().float())
        stop = None if stop_at is None else self.names.index(stop_at)
        for i in range(len(self.convs) - 1, -1, -1):
            y = self.saved[i]
            d = ops.relu_pool_bwd(y, g_same=None if self.pool_after[i] else g, g_pool=g if self.pool_after[i] else None,
                                  planes=self.planes, clamp_pos=guided)
            if stop is not None and i == stop:
                return d.to_nchw()
            cin = self.convs[i].in_channels
            g = ops.conv(d, bwd[i], 16 if cin == 3 else cin, ops.CONV_3X3, out_f32b=True)['f32b']
        return g.to_nchw()[:, :self.convs[0].in_channels].contiguous()
