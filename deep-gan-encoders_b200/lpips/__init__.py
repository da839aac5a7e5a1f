"""LPIPS (VGG16) perceptual distance -- stands in for the third-party `lpips` package the reference imports
(`lpips.LPIPS(net='vgg')`: E_align_s2.py:98, embedding_img.py:61, comparing-baseline.py:15; consumed by
`space_loss`, training_utils.py:93).  SURVEY 8(c): the package is an UNPINNED dependency (`requirements.txt:12`), is not
installed here and its weights cannot be fetched, so this module restates the published structure (Zhang et al. 2018,
"The Unreasonable Effectiveness of Deep Features as a Perceptual Metric", lpips v0.1, net='vgg'):

    x -> (x - shift) / scale                                   ScalingLayer, RGB constants below
      -> VGG16 relu1_2, relu2_2, relu3_3, relu4_3, relu5_3      torchvision `features[0:30]` split at the pools
      -> per tap: unit-normalise over channels (eps 1e-10), squared difference, 1x1 `lin` conv (C -> 1, no bias),
         spatial mean                                           -> summed over the five taps -> [N, 1, 1, 1]

with the package's parameter names (`net.slice{k}.{i}.weight|bias`, `lin{k}.model.1.weight`, `scaling_layer.shift|scale`) so
a state_dict saved from the real package loads.  PARITY UNPINNED: no reference vectors exist offline; the structure is
checked against `oracle/lpips.py` (a plain-torch restatement of the same published algorithm) with random weights.

Arithmetic: the twelve 3x3 convs with >= 64 input channels -- >99 % of the FLOPs -- run forward and, under autograd,
data-gradient (the VGG weights are frozen: `requires_grad=False` as in the package) on the tcgen05 kernels through
dge_b200.autograd.conv2d; the 3-channel first conv, ReLU, max-pool and the small reductions are torch CUDA ops in this
build.  CUDA-only (DgeError on CPU tensors), like every other module of this tree.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from dge_b200 import autograd as tc
from dge_b200 import ops

# torchvision vgg16.features: (index, in_channels, out_channels) of the convs of each LPIPS slice; a slice after the
# first starts with the 2x2 max-pool that precedes its first conv (features[4], [9], [16], [23])
_SLICES = [
    [(0, 3, 64), (2, 64, 64)],
    [(5, 64, 128), (7, 128, 128)],
    [(10, 128, 256), (12, 256, 256), (14, 256, 256)],
    [(17, 256, 512), (19, 512, 512), (21, 512, 512)],
    [(24, 512, 512), (26, 512, 512), (28, 512, 512)],
]
_CHNS = [64, 128, 256, 512, 512]


class ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer('shift', torch.Tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer('scale', torch.Tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, inp):
        return (inp - self.shift) / self.scale


class NetLinLayer(nn.Module):
    """A single 1x1 conv (C -> 1, no bias) behind the package's (inference-inert) dropout slot: `model.1.weight`."""

    def __init__(self, chn_in, chn_out=1, use_dropout=False):
        super().__init__()
        layers = [nn.Dropout()] if use_dropout else [nn.Identity()]
        layers += [nn.Conv2d(chn_in, chn_out, 1, stride=1, padding=0, bias=False)]
        self.model = nn.Sequential(*layers)

    def forward(self, x):
        return self.model(x)


class _VGG16Features(nn.Module):
    def __init__(self, requires_grad=False):
        super().__init__()
        for k, convs in enumerate(_SLICES):
            seq = nn.Sequential()
            for idx, cin, cout in convs:
                seq.add_module(str(idx), nn.Conv2d(cin, cout, 3, padding=1))
            setattr(self, f'slice{k + 1}', seq)
        self.planes = 2
        if not requires_grad:
            for p in self.parameters():
                p.requires_grad = False

    def forward(self, x):
        taps = []
        h = x
        for k in range(5):
            if k > 0:
                h = F.max_pool2d(h, 2, 2)
            for conv in getattr(self, f'slice{k + 1}'):
                w = conv.weight
                if w.shape[1] % 16 == 0 and w.shape[0] % 16 == 0:
                    h = tc.conv2d(h, w, self.planes) + conv.bias.view(1, -1, 1, 1)
                else:                                   # 3 input channels: point-wise cost, library conv
                    h = F.conv2d(h, w, conv.bias, padding=1)
                h = F.relu(h)
            taps.append(h)
        return taps


def normalize_tensor(x, eps=1e-10):
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


class LPIPS(nn.Module):
    def __init__(self, pretrained=True, net='vgg', version='0.1', lpips=True, spatial=False, pnet_rand=False,
                 pnet_tune=False, use_dropout=True, model_path=None, eval_mode=True, verbose=True):
        super().__init__()
        if net not in ('vgg', 'vgg16') or not lpips or spatial or version != '0.1':
            raise NotImplementedError("dge_b200 LPIPS: net='vgg', lpips=True, spatial=False, version='0.1' "
                                      "(the configuration the inversion scripts use)")
        self.pnet_type, self.pnet_tune, self.pnet_rand = net, pnet_tune, pnet_rand
        self.spatial, self.lpips, self.version = spatial, lpips, version
        self.scaling_layer = ScalingLayer()
        self.chns = _CHNS
        self.L = len(self.chns)
        self.net = _VGG16Features(requires_grad=pnet_tune)
        for k, c in enumerate(self.chns):
            setattr(self, f'lin{k}', NetLinLayer(c, use_dropout=use_dropout))
        self.lins = nn.ModuleList([getattr(self, f'lin{k}') for k in range(self.L)])
        if model_path is not None:
            self.load_state_dict(torch.load(model_path, map_location='cpu'), strict=False)
        elif pretrained and verbose:
            print('dge_b200 LPIPS: the pretrained VGG16 / linear-layer weights are not available offline; parameters are '
                  'randomly initialised -- pass model_path=... or load_state_dict() to use the published weights')
        if eval_mode:
            self.eval()

    def forward(self, in0, in1, retPerLayer=False, normalize=False):
        if not (in0.is_cuda and in1.is_cuda):
            raise ops.DgeError('LPIPS: dge_b200 runs on a B200 only; there is no CPU fallback')
        return self._distance(in0, in1, retPerLayer, normalize)

    def _distance(self, in0, in1, retPerLayer=False, normalize=False):
        tc.require_fp32_library_convs()
        if normalize:                                    # [0, 1] -> [-1, 1]
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        f0 = self.net(self.scaling_layer(in0.float()))
        f1 = self.net(self.scaling_layer(in1.float()))
        res = []
        for k in range(self.L):
            d = (normalize_tensor(f0[k]) - normalize_tensor(f1[k])) ** 2
            res.append(self.lins[k](d).mean(dim=(2, 3), keepdim=True))
        val = res[0]
        for k in range(1, self.L):
            val = val + res[k]
        return (val, res) if retPerLayer else val
