"""Time dge_sg2_layer_bwd / dge_up_fir_bwd_s2d on the shapes of the StyleGAN2-1024 backward at batch 8 (context for DESIGN 4.5).
usage: python tools/probe_sg2_bwd.py [iters=10]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
from dge_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
N = 8


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# (channels, size, has ToRGB, has a next layer, x2 layer -> F32B output)
cases = [(32, 1024, True, False, False), (32, 1024, False, True, True), (64, 512, True, True, False),
         (64, 512, False, True, True), (128, 256, True, True, False), (512, 64, True, True, False)]
tot = 0.0
for c, s, rgb, nxt, up in cases:
    ya = ops.nchw_to_act(torch.randn(N, c, s, s, device="cuda"))
    dxs = ops.nchw_to_f32b(torch.randn(N, c, s, s, device="cuda")) if nxt else None
    scale = torch.rand(N, c, device="cuda") + 0.5 if nxt else None
    dimg = torch.randn(N, 3, s, s, device="cuda") if rgb else None
    rgbw = torch.randn(N, 3, c, device="cuda") if rgb else None
    noise = torch.randn(1, s, s, device="cuda")
    bias = torch.randn(c, device="cuda")
    demod = torch.rand(N, c, device="cuda") + 0.5
    ms = timed(lambda: ops.sg2_layer_bwd(ya, scale, dxs, dimg, rgbw, noise, False, 0.1, bias, demod, 2 ** 0.5, 0.2,
                                         out_kind="f32b" if up else "act"))
    gb = N * c * s * s * (4 + (4 if nxt else 0) + 4) / 1e9 + (N * 3 * s * s * 4 / 1e9 if rgb else 0)
    tot += ms
    print(f"sg2_layer_bwd c={c:4d} @{s:4d}^2 rgb={int(rgb)} dxs={int(nxt)} out={'f32b' if up else 'act '}: {ms * 1e3:8.1f} us "
          f"{gb / ms * 1e3:7.0f} GB/s alg")
    if up:
        d = ops.nchw_to_f32b(torch.randn(N, c, s, s, device="cuda"))
        ms2 = timed(lambda: ops.up_fir_bwd_s2d(d, 2))
        gb2 = N * c * s * s * 8 / 1e9
        tot += ms2
        print(f"up_fir_bwd_s2d c={c:4d} @{s:4d}^2: {ms2 * 1e3:8.1f} us {gb2 / ms2 * 1e3:7.0f} GB/s alg")
print(f"total {tot:.3f} ms")
