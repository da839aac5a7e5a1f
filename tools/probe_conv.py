"""Run one conv configuration a few times (for ncu captures / quick timing).
usage: python tools/probe_conv.py N H W CIN COUT [kind=0] [planes=2] [iters=5]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
from dge_b200 import ops

a = [int(v) for v in sys.argv[1:]]
n, h, w, cin, cout = a[:5]
kind = a[5] if len(a) > 5 else 0
planes = a[6] if len(a) > 6 else 2
iters = a[7] if len(a) > 7 else 5
x = ops.Act(n, cin, h, w, planes)
x.t.normal_()
wt = torch.randn(cout, cin, 3 if kind != 1 else 1, 3 if kind != 1 else 1, device="cuda") * 0.05
wpk = ops.pack_conv_weight(wt, flip=(kind == 2), planes=planes)
dm = torch.rand(n, cout, device="cuda") + 0.5
bias = torch.randn(cout, device="cuda")
noise = torch.randn(h, w, device="cuda")
s = torch.rand(n, cout, device="cuda") + 0.5


def run():
    if kind == 2:
        return ops.conv(x, wpk, cout, kind)
    return ops.conv(x, wpk, cout, kind, demod=dm, noise=noise, noise_scalar=0.3, bias=bias, slope=0.2, gain=1.414,
                    out_act=True, out_scale=s)


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
taps = 1 if kind == 1 else 9
fl = 2.0 * n * h * w * cin * cout * taps
print(f"conv kind={kind} n={n} {h}x{w} {cin}->{cout} planes={planes}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s algorithmic")
