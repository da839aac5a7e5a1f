"""Time dge_conv_wgrad on the conv shapes of BE(startf=16, layer_count=9) at batch 8 (the StyleGAN2-1024 workload) and,
next to it, torch's fp32 weight gradient (cuDNN, TF32 off) of the same layer.  Not a bench line: context for DESIGN §7.
usage: python tools/probe_wgrad.py [planes=2] [iters=5]   (DGE_WGRAD_NO_STACK=1 disables the tap-stacked mode)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
import torch.nn.functional as F
from dge_b200 import ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
planes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N = int(os.environ.get("PROBE_N", "8"))
# (cin, cout, size, ksize): conv_1 / conv_2 of every BEBlock (model/E/E.py:27-36), 1024 -> 4
shapes = []
c, size = 16, 1024
while size >= 4:
    co = min(2 * c, 512)
    shapes += [(c, c, size, 3), (c, co, size, 3)]
    c, size = co, size // 2
sel = os.environ.get("PROBE_SEL")
if sel:
    shapes = [shapes[int(i)] for i in sel.split(",")]


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


rows, tot, tot_ref = [], 0.0, 0.0
for cin, cout, size, k in shapes:
    x = torch.randn(N, cin, size, size, device="cuda")
    dy = torch.randn(N, cout, size, size, device="cuda")
    xa, dya = ops.nchw_to_act(x, planes=planes), ops.nchw_to_act(dy, planes=planes)
    out = torch.empty(cout, cin, k, k, device="cuda")
    ms = timed(lambda: ops.conv_wgrad(dya, xa, k, out=out))
    wt = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)

    def ref():
        return torch.ops.aten.convolution_backward(dy, x, wt, None, [1, 1], [k // 2, k // 2], [1, 1], False, [0, 0], 1,
                                                   [False, True, False])[1]

    ms_ref = timed(ref)
    err = ((out - ref()).abs().max() / ref().abs().max()).item()
    fl = 2.0 * N * size * size * cin * cout * k * k
    gb = N * size * size * (cin + cout) * 2 * planes / 1e9
    rows.append({"cin": cin, "cout": cout, "size": size, "ms": round(ms, 4), "ms_torch_fp32": round(ms_ref, 4),
                 "tflops_alg": round(fl / ms / 1e9, 1), "gbps_alg": round(gb / ms * 1e3, 0), "rel_err": err})
    tot += ms
    tot_ref += ms_ref
    print(f"wgrad {cin:4d}->{cout:4d} @{size:4d}^2: {ms:8.3f} ms ({fl / ms / 1e9:7.1f} TFLOP/s alg, {gb / ms * 1e3:6.0f} GB/s alg)"
          f"   torch fp32 {ms_ref:8.3f} ms   rel err {err:.1e}", flush=True)
    del x, dy, xa, dya
print(f"total {tot:.3f} ms   torch fp32 total {tot_ref:.3f} ms")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"planes": planes, "batch": N, "rows": rows, "total_ms": tot, "total_ms_torch_fp32": tot_ref,
           "no_stack": bool(os.environ.get("DGE_WGRAD_NO_STACK"))},
          open(os.path.join(ROOT, "gpurun_out", "wgrad_probe%s.json" % ("_nostack" if os.environ.get("DGE_WGRAD_NO_STACK") else "")), "w"), indent=1)
