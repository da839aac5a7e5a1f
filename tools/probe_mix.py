"""Run the dominant launch classes of the bench step once each, with the epilogue options the models use
(for `ncu --set full` captures and quick CUDA-event timing).  usage: python tools/probe_mix.py [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
import torch
from dge_b200 import ops

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = 8
dev = "cuda"


def g_plain(h, c):
    x = ops.Act(n, c, h, h, 2)
    x.t.normal_()
    wpk = ops.pack_conv_weight(torch.randn(c, c, 3, 3, device=dev) * 0.05)
    dm, sc = torch.rand(n, c, device=dev) + 0.5, torch.rand(n, c, device=dev) + 0.5
    bias, noise = torch.randn(c, device=dev), torch.randn(h, h, device=dev)
    rgb_w = torch.randn(n, 3, c, device=dev)
    img = torch.zeros(n, 3, h, h, device=dev)
    return lambda: ops.conv(x, wpk, c, ops.CONV_3X3, demod=dm, noise=noise, noise_scalar=0.3, bias=bias, slope=0.2,
                            gain=1.414, out_act=True, out_scale=sc, rgb_w=rgb_w, rgb_out=img)


def g_up(h, cin, cout):
    x = ops.Act(n, cin, h, h, 2)
    x.t.normal_()
    wpk = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05, flip=True)
    dm, sc = torch.rand(n, cout, device=dev) + 0.5, torch.rand(n, cout, device=dev) + 0.5
    bias, noise = torch.randn(cout, device=dev), torch.randn(2 * h, 2 * h, device=dev)

    def run():
        raw = ops.conv(x, wpk, cout, ops.CONV_UP3X3)["raw_up"]
        return ops.up_fir_epilogue(raw, n, cout, 2 * h, 2 * h, demod=dm, noise=noise, noise_scalar=0.3, bias=bias,
                                   slope=0.2, gain=1.414, out_scale=sc)
    return run


def e_conv(h, cin, cout):
    x = ops.Act(n, cin, h, h, 2)
    x.t.normal_()
    wpk = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device=dev) * 0.05)
    nw, bias = torch.randn(cout, device=dev), torch.randn(cout, device=dev)
    noise = torch.randn(n, 1, h, h, device=dev)
    return lambda: ops.conv(x, wpk, cout, ops.CONV_3X3, noise=noise, noise_batched=True, noise_w=nw, bias=bias,
                            slope=0.2, out_f32b=True)


def e_stats_norm(h, c):
    x = ops.F32B(n, c, h, h)
    x.t.normal_()

    def run():
        st, mr = ops.instance_stats(x)
        return ops.instance_norm(x, mr)
    return run


def fir_only(h, c):
    raw = torch.randn(n, c // 8, 2 * h + 1, 2 * h + 1, 8, device=dev)
    dm, sc = torch.rand(n, c, device=dev) + 0.5, torch.rand(n, c, device=dev) + 0.5
    bias, noise = torch.randn(c, device=dev), torch.randn(2 * h, 2 * h, device=dev)
    return lambda: ops.up_fir_epilogue(raw, n, c, 2 * h, 2 * h, demod=dm, noise=noise, noise_scalar=0.3, bias=bias,
                                       slope=0.2, gain=1.414, out_scale=sc)


cases = [
    ("FIR 32ch -> 1024", fir_only(512, 32)),
    ("FIR 64ch -> 512", fir_only(256, 64)),
    ("FIR 512ch -> 64", fir_only(32, 512)),
    ("G conv3x3 32->32 @1024 (+ToRGB)", g_plain(1024, 32)),
    ("G conv3x3 64->64 @512 (+ToRGB)", g_plain(512, 64)),
    ("G conv3x3 128->128 @256 (+ToRGB)", g_plain(256, 128)),
    ("G conv3x3 256->256 @128 (+ToRGB)", g_plain(128, 256)),
    ("G conv3x3 512->512 @64 (+ToRGB)", g_plain(64, 512)),
    ("G up 64->32 @512->1024 (conv_up + fir)", g_up(512, 64, 32)),
    ("G up 256->128 @128->256 (conv_up + fir)", g_up(128, 256, 128)),
    ("E conv3x3 16->16 @1024", e_conv(1024, 16, 16)),
    ("E conv3x3 16->32 @1024", e_conv(1024, 16, 32)),
    ("E conv3x3 32->64 @512", e_conv(512, 32, 64)),
    ("E stats+IN 16 @1024", e_stats_norm(1024, 16)),
    ("small G conv3x3 512->512 @8", g_plain(8, 512)),
    ("small G conv3x3 512->512 @16", g_plain(16, 512)),
    ("small E conv3x3 512->512 @4", e_conv(4, 512, 512)),
]
sel = os.environ.get("PROBE_SEL")
for name, fn in cases:
    if sel and not any(k in name for k in sel.split(",")):
        continue
    fn()
    torch.cuda.synchronize()
    if os.environ.get("PROBE_ONCE"):
        continue
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"{name:45s} {ms:.3f} ms")
    L = ops.lib()
    if hasattr(L, "dge_exp_role_cycles"):      # experiment build only (DGE_LIB_PATH=tools/_exp/libdge_exp.so)
        import ctypes
        buf = (ctypes.c_uint64 * 16)()
        L.dge_exp_role_cycles(None, 1)
        fn()
        L.dge_exp_role_cycles(buf, 1)
        names = ["prod wait a_empty", "prod work", "mma wait tm_empty", "mma wait a_full", "mma issue",
                 "epi wait tm_full", "epi head", "epi tmem-ld wait", "epi groups", "epi tail"]
        tot = buf[0] + buf[1]
        print("    role cycles (sum over CTAs; % of the producer's total): " +
              ", ".join(f"{nm} {100.0 * buf[i] / max(tot, 1):.0f}%" for i, nm in enumerate(names)) +
              f"; cycles/CTA ~{tot / 296:.0f} (if 296 CTAs)")
