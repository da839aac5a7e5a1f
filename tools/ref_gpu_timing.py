"""Context measurement (not a bench arm): the UNMODIFIED reference modules on the B200 through PyTorch/cuDNN,
fp32 with TF32 off / on and bf16 autocast, same workload as bench.py.
Needs a copy of the reference under baseline/_ref (git-ignored) or $DGE_REF."""
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DGE_REF", os.path.join(ROOT, "baseline", "_ref"))
for n in ["matplotlib", "matplotlib.pyplot", "boto3", "botocore", "botocore.exceptions", "lpips", "tensorboardX"]:
    sys.modules.setdefault(n, types.ModuleType(n))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.modules["botocore.exceptions"].ClientError = Exception
sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
sys.path.insert(0, REF)
import torch
import model.stylegan2_generator as sg2
import model.E.E as E

torch.manual_seed(0)
dev = torch.device("cuda")
G = sg2.StyleGAN2Generator(1024).eval().to(dev)
Enc = E.BE(16, 512, 9, 512, 3).eval().to(dev)
out = {"cpu_count": os.cpu_count(), "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__,
       "cudnn": torch.backends.cudnn.version()}


def timeit(fn, warm=2, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


with torch.no_grad():
    z = torch.randn(8, 512, device=dev)
    imgs1 = G(z, trunc_psi=0.7, trunc_layers=8, randomize_noise=False)["image"]
    _, w = Enc(imgs1)
    for name, tf32, ac in (("fp32_tf32off", False, False), ("fp32_tf32on", True, False), ("bf16_autocast", True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if ac else torch.autocast("cuda", enabled=False)
        with ctx:
            tg = timeit(lambda: G.synthesis(w))
            te = timeit(lambda: Enc(imgs1))

            def both():
                c2, w2 = Enc(imgs1)
                return G.synthesis(w2)["image"]
            tb = timeit(both)
        out[name] = {"G_synthesis_ms": tg, "E_ms": te, "E_plus_G_ms": tb, "images_per_s": 8 / (tb / 1e3)}
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
print(json.dumps(out))
