#!/usr/bin/env python
"""bench.py -- E+G forward images/s on StyleGAN2-FFHQ-1024 (BASELINE.json metric / configs[2]).

One "step" = the two forwards every iteration of the reference training loop performs
(E_align_s2.py:153,160) on one batch of 8 synthetic 1024x1024 images per GPU:
    const2, w2 = E(imgs1);  imgs2 = G.synthesis(w2)['image']
with E = BE(startf=16, maxf=512, layer_count=9) and G = StyleGAN2Generator(1024), random-init
weights (no checkpoints offline), fp32-equivalent split-precision bf16x3 tensor-core math.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
N > 1 is launched by the driver through torch.distributed.run (one rank per GPU); the forward path shards
by sample with no data-path collective (forward-only => replicas), so scaling is "weak".
Prints ONE JSON line on rank 0.

Beside the headline the line carries
  train          the whole E_align_s2.py training iteration (:140-221) through the drop-in modules at every N: one
                 replica per GPU, encoder gradients exchanged by the bucketed NCCL all-reduce that overlaps the backward
                 (dge_b200.dist.GradBucket, hooked into LREQAdam) -- encoder-train images/s, all-reduce and exposed-comm ms
  reference_gpu  (N = 1) the UNMODIFIED reference on the same GPU in the same run: forward and training iteration
                 (tools/ref_gpu_timing.py as a subprocess)
  inversion      (N = 1) BASELINE configs[4]: the embedding_img.py inversion loop with StyleGAN2-1024 (tools/bench_invert.py),
                 ours and the unmodified reference on this GPU, s/image and final reconstruction MSE
  cpu_baseline   (N = 1) the reference's CPU path on this box's host cores (`--impl reference` as a subprocess)
`--impl reference` times the unmodified reference modules (baseline/_ref, a git-ignored copy that travels with the
snapshot) on the host cores at the same batch 8; without that copy it falls back to the oracle port and says so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "deep-gan-encoders_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "E+G fwd images/sec (StyleGAN2-FFHQ1024, bs=8)"
UNIT = "images/s"
RES, BATCH, STARTF, LAYERS = 1024, 8, 16, 9
# algorithmic dense-conv work per image (2*MAC): G synthesis 148.5 GFLOP + E 86.7 GFLOP
GFLOP_PER_IMAGE = 235.2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        # keep stdout to the ONE JSON line: NCCL writes its version banner / debug log to stdout unless redirected
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    return rank, world, local


# -------------------------------------------------------------------------------------------------
# synthetic weights / inputs
# -------------------------------------------------------------------------------------------------
def perturb_zero_init(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in list(module.named_parameters()) + list(module.named_buffers()):
            if k.endswith(("bias", "noise_strength", "w_avg", "noise_weight_1", "noise_weight_2", "bias_1", "bias_2")):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)


def build_ours(device, res=RES, startf=STARTF, layers=LAYERS):
    from model.E.E import BE
    from model.stylegan2_generator import StyleGAN2Generator
    torch.manual_seed(0)
    G = StyleGAN2Generator(res).eval()
    E = BE(startf, 512, layers, 512, 3).eval()
    perturb_zero_init(G, 1)
    perturb_zero_init(E, 2)
    E.set_noise_mode("device")
    return G.to(device), E.to(device)


def oracle_state(res=RES, startf=STARTF, layers=LAYERS):
    """Same seeded random-init weights as build_ours, as plain CPU state dicts for the oracle."""
    G, E = build_ours("cpu", res, startf, layers)
    return ({k: v.detach().float() for k, v in G.state_dict().items()},
            {k: v.detach().float() for k, v in E.state_dict().items()})


def oracle_step(gsd, esd, imgs1, res=RES, layers=LAYERS):
    from oracle import encoder as oenc
    from oracle import stylegan2 as osg2
    const2, w2 = oenc.be_forward(esd, imgs1, layers)
    return osg2.synthesis(gsd, w2, res)["image"], const2, w2


REF_DIR = os.environ.get("DGE_REF", os.path.join(ROOT, "baseline", "_ref"))


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "model", "stylegan2_generator.py"))


def time_cpu_reference(steps, warmup, batch=BATCH):
    """The reference's own CPU path: its unmodified modules (baseline/_ref) when the copy is there (kind "reference"), else
    the oracle restatement (kind "port").  Must run in a process that has NOT imported the drop-in package (same module
    names)."""
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(3)
    imgs1 = torch.randn(batch, 3, RES, RES, generator=g).clamp_(-1, 1)
    if reference_available():
        import types
        for n in ["matplotlib", "matplotlib.pyplot", "boto3", "botocore", "botocore.exceptions", "lpips", "tensorboardX"]:
            sys.modules.setdefault(n, types.ModuleType(n))
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.modules["botocore.exceptions"].ClientError = Exception
        sys.modules["botocore"].exceptions = sys.modules["botocore.exceptions"]
        pkg = os.path.join(ROOT, "deep-gan-encoders_b200")
        sys.path[:] = [REF_DIR] + [p for p in sys.path if os.path.abspath(p or ".") != pkg]
        import model.E.E as ref_e
        import model.stylegan2_generator as ref_g
        assert os.path.abspath(ref_g.__file__).startswith(os.path.abspath(REF_DIR))
        torch.manual_seed(0)
        G = ref_g.StyleGAN2Generator(RES).eval()
        E = ref_e.BE(STARTF, 512, LAYERS, 512, 3).eval()

        def step():
            const2, w2 = E(imgs1)
            return G.synthesis(w2)["image"]
        kind = "reference"
    else:
        gsd, esd = oracle_state()

        def step():
            return oracle_step(gsd, esd, imgs1)
        kind = "port"
    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, kind


# -------------------------------------------------------------------------------------------------
# arms
# -------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, _ = dist_setup(args.gpus)
    if rank != 0:
        return
    ips, dt, kind = time_cpu_reference(args.steps, args.warmup, batch=BATCH)
    cores = os.cpu_count()
    src = ("the UNMODIFIED reference modules (baseline/_ref: model/stylegan2_generator.py, model/E/E.py), PyTorch fp32 on the "
           "host cores" if kind == "reference" else
           "oracle/ torch-fp32 restatement of the reference forward (no copy of the reference under baseline/_ref)")
    sample = f"the full step at the benchmark batch of {BATCH} ({dt:.1f} s per step), all {cores} host threads; {src}"
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: StyleGAN2-FFHQ-1024 synthesis + BE(startf=16, L=9) encoder forward, "
                                   "batch 8, random-init weights, CPU", "global_batch": BATCH,
                       "note": "one process on the host cores whatever --gpus says (rank 0 only under torchrun)"},
            "cpu_baseline": {"value": ips, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": ips, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (ours) needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    from dge_b200 import ops
    dev = torch.device("cuda", local if world > 1 else 0)
    torch.cuda.set_device(dev)
    ops.lib()
    G, E = build_ours(dev)
    pk = peaks()

    with torch.no_grad():
        from dge_b200 import dist as ddist
        # the GLOBAL latent batch is drawn from one seed and sliced by rank (SURVEY 8e caveat 3)
        z = ddist.global_latents(100, BATCH * world, 512, rank, world, device=dev)
        imgs1 = G(z, trunc_psi=0.7, trunc_layers=8, randomize_noise=False)["image"].contiguous()
        imgs1_host = imgs1.cpu().pin_memory()
        # w2 [8,18,512] and const2 [8,512,4,4]: two CONTIGUOUS pinned buffers (a strided slice of one pinned tensor makes
        # copy_ stage through pageable memory, i.e. a blocking D2H that serialises the host with every step)
        w2_host = torch.empty((BATCH, 18, 512), dtype=torch.float32).pin_memory()
        const2_host = torch.empty((BATCH, 512, 4, 4), dtype=torch.float32).pin_memory()

        def step(x):
            const2, w2 = E(x)
            return G.synthesis(w2)["image"], const2, w2

        def barrier():
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            torch.cuda.synchronize()

        def timed(fn, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            return ddist.max_over_ranks(e0.elapsed_time(e1), dev) / steps

        # ---- device-resident throughput ----------------------------------------------------------
        for _ in range(max(args.warmup, 3)):
            step(imgs1)
        ops.launch_count_reset()
        step(imgs1)
        launches_per_step = ops.launch_count()
        # The ~190 launches of one step are captured once into a CUDA graph (streams and graphs instead of a tracing
        # compiler); every timed step replays it on a static input buffer.
        static_in = imgs1.clone()
        graph, graph_out = None, None
        if not args.no_graph:
            try:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    step(static_in)
                torch.cuda.current_stream().wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    graph_out = step(static_in)
            except Exception as exc:  # noqa: BLE001 -- report and fall back to eager launches
                print(f"bench: CUDA graph capture failed ({exc!r}); timing eager launches", file=sys.stderr)
                graph = None
                torch.cuda.synchronize()

        def run_step():
            if graph is not None:
                graph.replay()
                return graph_out
            return step(static_in)

        for _ in range(3):
            run_step()
        sampler = ClockSampler(dev.index or 0)
        sampler.start()
        ms = timed(run_step, args.steps)
        clocks = sampler.stop()
        launches = launches_per_step * args.steps

        # ---- end to end: pinned host images in, latents + reconstruction MSE out, every step ---------
        # Every step copies ITS input batch host->device and its results device->host inside the timed region.  The
        # H2D of step i+1 (100 MB over PCIe, ~2 ms) runs on a copy stream into a staging buffer while step i computes;
        # the compute stream picks it up with a device-to-device copy.  Nothing is skipped: all copies are bracketed
        # by the same barrier + synchronize as the kernels.
        mom = torch.zeros(6, dtype=torch.float64, device=dev)
        mom_host = torch.empty(6, dtype=torch.float64).pin_memory()
        copy_stream = torch.cuda.Stream()
        staging = [torch.empty_like(static_in), torch.empty_like(static_in)]
        staged = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        main = torch.cuda.current_stream()
        state = {"i": 0}

        def prefetch(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                staging[slot].copy_(imgs1_host, non_blocking=True)
                staged[slot].record(copy_stream)

        def e2e_step():
            slot = state["i"] & 1
            state["i"] += 1
            prefetch(slot ^ 1)                      # next step's input, overlapped with this step's kernels
            main.wait_event(staged[slot])
            static_in.copy_(staging[slot], non_blocking=True)
            consumed[slot].record(main)
            img2, const2, w2 = run_step()
            w2_host.copy_(w2, non_blocking=True)
            const2_host.copy_(const2, non_blocking=True)
            # reconstruction MSE: one pass of the fused moments kernel (dge_pair_moments), 6 doubles back to the host
            ops.check(ops.lib().dge_pair_moments(ops._p(img2), ops._p(static_in), img2.numel(), ops._p(mom),
                                                 ops._stream()))
            mom_host.copy_(mom, non_blocking=True)

        for ev in consumed:
            ev.record(main)
        prefetch(0)
        for _ in range(2):
            e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
        copy_stream.synchronize()
        h2d = imgs1_host.numel() * 4
        d2h = (w2_host.numel() + const2_host.numel()) * 4 + mom_host.numel() * 8

        # ---- per-kernel roofline: CUDA events around every launch of one more step ------------------
        with ops.profile() as rec:
            for _ in range(3):
                step(imgs1)
        prof = rec.summary()

    roof, roof_named, top = roofline_from_profile(prof, pk)
    ips = BATCH * world / (ms / 1e3)
    ips_e2e = BATCH * world / (ms_e2e / 1e3)
    line = {
        "metric": METRIC, "value": ips, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16x3 (split bf16 hi+lo operands, fp32 accumulate: fp32-equivalent, meets the 1e-3 parity bar)",
        "data": "synthetic",
        "config": {"workload": "configs[2]: StyleGAN2-FFHQ-1024 synthesis + BE(startf=16, L=9) encoder forward, "
                               "batch 8 per GPU, random-init weights, encoder noise drawn on device",
                   "global_batch": BATCH * world, "parallelism": f"replicas x{world} (no data-path collective)",
                   "l2": "per-step working set ~25 GB of activations >> 126 MB L2; no explicit flush",
                   "launch": "CUDA graph replay of the step" if graph is not None else "eager launches"},
        "e2e": {"value": ips_e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "what": "pinned host imgs1 -> H2D (copy stream, overlapped with the previous step's kernels) -> E -> "
                        "G.synthesis -> D2H of (w2, const2) + the 6 recon moments (sum (a-b)^2 = MSE * numel)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roof,
        "roofline_modconv_128_256": roof_named,
        "achieved_conv_tflops": GFLOP_PER_IMAGE * ips / world / 1e3,
        "top_kernels": top,
        "peaks": pk,
    }
    # ---- the training iteration (SURVEY 8f-1 / north star: encoder-train images/s at 1, 2, 4, 8 GPUs) ----------
    if not args.no_train:
        del graph, graph_out
        torch.cuda.empty_cache()
        line["train"] = train_leg(G, E, dev, rank, world, max(3, min(args.steps, 10)), 3)
    if rank == 0:
        if world == 1 and not args.no_reference_gpu:
            torch.cuda.empty_cache()
            line["reference_gpu"] = run_json_subprocess(
                [sys.executable, os.path.join(ROOT, "tools", "ref_gpu_timing.py")] + (["--no-train"] if args.no_train else []),
                600)
        if world == 1 and not args.no_inversion:
            # BASELINE configs[4]: the embedding_img.py inversion loop (batch 1), ours and the unmodified reference on this GPU
            torch.cuda.empty_cache()
            inv = [sys.executable, os.path.join(ROOT, "tools", "bench_invert.py"), "--images", "3", "--iterations", "6"]
            ours_inv = run_json_subprocess(inv, 600)                                     # encoder noise drawn on the device
            ours_same = run_json_subprocess(inv + ["--noise", "reference"], 600)          # the reference's CPU noise stream
            ref_inv = run_json_subprocess(inv + ["--impl", "reference"], 900)
            line["inversion"] = {"ours": ours_inv, "ours_reference_noise": ours_same, "reference_gpu": ref_inv,
                                 "note": "timing: `ours` (device noise) vs `reference_gpu`; reconstruction MSE: "
                                         "`ours_reference_noise` vs `reference_gpu` (identical seeds and noise stream)"}
            if "ms_per_iteration" in ours_inv and "ms_per_iteration" in ref_inv:
                line["inversion"]["speedup_vs_reference_gpu"] = ref_inv["ms_per_iteration"] / ours_inv["ms_per_iteration"]
            # opt-in switch DGE_TRAIN_GRAPHS=1: the synthesis node replays CUDA graphs (dge_b200/train_g.py)
            ours_g = run_json_subprocess(inv + ["--graphs"], 600)
            line["inversion"]["ours_graphs"] = ours_g
            if "ms_per_iteration" in ours_g and "ms_per_iteration" in ref_inv:
                line["inversion"]["speedup_vs_reference_gpu_graphs"] = ref_inv["ms_per_iteration"] / ours_g["ms_per_iteration"]
            # the same loop with the case-2 encoder, the class embedding_img.py:9 actually imports (model/E/E_Blur.py)
            blur = [sys.executable, os.path.join(ROOT, "tools", "bench_invert.py"), "--encoder", "blur", "--images", "3",
                    "--iterations", "6"]
            ours_b, ref_b = run_json_subprocess(blur, 600), run_json_subprocess(blur + ["--impl", "reference"], 900)
            line["inversion"]["e_blur"] = {"ours": ours_b, "reference_gpu": ref_b}
            if "ms_per_iteration" in ours_b and "ms_per_iteration" in ref_b:
                line["inversion"]["e_blur"]["speedup_vs_reference_gpu"] = ref_b["ms_per_iteration"] / ours_b["ms_per_iteration"]
        if world == 1 and not args.no_cpu_baseline:
            ref = run_json_subprocess([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3",
                                       "--warmup", "1"], 900)
            line["cpu_baseline"] = ref.get("cpu_baseline", ref)
        print(json.dumps(line), flush=True)


def run_json_subprocess(cmd, timeout):
    """Run a helper in its own process (the reference's module names clash with the drop-in's) -> its last JSON line."""
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout,
                           env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (r.stderr or r.stdout)[-400:]}
    except Exception as exc:  # noqa: BLE001
        return {"error": repr(exc)[:400]}


def train_leg(G, E, dev, rank, world, steps, warmup):
    """E_align_s2.py:140-221, mtype 2, through the drop-in modules: per iteration G(z) under no_grad, E(imgs1),
    G.synthesis(w2), the three image-space `space_loss` calls (full image, AT1 and AT2 crops) + the latent one, and the two
    `zero_grad / backward / step` pairs.  One replica per GPU on its own slice of the seeded global z batch; LREQAdam owns
    the gradient exchange (bucketed all-reduce overlapped with the backward)."""
    import lpips
    import torch.distributed as dist
    import training_utils as tu
    from dge_b200 import dist as ddist
    from dge_b200 import ops
    from model.utils.custom_adam import LREQAdam
    E.train()
    opt = LREQAdam([{"params": E.parameters()}], lr=0.0015, betas=(0.0, 0.99), weight_decay=0)
    lp = lpips.LPIPS(net="vgg", pretrained=False, pnet_rand=True, verbose=False).to(dev)
    kw = dict(trunc_psi=0.7, trunc_layers=8, randomize_noise=False)

    def iteration(it):
        z = ddist.global_latents(it % 30000, BATCH * world, 512, rank, world, device=dev)     # set_seed(iteration % 30000)
        with torch.no_grad():
            r = G(z, **kw)
            imgs1, w1 = r["image"], r["wp"]
        const2, w2 = E(imgs1)
        imgs2 = G.synthesis(w2)["image"]
        opt.zero_grad()
        l0, _ = tu.space_loss(imgs1, imgs2, lpips_model=lp)
        m = imgs1.shape[3] // 8
        l1, _ = tu.space_loss(imgs1[:, :, :, m:-m], imgs2[:, :, :, m:-m], lpips_model=lp)
        m2 = m + imgs1.shape[2] // 32
        l2, _ = tu.space_loss(imgs1[:, :, m2:-m2, m2:-m2], imgs2[:, :, m2:-m2, m2:-m2], lpips_model=lp)
        opt.zero_grad()
        (l0 + l1 * 5 + l2 * 9).backward(retain_graph=True)
        opt.step()
        lw, _ = tu.space_loss(w1, w2, image_space=False)
        opt.zero_grad()
        (lw * 0.01).backward()
        opt.step()

    t_w = torch.randn(BATCH, 18, 512, device=dev)

    def encoder_step(it, imgs):
        const, w = E(imgs)
        loss = ((w - t_w) ** 2).mean() + (const ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        return ddist.max_over_ranks(e0.elapsed_time(e1), dev) / n

    for i in range(warmup):
        iteration(i)
    ops.launch_count_reset()
    iteration(warmup)
    launches = ops.launch_count()
    torch.cuda.reset_peak_memory_stats()
    ms = timed(iteration, steps)
    peak = torch.cuda.max_memory_allocated() / 2 ** 30
    with torch.no_grad():
        imgs = G(ddist.global_latents(7, BATCH * world, 512, rank, world, device=dev), **kw)["image"]
    for i in range(2):
        encoder_step(i, imgs)
    ms_enc = timed(lambda i: encoder_step(i, imgs), steps)
    out = {"metric": "encoder-train images/sec (E_align_s2.py iteration: StyleGAN2-FFHQ1024 + BE(16,9), bs=8/GPU)",
           "value": BATCH * world / (ms / 1e3), "unit": UNIT, "n_gpus": world, "ms_per_iteration": ms, "steps": steps,
           "warmup": warmup + 1, "scaling": "weak", "peak_gib": peak, "gpu_launches_per_iteration": launches,
           "what": "G(z) no_grad + E fwd + G.synthesis fwd + 3 image-space space_loss (MSE, cosine, SSIM, LPIPS-VGG16 with "
                   "random weights) + latent space_loss + 2 x (zero_grad, backward, gradient all-reduce, LREQAdam.step); "
                   "fused training nodes (dge_b200/train_e.py, train_g.py), bf16x3 split-precision convs",
           "encoder_step": {"ms": ms_enc, "images_per_s": BATCH * world / (ms_enc / 1e3),
                            "what": "BE(16,9) forward + backward + gradient all-reduce + LREQAdam.step only"}}
    # data-parallel correctness, not just speed: every rank started from the same weights, saw different images and applied the
    # averaged gradients -> after all those steps the replicas must still be bit-identical (and finite)
    chk = torch.stack([p.detach().double().sum() for p in E.parameters()])
    out["train_state"] = {"parameters_finite": bool(torch.isfinite(chk).all())}
    if world > 1:
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        out["train_state"]["replicas_bit_identical"] = bool((hi == lo).all())
    bucket = opt._bucket
    if world > 1 and bucket is not None:
        # the same iterations without the exchange -> exposed communication; and the exchange alone
        bucket.enabled = False
        ms_nocomm = timed(iteration, steps)
        bucket.enabled = True
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.all_reduce(bucket.flat)
        e0.record()
        for _ in range(5):
            dist.all_reduce(bucket.flat)
        e1.record()
        barrier()
        ar = ddist.max_over_ranks(e0.elapsed_time(e1), dev) / 5
        out.update({"allreduce_bytes_per_iteration": 2 * bucket.bytes_per_exchange, "allreduce_ms_standalone": ar,
                    "allreduce_busbw_gbs": 2 * (world - 1) / world * bucket.bytes_per_exchange / (ar / 1e3) / 1e9,
                    "ms_per_iteration_without_exchange": ms_nocomm, "exposed_comm_ms_per_iteration": ms - ms_nocomm,
                    "chunks": len(bucket.chunks), "exchange": "2 per iteration (one per backward), NCCL AVG all-reduce of "
                    "the flat fp32 gradient bucket in 4 chunks launched from post-accumulate hooks during the backward"})
    else:
        out.update({"allreduce_bytes_per_iteration": 0, "exposed_comm_ms_per_iteration": 0.0})
    E.eval()
    return out


def conv_flops(key, name):
    n, h, w, cin, cout = key[:5]
    taps = 1 if name == "conv1x1" else 9
    return 2.0 * n * h * w * cin * cout * taps     # conv_up3x3: transposed conv over the INPUT grid


def conv_bytes(key, name):
    """compulsory HBM bytes: read the ACT operand once, write the output once (4 B/element each in bf16x3 mode)."""
    n, h, w, cin, cout, planes = key
    out_px = (2 * h + 1) * (2 * w + 1) if name == "conv_up3x3" else h * w
    return n * (h * w * cin * 2 * planes + out_px * cout * 4)


def ncu_traffic(name, key):
    """DRAM bytes per launch of this launch class from the committed ncu capture (None if it was not captured)."""
    path = os.path.join(ROOT, "profiles", "r1d_top_kernels_ncu.json")
    if not os.path.exists(path) or name != "conv3x3":
        return None
    n, h, w, cin, cout = key[:5]
    want = f"conv3x3 {cin}->{cout} @{h}^2"
    try:
        for k, d in json.load(open(path)).items():
            if k.startswith(want):
                tot = 0.0
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    v, u = d[m]
                    tot += float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
                return tot
    except Exception:
        return None
    return None


def conv_roofline(d, pk):
    """Roofline entry of one conv launch class: bound = the larger of the x3-MMA tensor time and the HBM time."""
    key, name = tuple(d["key"]), d["kernel"]
    fl, by = conv_flops(key, name), conv_bytes(key, name)
    sec = d["ms_per_launch"] / 1e3
    t_tc = fl * 3 / (pk["bf16_tflops_sustained"] * 1e12)      # bf16x3 issues 3 MMAs per algorithmic MAC
    t_hbm = by / (pk["hbm_gbs"] * 1e9)
    if t_tc >= t_hbm:
        roof = {"bound": "tensor", "achieved": fl / sec / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": fl / sec / 1e12 / pk["bf16_tflops_sustained"],
                "mma_tflops": 3 * fl / sec / 1e12,
                "note": "achieved = ALGORITHMIC conv FLOPs (2*MAC) over the CUDA-event launch time; the parity mode "
                        "issues 3 bf16 MMAs per MAC (mma_tflops = 3 x achieved is the tensor-core work actually done: "
                        "compare THAT with the peak; frac = achieved/peak is <= 1/3 of whatever peak the clocks allow)"}
    else:
        roof = {"bound": "hbm", "achieved": by / sec / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": by / sec / 1e9 / pk["hbm_gbs"],
                "note": "algorithmic bytes = ACT operand read once + output written once (4 B/element each)"}
    roof["traffic"] = ncu_traffic(name, key)
    roof["traffic_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum of one launch of this shape from the committed "
                            "ncu --set full capture (profiles/r1d_top_kernels_ncu.json)") if roof["traffic"] else None
    roof["kernel"] = f"{name} n,h,w,cin,cout,planes={list(key)}"
    roof["ms_per_launch"] = d["ms_per_launch"]
    roof["peak_source"] = pk["source"] + " (sustained bf16 / copy bandwidth, MEASURED_PEAKS.json)"
    return roof


def roofline_from_profile(prof, pk):
    rows = []
    for (name, key), (cnt, ms) in prof.items():
        rows.append({"kernel": name, "key": list(key), "launches": cnt, "ms_per_launch": ms / cnt, "ms_total": ms})
    total = sum(r["ms_total"] for r in rows) or 1.0
    rows.sort(key=lambda r: -r["ms_total"])
    for r in rows:
        r["share"] = r["ms_total"] / total
    top = rows[:12]
    conv = [r for r in rows if r["kernel"].startswith("conv")]
    if not conv:
        return None, None, top
    roof = conv_roofline(conv[0], pk)                       # the launch class with the largest share of the step
    # ... and the StyleGAN2-1024 modulated conv the north star names (128 -> 128 channels at 256^2, tensor-bound)
    named = [r for r in conv if r["kernel"] == "conv3x3" and r["key"][1:5] == [256, 256, 128, 128]]
    roof_named = conv_roofline(named[0], pk) if named else None
    return roof, roof_named, top


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-iteration leg")
    ap.add_argument("--no-reference-gpu", action="store_true", help="skip the unmodified reference on this GPU")
    ap.add_argument("--no-inversion", action="store_true", help="skip the configs[4] inversion-loop leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager launches instead of a CUDA-graph replay")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
