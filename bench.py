#!/usr/bin/env python
"""Headline benchmark: 512x512 txt2img images/s (50-step DDPM, CFG 7.5) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--config 1|3|4] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: CLIP x2 -> 50 x (UNet with CFG pair + fused
CFG/DDPM step, replayed from one CUDA graph) -> VAE decode -> uint8, for B images per GPU
(BASELINE.json configs[1]: B = 8 on one B200). Multi-GPU = independent seeds per rank, no data-path
collective. --config 1 (default): weak scaling, B per GPU fixed; --config 3 / --strong: BASELINE.json configs[3],
64 images split across the ranks; --config 4: 768x768, B = 8 per GPU.

  value     images/s with every input resident in HBM when the timed region starts
  e2e       the same through pipeline.generate() (prompt strings in, host uint8 images out)
  roofline  the dominant variant of the dominant kernel (implicit-GEMM 3x3 conv, gemm_tc_kernel): algorithmic FLOPs
            of THAT launch variant / its duration re-timed inside a CUDA graph, against the measured bf16 peak of
            MEASURED_PEAKS.json; roofline_attention / roofline_hbm / detail.rooflines: the same for the flash
            attention kernel and the HBM-bound kernels (projection GEMM with fp32 residual, GroupNorm, LayerNorm)
  cpu_baseline  the reference's own pipeline.generate(device="cpu") (oracle/_ref: the unmodified reference compiled
            to bytecode by oracle/build_ref.py) on the host cores, on a bounded sample

--impl reference times that CPU path alone, K + W real generate() calls with --ref-steps denoising steps each
(--full: 50), the UNet evaluations beyond --ref-steps extrapolated from the measured ones.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "512x512 txt2img images/s (50-step DDPM, CFG 7.5)"
UNIT = "images/s"
N_STEPS = 50
CFG = 7.5
H = W = 512
# Algorithmic work per image (CFG pair per UNet evaluation, dead GEGLU gate and hoisted cross-attention
# K/V excluded). "fold": linear_geglu_2 . linear_geglu_1[:4C] composed into one C x C map at pack time - the saved
# GEMM FLOPs leave the numerator, as §8d prescribes (token-proportional: x2.25 at 96x96 latents). "upfold": the three
# Upsample convs (1280 @ 16x16 and 32x32, 640 @ 64x64: 135.9 GFLOP per image-step as the reference executes them) run
# as four 2x2 phase convolutions of the low-resolution input, 4/9 of the FLOPs: 75.5 GFLOP leave the numerator too, and
# so do 386.5 GFLOP per image of the VAE decoder's three Upsample -> conv pairs (695.8 GFLOP as the reference runs them).
# "cfg_prefix": the two members of a classifier-free-guidance pair share the latent and the time step, so everything
# before the first cross-attention (stem conv, first resblock, first self-attention with its projections) is evaluated
# once per pair: 40.86 GFLOP per image-step (0.09 + 15.10 + 0.84 + 2.52 + 21.47 + 0.84) are no longer executed.
WORKLOADS = {
    1: {"hw": 512, "unet_gflop": 1498.25, "fold_gflop": 179.1, "upfold_gflop": 75.5, "cfg_prefix_gflop": 40.86, "vae_gflop": 2514.52,
        "vae_upfold_gflop": 386.5,
        "name": "configs[1]: SD1.5-arch random-init txt2img 512x512 (4x64x64 latent)"},
    3: {"hw": 512, "unet_gflop": 1498.25, "fold_gflop": 179.1, "upfold_gflop": 75.5, "cfg_prefix_gflop": 40.86, "vae_gflop": 2514.52,
        "vae_upfold_gflop": 386.5,
        "name": "configs[3]: SD1.5-arch random-init txt2img 512x512, batch 64 sharded by seed across the ranks"},
    4: {"hw": 768, "unet_gflop": 4060.02, "fold_gflop": 179.1 * 2.25, "upfold_gflop": 75.5 * 2.25, "cfg_prefix_gflop": 152.3, "vae_gflop": 5754.30,
        "vae_upfold_gflop": 386.5 * 2.25,
        "name": "configs[4]: SD1.5-arch random-init txt2img 768x768 (4x96x96 latent, 9216-token self-attention)"},
}
CLIP_GFLOP_PER_PROMPT = 13.30


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops": d.get("bf16_tflops_sustained", 1390.6), "tflops_burst": d.get("bf16_tflops", 1654.9),
                "hbm_gbs": d.get("hbm_gbs", 6546.9), "source": "MEASURED_PEAKS.json (sustained bf16)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------- CPU arm
class ReferenceCPU:
    """The reference's own CPU implementation of the path on the host cores.

    kind "reference": the UNMODIFIED reference (sd/*.py compiled to bytecode under oracle/_ref by
    oracle/build_ref.py) - its own module classes built with the canonical seed-0 weights and its own
    pipeline.generate(device="cpu"), stub tokenizer (sd/pipeline.py:109 duck type), seed 42, CFG 7.5.
    kind "port": oracle/sd_oracle.py, only when oracle/_ref is absent or was built by another CPython.

    One sample = ONE real generate() call with `n_steps` denoising steps; forward hooks (which do not touch
    the reference's code) time every Diffusion.forward inside it. images/s for the 50-step workload =
    1 / (wall time of the call + (50 - n_steps) x mean Diffusion.forward time): CLIP x2, sampler, VAE decode and
    post-processing are measured as they ran, only the identical-work UNet evaluations are scaled. With
    n_steps = 50 nothing is extrapolated (bench.py --impl reference --full)."""

    def __init__(self, torch):
        self.torch = torch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        from pytorch_stable_diffusion_b200 import synthetic
        self.synthetic = synthetic
        self.cond, self.uncond = synthetic.canonical_tokens()
        self.ref = None
        try:
            import build_ref
            self.ref = build_ref.load()
        except Exception as e:                                  # noqa: BLE001 - any failure means "use the port"
            print(f"bench.py: oracle/_ref unusable ({e}); timing the oracle port instead", file=sys.stderr)
        t0 = time.perf_counter()
        if self.ref is not None:
            self.kind = "reference"
            torch.manual_seed(0)          # canonical weights: loader order (sd/model_loader.py:28-41), default init
            r = self.ref
            self.models = {"encoder": r["encoder"].VAE_Encoder().eval(), "decoder": r["decoder"].VAE_Decoder().eval(),
                           "diffusion": r["diffusion"].Diffusion().eval(), "clip": r["clip"].CLIP().eval()}
            self.unet_s = []
            self._t = None
            self.models["diffusion"].register_forward_pre_hook(self._pre)
            self.models["diffusion"].register_forward_hook(self._post)
        else:
            self.kind = "port"
            import sd_oracle
            self.sd_oracle = sd_oracle
            models = synthetic.build_models("cpu", which=("decoder", "diffusion", "clip"))
            self.weights = synthetic.state_dicts(models)
        self.build_s = time.perf_counter() - t0

    def _pre(self, mod, inp):
        self._t = time.perf_counter()

    def _post(self, mod, inp, out):
        self.unet_s.append(time.perf_counter() - self._t)

    def sample(self, n_steps):
        """{'call_s', 'unet_s' (mean per evaluation), 'n_steps'} of one real generate() call."""
        torch = self.torch
        if self.kind == "reference":
            tok = self.synthetic.StubTokenizer()
            self.unet_s = []
            t0 = time.perf_counter()
            img = self.ref["pipeline"].generate(prompt="a", uncond_prompt="b", input_image=None, strength=0.8,
                                                do_cfg=True, cfg_scale=CFG, sampler_name="ddpm",
                                                n_inference_steps=n_steps, models=self.models, seed=42,
                                                device="cpu", idle_device=None, tokenizer=tok)
            call = time.perf_counter() - t0
            assert img.shape == (H, W, 3) and len(self.unet_s) == n_steps
            return {"call_s": call, "unet_s": sum(self.unet_s) / n_steps, "n_steps": n_steps}
        with torch.no_grad():
            t0 = time.perf_counter()
            img, _ = self.sd_oracle.generate(self.weights, self.cond, self.uncond, seed=42, cfg_scale=CFG,
                                             n_inference_steps=n_steps, device="cpu")
            call = time.perf_counter() - t0
            t1 = time.perf_counter()
            lat = torch.zeros(2, 4, H // 8, W // 8)
            ctx = torch.zeros(2, 77, 768)
            self.sd_oracle.diffusion_forward(self.weights["diffusion"], lat, ctx, self.sd_oracle.get_time_embedding(500))
            unet = time.perf_counter() - t1
        return {"call_s": call, "unet_s": unet, "n_steps": n_steps}

    @staticmethod
    def images_per_s(s):
        return 1.0 / (s["call_s"] + (N_STEPS - s["n_steps"]) * s["unet_s"])

    def describe(self, n_steps, n_calls):
        what = ("the unmodified reference (oracle/_ref bytecode of sd/*.py): pipeline.generate(device='cpu')"
                if self.kind == "reference" else "oracle/sd_oracle.py port of the reference algorithm")
        if n_steps >= N_STEPS:
            return f"{what}, {n_calls} full 50-step txt2img call(s), batch 1, CFG 7.5, fp32 - nothing extrapolated"
        return (f"{what}, {n_calls} real call(s) with n_inference_steps={n_steps} (CLIP x2 + {n_steps} UNet CFG-pair "
                f"evaluation(s) + VAE decode + post-processing, batch 1, fp32); images/s = 1 / (call wall time + "
                f"{N_STEPS - n_steps} x mean Diffusion.forward time measured by forward hooks inside the call)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    cpu = ReferenceCPU(torch)
    n = N_STEPS if args.full else max(1, args.ref_steps)
    for _ in range(args.warmup if not args.full else 0):
        cpu.sample(1)
    samples = [cpu.sample(n) for _ in range(args.steps)]
    mean = {k: sum(s[k] for s in samples) / len(samples) for k in ("call_s", "unet_s")}
    mean["n_steps"] = n
    v = cpu.images_per_s(mean)
    sample = cpu.describe(n, args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * mean["call_s"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "configs[0]: SD1.5-arch random-init txt2img 512x512, batch 1, 50 DDPM steps, "
                                   "CFG 7.5, CPU (" + ("the reference's own pipeline.generate" if cpu.kind == "reference"
                                                      else "oracle port of the reference algorithm") + ")",
                       "torch_threads": torch.get_num_threads(), "n_inference_steps_per_timed_call": n},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cpu.cores, "kind": cpu.kind, "sample": sample,
                             "seconds": {k: round(x, 3) for k, x in mean.items()}, "model_build_s": round(cpu.build_s, 1)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--config", type=int, default=1, choices=[1, 3, 4],
                    help="BASELINE.json configs index: 1 = 512^2 batch 8 per GPU (the headline, weak scaling), "
                         "3 = 512^2 batch 64 split across the ranks (strong scaling), 4 = 768^2 batch 8 per GPU")
    ap.add_argument("--strong", action="store_true", help="same as --config 3")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full", action="store_true",
                    help="--impl reference: every timed step is a full 50-step reference generate (minutes per step)")
    ap.add_argument("--ref-steps", type=int, default=1,
                    help="--impl reference: denoising steps per timed reference call (bounded sample)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-only", action="store_true",
                    help="one eager UNet evaluation + decode (short command for ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.warmup < 3 and not args.profile_only:
        print("bench.py: raising --warmup to 3 (timing rule)", file=sys.stderr)
        args.warmup = 3

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; "
                         "use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from pytorch_stable_diffusion_b200 import _ext, engine, ops, pipeline, synthetic
    from pytorch_stable_diffusion_b200.ddpm import DDPMSampler
    peaks = load_peaks()
    if args.strong:
        args.config = 3
    wl = WORKLOADS[args.config]
    H = W = wl["hw"]
    metric = METRIC if H == 512 else f"{H}x{W} txt2img images/s (50-step DDPM, CFG 7.5)"
    UNET_GFLOP_PER_IMAGE_STEP = wl["unet_gflop"] - (wl["fold_gflop"] if engine.FOLD_GEGLU else 0.0) - (
        wl["upfold_gflop"] if engine.FOLD_UPSAMPLE else 0.0) - (wl["cfg_prefix_gflop"] if engine.SHARE_CFG_PREFIX else 0.0)
    VAE_GFLOP_PER_IMAGE = wl["vae_gflop"] - (wl["vae_upfold_gflop"] if engine.FOLD_UPSAMPLE else 0.0)
    if args.config == 3:
        if 64 % world:
            raise SystemExit("bench.py --config 3: 64 images do not split evenly over this many ranks")
        B = 64 // world                # BASELINE.json configs[3]: global batch 64, fixed
    else:
        B = args.batch
    lh, lw = H // 8, W // 8

    t_build = time.perf_counter()
    models = synthetic.build_models("cpu")
    for m in models.values():
        m.to(dev)
    t_build = time.perf_counter() - t_build

    # ---- device-resident inputs of one step (seeds are disjoint across ranks)
    cond, uncond = synthetic.canonical_tokens()
    tokens = torch.stack([cond, uncond]).to(dev)
    seeds = [42 + rank * B + i for i in range(B)]
    gens = [torch.Generator().manual_seed(s) for s in seeds]
    latents = torch.cat([torch.randn((1, 4, lh, lw), generator=g) for g in gens]).to(dev)
    sampler = DDPMSampler(torch.Generator())
    sampler.set_inference_timesteps(N_STEPS)
    step_noise = torch.stack([
        torch.cat([torch.randn((1, 4, lh, lw), generator=g) for g in gens]) if int(t) > 0
        else torch.zeros((B, 4, lh, lw)) for t in sampler.timesteps]).to(dev)
    coef = sampler.coefficient_table(dev)
    temb = torch.cat([pipeline.get_time_embedding(int(t)) for t in sampler.timesteps]).to(dev)
    idx = torch.tensor([0] * B + [1] * B, device=dev)

    def step_device():
        ctx = models["clip"](tokens)                   # (2, 77, 768): cond, uncond
        context = ctx[idx]                             # (2B, 77, 768), conditional rows first
        return pipeline.sample_on_device(models["diffusion"], models["decoder"], context, latents, step_noise,
                                         coef, temb, do_cfg=True, cfg_scale=CFG)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def mark(msg):
        if os.environ.get("SDB_BENCH_DEBUG"):
            torch.cuda.synchronize()
            print(f"[bench rank {rank}] {msg} fault={_ext.read_fault()}", file=sys.stderr, flush=True)

    if args.profile_only:
        with torch.no_grad():
            eng = models["diffusion"]._engine()
            ctx = models["clip"](tokens)[idx]
            kvs = eng.context_kv(ctx)
            tv = eng.time_vectors(temb)
            x = ops.nchw_to_nhwc(latents, repeat=2, out_fp32=True)
            eng.forward_nhwc(x, tv[0], kvs, cfg_pairs=True)
            models["decoder"].decode_nhwc(latents[:1])
        torch.cuda.synchronize()
        print(json.dumps({"profile_only": True, "fault": _ext.read_fault()}))
        return 0

    with torch.no_grad():
        # ---- warm-up (first call captures the CUDA graph of the 50-step loop)
        n_before_capture = _ext.launch_count()
        mark("inputs ready")
        t_cap = time.perf_counter()
        img = step_device()
        torch.cuda.synchronize()
        t_cap = time.perf_counter() - t_cap
        mark("first step (graph captured)")
        loop = next(iter(pipeline._GRAPH_CACHE.values()))
        for _ in range(args.warmup - 1):
            img = step_device()
        mark("warm-up done")
        barrier()
        mark("barrier passed")

        # ---- timed region: K steps, inputs already in HBM
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        clocks = ClockSampler(vis.split(",")[local] if vis else local)
        clocks.start()
        n0 = _ext.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            img = step_device()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        mark("timed region done")
        eager_launches = _ext.launch_count() - n0
        clk = clocks.stop()
        graph_launches = loop.graph_launches
        gpu_launches = eager_launches + args.steps * graph_launches
        tms = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
        value = world * B * args.steps / (ms / 1e3)

        # ---- phase split on the device (graph replay alone, decode alone), rank 0 only, not the headline
        def timed(fn, reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        loop_ms = timed(loop.replay, 2)
        dec_ms = timed(lambda: models["decoder"].decode_nhwc(loop.latents), 3)
        clip_ms = timed(lambda: models["clip"](tokens), 3)

        # ---- end to end through the public API: prompt strings in, host uint8 images out
        e2e = None
        if not args.no_e2e:
            tok = synthetic.StubTokenizer()
            kw = dict(models=models, batch_size=B, n_inference_steps=N_STEPS, cfg_scale=CFG, device=dev,
                      tokenizer=tok, return_all=True, height=H, width=W)
            out = pipeline.generate("a", "b", seed=1000 + rank, **kw)     # warm (same graph)
            barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            reps = max(1, min(args.steps, 3))
            for i in range(reps):
                out = pipeline.generate("a", "b", seed=2000 + rank * 100 + i, **kw)
            b.record()
            barrier()
            t = torch.tensor([a.elapsed_time(b)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            assert out.shape == (B, H, W, 3) and out.dtype.name == "uint8"
            e2e = {"value": world * B * reps / (float(t.item()) / 1e3), "unit": UNIT,
                   "h2d_bytes_per_step": 2 * 77 * 8 + N_STEPS * 5 * 4 + N_STEPS * 320 * 4,
                   "d2h_bytes_per_step": B * H * W * 3, "steps": reps,
                   "note": "pipeline.generate(prompt, uncond_prompt, batch_size=B, seed=...): tokens, DDPM "
                           "coefficient table and time embeddings copied H2D, device RNG, uint8 images copied D2H"}

        # ---- roofline pass: one eager UNet evaluation with per-launch CUDA events classifies the launches; the
        # dominant VARIANT of each kernel family is then re-timed the way the captured loop runs it
        roof, roof_hbm, roof_attn, rooflines, breakdown = None, None, None, None, None
        if rank == 0:
            eng = models["diffusion"]._engine()
            ops.PROFILER = ops.LaunchProfiler()
            eng.forward_nhwc(loop.x_in, loop.tvecs[0], loop.kvs, cfg_pairs=True)        # warm
            ops.PROFILER = ops.LaunchProfiler()
            ta = torch.cuda.Event(enable_timing=True)
            tb = torch.cuda.Event(enable_timing=True)
            ta.record()
            eng.forward_nhwc(loop.x_in, loop.tvecs[0], loop.kvs, cfg_pairs=True)
            tb.record()
            prof = ops.PROFILER
            ops.PROFILER = None
            summ = prof.summary()
            shapes = prof.summary(by_shape=True)
            unet_eager_ms = ta.elapsed_time(tb)
            relaunch = {}
            for name, _, _, _, _, shape, rl in prof.records:
                if rl is not None:
                    relaunch.setdefault((name, shape), []).append(rl)
            tpath = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
            ncu_traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
            def traffic_of(name, shape):
                """dram__bytes_read.sum + dram__bytes_write.sum of one cold-cache launch of this problem (ncu --set
                full, profiles/r02_ncu_traffic.json: keys are the tokens that identify the problem), or None."""
                have = set(f"{name} {shape}".split())
                for key, val in ncu_traffic.items():
                    if key != "_note" and set(key.split()) <= have:
                        return val
                return None

            KERNEL = {"gemm_tc_conv3x3": "gemm_tc_kernel<2> (CTA pairs), implicit-GEMM 3x3 conv",
                      "gemm_tc_linear": "gemm_tc_kernel<2> (CTA pairs), nn.Linear / 1x1 conv",
                      "attention": "attn2_tc_kernel / attn_tc_kernel (flash attention)",
                      "groupnorm": "gn_reduce_partials_kernel + gn_apply_kernel (or gn_fused_kernel)",
                      "layernorm": "layernorm_f32_kernel"}

            def graph_us(key, reps=24):
                """Average duration of one launch of variant `key`, the way the captured loop runs it: every launch
                of that variant in the UNet evaluation above (its own operand buffers, weights and outputs - real
                activations, not random data) issued back to back inside ONE CUDA graph, cycling through them until
                `reps` launches are recorded; CUDA events around a replay on the launching stream. Successive
                launches touch different buffers, so nothing is served from a previous launch's L2 lines unless the
                whole variant set fits (footprint is reported)."""
                rls = relaunch[key]
                n = max(reps, len(rls))
                n -= n % len(rls)
                for rl in rls:
                    rl()
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    for i in range(n):
                        rls[i % len(rls)]()
                gr.replay()
                torch.cuda.synchronize()
                best = None
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    gr.replay()
                    e1.record()
                    torch.cuda.synchronize()
                    us = 1e3 * e0.elapsed_time(e1) / n
                    best = us if best is None else min(best, us)
                return best, n, len(rls)

            def roof_of(key, bound):
                dv = shapes[key]
                name, shape = key
                fl, by = dv["flops"] / dv["launches"], dv["bytes"] / dv["launches"]
                us_eager = 1e3 * dv["ms"] / dv["launches"]
                us, n, nd = graph_us(key)
                if bound == "tensor":
                    # the kernel is timed ALONE (a few ms of back-to-back launches): the burst bf16 figure of
                    # MEASURED_PEAKS.json is its denominator; the sustained one is kept for the whole-job fractions
                    a, pk, unit = fl / (us * 1e-6) / 1e12, peaks["tflops_burst"], "TFLOP/s"
                else:
                    a, pk, unit = by / (us * 1e-6) / 1e9, peaks["hbm_gbs"], "GB/s"
                return {"bound": bound, "kernel": f"{KERNEL.get(name, name)}: {shape}", "achieved": a, "peak": pk,
                        "unit": unit, "frac": a / pk, "traffic": traffic_of(name, shape),
                        "peak_source": ("MEASURED_PEAKS.json (burst bf16: kernel timed alone)" if bound == "tensor" and
                                        "MEASURED" in peaks["source"] else peaks["source"]),
                        "frac_of_sustained_peak": (a / peaks["tflops"]) if bound == "tensor" else None,
                        "launches_per_unet_eval": dv["launches"],
                        "flops_per_launch": fl, "algorithmic_bytes_per_launch": by, "us_per_launch": us,
                        "us_per_launch_eager_events": us_eager, "share_of_unet_eval": dv["ms"] / unet_eager_ms,
                        "timing": f"CUDA events around a CUDA-graph replay of {n} back-to-back launches cycling through "
                                  f"the {nd} launches of this exact variant in one UNet evaluation (their own operands; "
                                  f"{nd * by / 1e6:.0f} MB distinct footprint vs 126 MB L2); achieved = "
                                  "flops_per_launch (or algorithmic_bytes_per_launch) / us_per_launch"}

            ridge = peaks["tflops"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
            fam = lambda n: [kv for kv in shapes.items() if kv[0][0] == n and kv[0] in relaunch]
            top = lambda kvs: max(kvs, key=lambda kv: kv[1]["ms"])[0] if kvs else None
            # `roofline`: the north-star's named kernel, the implicit-GEMM 3x3 conv (the dominant kernel by time
            # and by FLOPs), its dominant tensor-bound variant
            # (the problem size with the most FLOPs per evaluation first, full resolution winning ties - 320 -> 320 @
            # 64x64, nine launches per evaluation - then its most expensive launch variant: a stable choice)
            convs = [kv for kv in fam("gemm_tc_conv3x3") if kv[1]["flops"] / max(kv[1]["bytes"], 1.0) >= ridge]
            # a problem size = (cin, cout, taps, output pixels per sample): the shared CFG prefix runs two of the
            # 320 -> 320 @ 64x64 convs on half the batch, they are the same problem
            size_of = lambda kv: " ".join(t for t in kv[0][1].split() if t.split("=")[0] in ("cin", "cout", "taps", "hw"))
            by_size = {}
            for kv in convs:     # rank problem sizes by their FLOPs per evaluation (rounded to 1 %: the 64x64, 32x32 and
                # 16x16 levels tie exactly), then by resolution: no timing noise in the choice
                hw_ = int(dict(t.split("=") for t in kv[0][1].split())["hw"])
                fl_, _ = by_size.get(size_of(kv), (0.0, hw_))
                by_size[size_of(kv)] = (fl_ + kv[1]["flops"], hw_)
            by_size = {k: (round(v[0] / 1e10), v[1]) for k, v in by_size.items()}
            dom_size = max(by_size, key=by_size.get) if by_size else None
            k_conv = top([kv for kv in convs if size_of(kv) == dom_size])
            k_attn = top(fam("attention"))
            k_lin = top([kv for kv in fam("gemm_tc_linear") if kv[1]["flops"] / max(kv[1]["bytes"], 1.0) < ridge])
            # GroupNorm / LayerNorm: the variant that moves the most algorithmic bytes per evaluation (a deterministic
            # choice; by time the 8 x 8 level's one-pass GroupNorm - 10 MB per launch, L2-resident, launch-latency-bound -
            # can come out on top and says nothing about HBM)
            top_bytes = lambda kvs: max(kvs, key=lambda kv: kv[1]["bytes"])[0] if kvs else None
            k_gn, k_ln = top_bytes(fam("groupnorm")), top_bytes(fam("layernorm"))
            roof = roof_of(k_conv, "tensor")
            gemm = [v for k, v in summ.items() if k.startswith("gemm_tc")]
            roof["all_gemm_tc_launches"] = {"launches": sum(v["launches"] for v in gemm), "ms": sum(v["ms"] for v in gemm),
                                            "tflops": sum(v["flops"] for v in gemm) / (sum(v["ms"] for v in gemm) * 1e-3) / 1e12}
            roof_attn = roof_of(k_attn, "tensor") if k_attn else None
            hbm = [roof_of(k, "hbm") for k in (k_lin, k_gn, k_ln) if k]
            roof_hbm = max(hbm, key=lambda r: r["share_of_unet_eval"]) if hbm else None
            rooflines = [r for r in [roof, roof_attn] + hbm if r]
            breakdown = {k: {"launches": v["launches"], "ms": round(v["ms"], 3),
                             "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                             "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1) if v["bytes"] else None}
                         for k, v in sorted(summ.items())}
            breakdown["unet_eval_eager_ms"] = round(unet_eager_ms, 3)
            breakdown["shapes_eager_events"] = [
                {"shape": f"{k[0]} {k[1]}", "launches": v["launches"], "us": round(1e3 * v["ms"] / v["launches"], 1),
                 "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] else None,
                 "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)}
                for k, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:80]]
            del prof, relaunch, shapes, summ

    fault = _ext.read_fault()
    if fault:
        raise SystemExit(f"bench.py: device watchdog fault 0x{fault:x}")

    # ---- CPU baseline beside it (rank 0, N = 1): the reference's own generate() on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile_only:
        torch.cuda.empty_cache()
        ref_cpu = ReferenceCPU(torch)
        s = ref_cpu.sample(2)
        cpu = {"value": ref_cpu.images_per_s(s), "unit": UNIT, "cores": ref_cpu.cores, "kind": ref_cpu.kind,
               "sample": ref_cpu.describe(2, 1), "seconds": {k: round(v, 3) for k, v in s.items()},
               "model_build_s": round(ref_cpu.build_s, 1)}

    if rank == 0:
        unet_ms = loop_ms / N_STEPS
        alg_tflop_img = (N_STEPS * UNET_GFLOP_PER_IMAGE_STEP + VAE_GFLOP_PER_IMAGE) / 1e3
        line = {
            "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.config == 3 else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{wl['name']}, batch {B} per GPU, 50 DDPM steps, CFG 7.5, "
                                   "CUDA-graph-captured loop",
                       "batch_per_gpu": B, "global_batch": B * world, "n_inference_steps": N_STEPS,
                       "cfg_scale": CFG, "parallelism": f"seed-sharded x{world}, no collective",
                       "geglu_folded": bool(engine.FOLD_GEGLU), "upsample_folded": bool(engine.FOLD_UPSAMPLE),
                       "cfg_prefix_shared": bool(engine.SHARE_CFG_PREFIX),
                       "storage": "fp32 residual stream, latents and noise; 16-bit tensor-core operands (IEEE half at "
                                  "the full-resolution level, bf16 below); IEEE-half resblock hidden tensor "
                                  f"({bool(engine.HID_F16)}) and attention-block token stream ({bool(engine.TOK_F16)})",
                       "unet_gflop_per_image_step_algorithmic": UNET_GFLOP_PER_IMAGE_STEP,
                       "l2": "inputs larger than L2: 1.7 GB of bf16 weights streamed per UNet evaluation"},
            "clocks": clk, "e2e": e2e, "gpu_launches": gpu_launches,
            "roofline": roof, "roofline_hbm": roof_hbm, "roofline_attention": roof_attn, "cpu_baseline": cpu,
            "detail": {"unet_step_ms": unet_ms, "loop_ms": loop_ms, "vae_decode_ms": dec_ms, "clip_ms": clip_ms,
                       "graph_capture_s": t_cap, "model_build_s": t_build,
                       "launches_per_graph": graph_launches,
                       "unet_tensor_frac_of_sustained_peak":
                           B * UNET_GFLOP_PER_IMAGE_STEP / 1e3 / (unet_ms * 1e-3) / peaks["tflops"],
                       "whole_job_tensor_frac_of_sustained_peak":
                           value / world * alg_tflop_img / peaks["tflops"],
                       "rooflines": rooflines, "kernels": breakdown},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def _run():
    # The contract is ONE JSON line on stdout. Libraries (NCCL's version banner, for one) write to file
    # descriptor 1 on their own, so fd 1 is pointed at stderr for the whole run and the JSON line goes to
    # the saved original.
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real, "w")
    sys.stdout = out
    try:
        return main()
    finally:
        out.flush()


if __name__ == "__main__":
    sys.exit(_run())
