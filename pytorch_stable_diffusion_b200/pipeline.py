"""pipeline.generate() with the reference's signature (sd/pipeline.py:13-27), running CLIP, the
optional VAE encoder, the classifier-free-guidance DDPM loop and the VAE decoder on hand-written
sm_100a kernels. The whole denoising loop (n_steps x [UNet + fused CFG/DDPM step]) is captured
once in a CUDA graph and replayed.

Extensions are keyword-only and default to the reference's behaviour (one 512x512 image, RNG drawn
from torch.Generator(device)): batch_size, height/width, per-sample seeds, injected noise,
return_all, use_cuda_graph, trace.
"""
import collections

import numpy as np
import torch

from . import _ext, ops
from .ddpm import DDPMSampler

WIDTH = 512
HEIGHT = 512
LATENTS_WIDTH = WIDTH // 8
LATENTS_HEIGHT = HEIGHT // 8
MAX_LATENT_TOKENS = 24576     # row length limit of sdb_softmax_rows (VAE attention, sd/decoder.py:34-73)

# Captured denoising loops, least recently used first. Each entry pins its static buffers and its graph's memory
# pool (a few GB at batch 8), so only the GRAPH_CACHE_SIZE most recently used configurations stay resident:
# alternating txt2img (50 steps) and img2img (40 steps), or two batch sizes, does not re-capture.
GRAPH_CACHE_SIZE = 3
_GRAPH_CACHE = collections.OrderedDict()


def rescale(x, old_range, new_range, clamp=False):
    """In-place affine range map, optional clamp (sd/pipeline.py:265-307)."""
    old_min, old_max = old_range
    new_min, new_max = new_range
    x -= old_min
    x *= (new_max - new_min) / (old_max - old_min)
    x += new_min
    if clamp:
        x = x.clamp(new_min, new_max)
    return x


def get_time_embedding(timestep):
    """(1, 320) sinusoidal embedding, cosine half first (sd/pipeline.py:310-349)."""
    freqs = torch.pow(10000, -torch.arange(start=0, end=160, dtype=torch.float32) / 160)
    x = torch.tensor([timestep], dtype=torch.float32)[:, None] * freqs[None]
    return torch.cat([torch.cos(x), torch.sin(x)], dim=-1)


def _encode_prompts(clip, tokenizer, prompts, device):
    ids = [tokenizer.batch_encode_plus([p], padding="max_length", max_length=77).input_ids[0] for p in prompts]
    uniq = {}
    for row in ids:
        uniq.setdefault(tuple(row), None)
    keys = list(uniq)
    tokens = torch.tensor(keys, dtype=torch.long, device=device)
    ctx = clip(tokens)                                   # (U, 77, 768) fp32
    index = [keys.index(tuple(row)) for row in ids]
    return ctx[index]


def _draw(shape, generator, seeds, device):
    """One noise tensor of `shape` = (B, ...). With per-sample seeds each sample owns a CPU
    generator (its stream equals an independent batch-1 reference run on CPU)."""
    if seeds is None:
        return torch.randn(shape, generator=generator, device=device)
    parts = [torch.randn((1,) + tuple(shape[1:]), generator=g) for g in seeds]
    return torch.cat(parts, 0).to(device)


_TABLE_CACHE = collections.OrderedDict()


def _step_tables(sampler, device):
    """Device-resident per-step constants of a schedule: DDPM coefficients [steps, 5] and sinusoidal time embeddings
    [steps, 320] (sd/pipeline.py:211, sd/ddpm.py:102-139). Both depend only on the timestep list and the betas, so
    they are built once per (schedule, device) instead of on every generate() call (100 small host ops + 2 copies)."""
    key = (tuple(int(t) for t in sampler.timesteps), float(sampler.betas[0]), float(sampler.betas[-1]),
           int(sampler.num_train_timesteps), str(device))
    hit = _TABLE_CACHE.get(key)
    if hit is None:
        coef = sampler.coefficient_table(device)
        temb = torch.cat([get_time_embedding(int(t)) for t in sampler.timesteps]).to(device)
        hit = (coef, temb)
        _TABLE_CACHE[key] = hit
        while len(_TABLE_CACHE) > 8:
            _TABLE_CACHE.popitem(last=False)
    else:
        _TABLE_CACHE.move_to_end(key)
    return hit


class _Loop:
    """Static buffers + captured CUDA graph of one denoising-loop configuration."""

    def __init__(self, eng, B, h, w, n_steps, do_cfg, cfg_scale, device):
        self.eng, self.B, self.n_steps, self.do_cfg, self.cfg_scale = eng, B, n_steps, do_cfg, cfg_scale
        N = 2 * B if do_cfg else B
        self.latents = torch.zeros((B, 4, h, w), device=device, dtype=torch.float32)
        self.x_in = torch.zeros((N, h, w, 4), device=device, dtype=torch.float32)   # UNet input stays fp32
        self.noise = torch.zeros((n_steps, B, 4, h, w), device=device, dtype=torch.float32)
        self.coef = torch.zeros((n_steps, 5), device=device, dtype=torch.float32)
        self.tvecs = torch.zeros((n_steps, eng.time_total), device=device, dtype=torch.float32)
        self.kvs = None
        self.graph = None
        self.graph_launches = 0

    def set_inputs(self, latents, noise, coef, tvecs, kvs):
        self.latents.copy_(latents)
        self.noise.copy_(noise)
        self.coef.copy_(coef)
        self.tvecs.copy_(tvecs)
        if self.kvs is None:
            self.kvs = [(k.clone(), v.clone()) for k, v in kvs]
        else:
            for (dk, dv), (k, v) in zip(self.kvs, kvs):
                dk.copy_(k)
                dv.copy_(v)
        self.x_in.copy_(ops.nchw_to_nhwc(self.latents, repeat=2 if self.do_cfg else 1, out_fp32=True))

    def run_steps(self, trace=None, timesteps=None):
        for i in range(self.n_steps):
            eps = self.eng.forward_nhwc(self.x_in, self.tvecs[i], self.kvs, cfg_pairs=self.do_cfg)
            if trace is not None:
                trace.append((int(timesteps[i]), self.latents.clone(), ops.nhwc_to_nchw_f32(eps)))
            ops.cfg_ddpm_step(self.latents, eps, self.noise[i], self.coef, i, self.cfg_scale, self.do_cfg,
                              self.x_in)

    def capture(self):
        # one eager UNet evaluation first: lazy CUDA/module initialisation must not happen under capture
        self.eng.forward_nhwc(self.x_in, self.tvecs[0], self.kvs, cfg_pairs=self.do_cfg)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = _ext.launch_count()
        with torch.cuda.graph(g):
            self.run_steps()
        self.graph_launches = _ext.launch_count() - n0     # kernels replayed by every graph launch
        self.graph = g

    def replay(self):
        self.graph.replay()


def generate(
    prompt,
    uncond_prompt=None,
    input_image=None,
    strength=0.8,
    do_cfg=True,
    cfg_scale=7.5,
    sampler_name="ddpm",
    n_inference_steps=50,
    models={},
    seed=None,
    device=None,
    idle_device=None,
    tokenizer=None,
    *,
    batch_size=1,
    height=None,
    width=None,
    seeds=None,
    noise=None,
    use_cuda_graph=True,
    return_all=False,
    trace=None,
):
    """Text-to-image / image-to-image sampling; returns a uint8 (H, W, 3) array like the reference
    (sd/pipeline.py:72-262), or (B, H, W, 3) with return_all=True.

    noise: optional dict of injected fp32 tensors {'latents' | ('encoder', 'add'), 'steps'} replacing
    every torch.randn draw (parity runs against the CPU oracle). trace: optional list receiving
    (timestep, latents_in, unet_output) per step (forces eager execution).
    """
    with torch.no_grad():
        if not 0 < strength <= 1:
            raise ValueError(f"Strength must be between 0 and 1, got {strength}")
        if sampler_name != "ddpm":
            raise ValueError(f"Sampler {sampler_name} not found")
        if idle_device:
            to_idle = lambda x: x.to(idle_device)
        else:
            to_idle = lambda x: x
        if device is None:
            device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("this build runs on hand-written CUDA kernels for sm_100a only; "
                               f"device={device} is not supported (no CPU fallback)")
        H = HEIGHT if height is None else height
        W = WIDTH if width is None else width
        if H % 64 or W % 64:
            raise ValueError("height and width must be multiples of 64")
        if (H // 8) * (W // 8) > MAX_LATENT_TOKENS:
            # the VAE's single-head d = 512 attention keeps one fp32 score row per query in shared memory
            # (sdb_softmax_rows: 24 576 columns) and an S x S fp32 score matrix per sample in HBM
            raise ValueError(f"{H}x{W} exceeds the largest supported image: (H/8)*(W/8) must not exceed "
                             f"{MAX_LATENT_TOKENS} latent positions (e.g. 1248x1248)")
        lh, lw = H // 8, W // 8
        B = batch_size

        generator = torch.Generator(device=device)
        if seed is None:
            generator.seed()
        else:
            generator.manual_seed(seed)
        sample_gens = None
        if seeds is not None:
            if len(seeds) != B:
                raise ValueError("seeds must have batch_size entries")
            sample_gens = [torch.Generator(device="cpu").manual_seed(int(s)) for s in seeds]

        # ---- CLIP (sd/pipeline.py:101-134)
        clip = models["clip"]
        clip.to(device)
        prompts = list(prompt) if isinstance(prompt, (list, tuple)) else [prompt] * B
        if len(prompts) != B:
            raise ValueError("a list-valued prompt must have batch_size entries")
        if do_cfg:
            unconds = (list(uncond_prompt) if isinstance(uncond_prompt, (list, tuple))
                       else [uncond_prompt if uncond_prompt is not None else ""] * B)
            context = torch.cat([_encode_prompts(clip, tokenizer, prompts, device),
                                 _encode_prompts(clip, tokenizer, unconds, device)])
        else:
            context = _encode_prompts(clip, tokenizer, prompts, device)
        to_idle(clip)

        sampler = DDPMSampler(generator)
        sampler.set_inference_timesteps(n_inference_steps)
        latents_shape = (B, 4, lh, lw)

        # ---- initial latents (sd/pipeline.py:149-196)
        has_image = input_image is not None and not (isinstance(input_image, (list, tuple)) and not input_image)
        if has_image:
            encoder = models["encoder"]
            encoder.to(device)
            # resize (Pillow's bicubic resampler, byte-exact) + rescale to [-1, 1] on the device; one image for the
            # whole batch, or a list with one image per sample
            from . import imageio
            _, x = imageio.load_images(input_image, W, H, device)          # fp32 NHWC in [-1, 1]
            if x.shape[0] == 1 and B > 1:
                x = x.expand(B, -1, -1, -1).contiguous()
            elif x.shape[0] != B:
                raise ValueError("a list-valued input_image must have batch_size entries")
            enc_noise = noise["encoder"].to(device) if noise is not None else _draw(
                latents_shape, generator, sample_gens, device)
            latents = encoder._engine().forward_from_nhwc(x, enc_noise.to(torch.float32).contiguous())
            sampler.set_strength(strength=strength)
            t0 = int(sampler.timesteps[0])
            add = noise["add"].to(device) if noise is not None else _draw(
                latents_shape, generator, sample_gens, device)
            a = sampler.alphas_cumprod[t0]
            latents = ops.axpby(latents, add.to(torch.float32).contiguous(), float(a ** 0.5),
                                float((1 - a) ** 0.5))
            to_idle(encoder)
        else:
            latents = noise["latents"].to(device) if noise is not None else _draw(
                latents_shape, generator, sample_gens, device)
        latents = latents.to(torch.float32).contiguous()

        # ---- per-step constants: noise, DDPM coefficients, time embeddings
        timesteps = sampler.timesteps
        n_steps = len(timesteps)
        if noise is not None:
            step_noise = noise["steps"].to(device=device, dtype=torch.float32, non_blocking=True)
        else:
            draws = [_draw(latents_shape, generator, sample_gens, device) if int(t) > 0
                     else torch.zeros(latents_shape, device=device) for t in timesteps]
            step_noise = torch.stack(draws)
        if step_noise.shape[0] < n_steps:   # the last step (t = 0) draws nothing
            pad = torch.zeros((n_steps - step_noise.shape[0],) + tuple(latents_shape), device=device)
            step_noise = torch.cat([step_noise, pad])
        coef, temb = _step_tables(sampler, device)

        diffusion = models["diffusion"]
        diffusion.to(device)
        decoder = models["decoder"]
        decoder.to(device)
        images = sample_on_device(diffusion, decoder, context, latents, step_noise, coef, temb,
                                  do_cfg=do_cfg, cfg_scale=cfg_scale, use_cuda_graph=use_cuda_graph,
                                  trace=trace, timesteps=timesteps)
        images = images.to("cpu").numpy()
        check_device_fault("pipeline.generate")
        if idle_device:
            # idle_device means "give the GPU memory back" (sd/pipeline.py:80-85): besides the fp32 parameters that
            # is the packed bf16 engines and the captured loop, which would otherwise keep everything resident.
            # The next call repacks and re-captures.
            drop_cached_graphs(diffusion.__dict__.get("_sdb_engine", (None, None))[1])
            for m in (clip, diffusion, decoder) + ((models["encoder"],) if has_image else ()):
                m.invalidate_packed()
            diffusion.__dict__.pop("_sdb_ctx", None)
        to_idle(diffusion)
        to_idle(decoder)
        return images if return_all else images[0]


def sample_on_device(diffusion, decoder, context, latents, step_noise, coef, temb, *, do_cfg=True,
                     cfg_scale=7.5, use_cuda_graph=True, trace=None, timesteps=None):
    """Device-resident core of generate(): everything is already in HBM.

    context fp32 [N, 77, 768] (N = 2B with CFG: conditional rows first), latents fp32 [B, 4, h, w],
    step_noise fp32 [steps, B, 4, h, w], coef fp32 [steps, 5] (DDPMSampler.coefficient_table),
    temb fp32 [steps, 320] (get_time_embedding rows). Runs the denoising loop (sd/pipeline.py:205-237)
    as one CUDA graph and decodes (sd/pipeline.py:243-259). Returns uint8 NHWC [B, 8h, 8w, 3] on device.
    """
    device = latents.device
    B, _, lh, lw = latents.shape
    n_steps = step_noise.shape[0]
    eng = diffusion._engine()
    tvecs = eng.time_vectors(temb)
    kvs = eng.context_kv(context)
    key = (id(eng), B, lh, lw, n_steps, bool(do_cfg), float(cfg_scale), device.index)
    loop = _GRAPH_CACHE.get(key)
    if loop is None:
        loop = _Loop(eng, B, lh, lw, n_steps, bool(do_cfg), float(cfg_scale), device)
        _GRAPH_CACHE[key] = loop
        while len(_GRAPH_CACHE) > GRAPH_CACHE_SIZE:
            _GRAPH_CACHE.popitem(last=False)
    else:
        _GRAPH_CACHE.move_to_end(key)
    loop.set_inputs(latents, step_noise, coef, tvecs, kvs)
    if trace is not None or not use_cuda_graph:
        loop.run_steps(trace=trace, timesteps=timesteps)
    else:
        if loop.graph is None:
            loop.capture()
            loop.set_inputs(latents, step_noise, coef, tvecs, kvs)
        loop.replay()
    images = decoder.decode_nhwc(loop.latents)           # fp32 NHWC in ~[-1, 1]
    return ops.image_to_uint8(images)


def drop_cached_graphs(engine=None):
    """Releases the captured loops (all, or those of one UNet engine) and the memory their graphs pin."""
    for key in [k for k, lp in _GRAPH_CACHE.items() if engine is None or lp.eng is engine]:
        del _GRAPH_CACHE[key]


def check_device_fault(what):
    """Raises if a kernel's mbarrier watchdog tripped since the last check (a wait that timed out lets the kernel run
    on with data that was not ready: every result since then is suspect). Synchronises the device."""
    fault = _ext.read_fault()
    if fault:
        raise _ext.SdbError(f"{what}: device watchdog fault 0x{fault:x} (mbarrier wait timed out at site {fault >> 8}); "
                            "the results of this call are invalid - the fault word has been cleared, retry the call")
