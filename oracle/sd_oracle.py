"""CPU oracle for the Stable Diffusion sampling path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch restatement of the algorithm in dawmro/pytorch_stable_diffusion (sd/*.py),
written functionally over state_dicts so it shares no code with the package under test. Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference legs may import it.

Pinning: the reference repository ships no tests, golden vectors or fixtures ("parity unpinned" by
the reference's own tests). This oracle is pinned instead against outputs of the
reference itself, executed in the build container by oracle/make_golden.py and committed under
tests/golden/ (tests/test_oracle_golden.py re-checks them on every CPU run).

Every function cites the reference lines it restates (paths relative to the reference checkout).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

VAE_SCALE = 0.18215


# ------------------------------------------------------------------------------------------------
# helpers
def _sub(sd, prefix):
    """View of a state_dict under `prefix.` with the prefix stripped."""
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(x, sd[name + ".weight"], sd.get(name + ".bias"), stride=stride, padding=padding)


def _gn(sd, name, x, eps=1e-5):
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _ln(sd, name, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.shape[0],), w, sd[name + ".bias"], 1e-5)


# ------------------------------------------------------------------------------------------------
# sd/attention.py
def self_attention(sd, x, n_heads, causal=False):
    """SelfAttention.forward — sd/attention.py:27-93 (mask before scaling, :58-66)."""
    b, s, e = x.shape
    dh = e // n_heads
    q, k, v = _lin(sd, "in_proj", x).chunk(3, dim=-1)
    q = q.view(b, s, n_heads, dh).transpose(1, 2)
    k = k.view(b, s, n_heads, dh).transpose(1, 2)
    v = v.view(b, s, n_heads, dh).transpose(1, 2)
    w = q @ k.transpose(-1, -2)
    if causal:
        w = w.masked_fill(torch.ones_like(w, dtype=torch.bool).triu(1), -torch.inf)
    w = w / math.sqrt(dh)
    w = F.softmax(w, dim=-1)
    o = (w @ v).transpose(1, 2).reshape(b, s, e)
    return _lin(sd, "out_proj", o)


def cross_attention(sd, x, y, n_heads):
    """CrossAttention.forward — sd/attention.py:161-253."""
    b, s, e = x.shape
    dh = e // n_heads
    q = _lin(sd, "q_proj", x).view(b, -1, n_heads, dh).transpose(1, 2)
    k = _lin(sd, "k_proj", y).view(b, -1, n_heads, dh).transpose(1, 2)
    v = _lin(sd, "v_proj", y).view(b, -1, n_heads, dh).transpose(1, 2)
    w = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    w = F.softmax(w, dim=-1)
    o = (w @ v).transpose(1, 2).contiguous().view(b, s, e)
    return _lin(sd, "out_proj", o)


# ------------------------------------------------------------------------------------------------
# sd/clip.py
def clip_layer(sd, x, n_heads=12):
    """CLIPLayer.forward — sd/clip.py:123-176 (quick-GELU at :166)."""
    r = x
    x = _ln(sd, "layernorm_1", x)
    x = self_attention(_sub(sd, "attention"), x, n_heads, causal=True) + r
    r = x
    x = _ln(sd, "layernorm_2", x)
    x = _lin(sd, "linear_1", x)
    x = x * torch.sigmoid(1.702 * x)
    return _lin(sd, "linear_2", x) + r


def clip_forward(sd, tokens):
    """CLIP.forward — sd/clip.py:227-261; CLIPEmbedding.forward :38-66."""
    tokens = tokens.type(torch.long)
    x = F.embedding(tokens, sd["embedding.token_embedding.weight"]) + sd["embedding.position_embedding"]
    n_layers = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("layers."))
    for i in range(n_layers):
        x = clip_layer(_sub(sd, f"layers.{i}"), x)
    return _ln(sd, "layernorm", x)


# ------------------------------------------------------------------------------------------------
# sd/decoder.py, sd/encoder.py
def vae_attention_block(sd, x):
    """VAE_AttentionBlock.forward — sd/decoder.py:34-73.

    As written in the reference: the declared GroupNorm is never applied, and `x.transpose(-1, 2)`
    (:62) swaps an axis with itself, so the (n, hw, c) attention output is re-viewed raw as
    (n, c, h, w) before the residual add.
    """
    n, c, h, w = x.shape
    t = x.view(n, c, h * w).transpose(-1, -2)
    t = self_attention(_sub(sd, "attention"), t, 1)
    return t.reshape(n, c, h, w) + x


def vae_residual_block(sd, x):
    """VAE_ResidualBlock.forward — sd/decoder.py:135-189."""
    r = x
    x = _conv(sd, "conv_1", F.silu(_gn(sd, "groupnorm_1", x)), padding=1)
    x = _conv(sd, "conv_2", F.silu(_gn(sd, "groupnorm_2", x)), padding=1)
    if "residual_layer.weight" in sd:
        r = _conv(sd, "residual_layer", r)
    return x + r


def _vae_sequential(sd, x, n_entries, pad_stride2):
    for i in range(n_entries):
        m = _sub(sd, str(i))
        if "groupnorm_1.weight" in m:
            x = vae_residual_block(m, x)
        elif "attention.in_proj.weight" in m:
            x = vae_attention_block(m, x)
        elif "weight" in m and m["weight"].dim() == 4:
            k = m["weight"].shape[-1]
            if pad_stride2 is not None and i in pad_stride2:
                x = F.conv2d(F.pad(x, (0, 1, 0, 1)), m["weight"], m["bias"], stride=2)
            else:
                x = F.conv2d(x, m["weight"], m["bias"], padding=(k - 1) // 2)
        elif "weight" in m and m["weight"].dim() == 1:
            x = F.group_norm(x, 32, m["weight"], m["bias"], 1e-5)
        else:
            x = _VAE_PARAMLESS[(n_entries, i)](x)
    return x


# parameter-free entries of the two nn.Sequential containers, keyed by (container length, index)
_VAE_PARAMLESS = {
    (26, 8): lambda x: F.interpolate(x, scale_factor=2.0, mode="nearest"),   # sd/decoder.py:269
    (26, 13): lambda x: F.interpolate(x, scale_factor=2.0, mode="nearest"),  # :289
    (26, 18): lambda x: F.interpolate(x, scale_factor=2.0, mode="nearest"),  # :309
    (26, 24): F.silu,                                                         # :335
    (19, 16): F.silu,                                                         # sd/encoder.py:88
}


def vae_decoder_forward(sd, x):
    """VAE_Decoder.forward — sd/decoder.py:342-374 (26-entry nn.Sequential, :232-340)."""
    return _vae_sequential(sd, x / VAE_SCALE, 26, None)


def vae_encoder_forward(sd, x, noise):
    """VAE_Encoder.forward — sd/encoder.py:95-155; stride-2 convs (entries 3, 6, 9) are preceded
    by a right/bottom zero pad (:120-122)."""
    x = _vae_sequential(sd, x, 19, (3, 6, 9))
    mean, logvar = torch.chunk(x, 2, dim=1)
    stdev = torch.clamp(logvar, -30, 20).exp().sqrt()
    return (mean + stdev * noise) * VAE_SCALE


# ------------------------------------------------------------------------------------------------
# sd/diffusion.py
def time_embedding(sd, t):
    """TimeEmbedding.forward — sd/diffusion.py:44-80."""
    return _lin(sd, "linear_2", F.silu(_lin(sd, "linear_1", t)))


def unet_residual_block(sd, x, time):
    """UNET_ResidualBlock.forward — sd/diffusion.py:145-209."""
    r = x
    h = _conv(sd, "conv_feature", F.silu(_gn(sd, "groupnorm_feature", x)), padding=1)
    t = _lin(sd, "linear_time", F.silu(time))
    h = h + t.unsqueeze(-1).unsqueeze(-1)
    h = _conv(sd, "conv_merged", F.silu(_gn(sd, "groupnorm_merged", h)), padding=1)
    if "residual_layer.weight" in sd:
        r = _conv(sd, "residual_layer", r)
    return h + r


def unet_attention_block(sd, x, context, n_heads=8):
    """UNET_AttentionBlock.forward — sd/diffusion.py:271-381.

    As written in the reference (:359-363) the GEGLU gate half of linear_geglu_1 is discarded and no
    GELU is applied: x = linear_geglu_2(linear_geglu_1(x)[..., :4C]).
    """
    long_res = x
    x = _conv(sd, "conv_input", _gn(sd, "groupnorm", x, eps=1e-6))
    n, c, h, w = x.shape
    x = x.view(n, c, h * w).transpose(-1, -2)
    x = self_attention(_sub(sd, "attention_1"), _ln(sd, "layernorm_1", x), n_heads) + x
    x = cross_attention(_sub(sd, "attention_2"), _ln(sd, "layernorm_2", x), context, n_heads) + x
    g = _lin(sd, "linear_geglu_1", _ln(sd, "layernorm_3", x))
    g, _gate = g.chunk(2, dim=-1)
    x = _lin(sd, "linear_geglu_2", g) + x
    x = x.transpose(-1, -2).reshape(n, c, h, w)
    return _conv(sd, "conv_output", x) + long_res


def _switch_sequential(sd, x, context, time):
    """SwitchSequential.forward — sd/diffusion.py:458-496, dispatching on the parameter names."""
    n_entries = 1 + max(int(k.split(".")[0]) for k in sd)
    for i in range(n_entries):
        m = _sub(sd, str(i))
        if "groupnorm_feature.weight" in m:
            x = unet_residual_block(m, x, time)
        elif "attention_1.in_proj.weight" in m:
            x = unet_attention_block(m, x, context)
        elif "conv.weight" in m:  # Upsample — sd/diffusion.py:412-435
            x = _conv(m, "conv", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
        else:  # bare nn.Conv2d; the stride-2 ones are encoders 3, 6, 9 (:553,561,569)
            x = F.conv2d(x, m["weight"], m["bias"], stride=m["_stride"], padding=1)
    return x


def unet_forward(sd, x, context, time):
    """UNET.forward — sd/diffusion.py:628-676."""
    skips = []
    for i in range(12):
        m = _sub(sd, f"encoders.{i}")
        if i in (0, 3, 6, 9):
            m["0._stride"] = 1 if i == 0 else 2
        x = _switch_sequential(m, x, context, time)
        skips.append(x)
    x = _switch_sequential(_sub(sd, "bottleneck"), x, context, time)
    for i in range(12):
        x = torch.cat((x, skips.pop()), dim=1)
        x = _switch_sequential(_sub(sd, f"decoders.{i}"), x, context, time)
    return x


def diffusion_forward(sd, latent, context, time):
    """Diffusion.forward — sd/diffusion.py:797-837; UNET_OutputLayer :714-748."""
    t = time_embedding(_sub(sd, "time_embedding"), time)
    x = unet_forward(_sub(sd, "unet"), latent, context, t)
    f = _sub(sd, "final")
    return _conv(f, "conv", F.silu(_gn(f, "groupnorm", x)), padding=1)


# ------------------------------------------------------------------------------------------------
# sd/ddpm.py
class OracleDDPM:
    """DDPMSampler — sd/ddpm.py:30-186, restated with the noise source injectable."""

    def __init__(self, noise_fn, num_training_steps=1000, beta_start=0.000085, beta_end=0.012):
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_training_steps,
                                    dtype=torch.float32) ** 2          # :43
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)   # :45-48
        self.n_train = num_training_steps
        self.noise_fn = noise_fn
        self.timesteps = torch.arange(num_training_steps - 1, -1, -1)  # :53

    def set_inference_timesteps(self, n=50):                            # :56-63
        self.n_inf = n
        ratio = self.n_train // n
        self.timesteps = torch.from_numpy((np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64))

    def set_strength(self, strength=1.0):                               # :90-99
        start = self.n_inf - int(self.n_inf * strength)
        self.timesteps = self.timesteps[start:]
        self.start_step = start

    def coefficients(self, t):
        """(sqrt(1-abar_t), sqrt(abar_t), c_x0, c_xt, sigma_t) of step() — :104-133."""
        t = int(t)
        prev_t = t - self.n_train // self.n_inf
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_a = a_t / a_prev
        cur_b = 1 - cur_a
        c_x0 = (a_prev ** 0.5 * cur_b) / b_t
        c_xt = cur_a ** 0.5 * b_prev / b_t
        var = torch.clamp((1 - a_prev) / (1 - a_t) * cur_b, min=1e-20)
        sigma = var ** 0.5 if t > 0 else torch.tensor(0.0)
        return b_t ** 0.5, a_t ** 0.5, c_x0, c_xt, sigma

    def step(self, t, latents, model_output):                           # :102-139
        sb, sa, c_x0, c_xt, sigma = (c.to(latents.device) for c in self.coefficients(t))
        x0 = (latents - sb * model_output) / sa
        prev = c_x0 * x0 + c_xt * latents
        if int(t) > 0:
            prev = prev + sigma * self.noise_fn(model_output.shape)
        return prev

    def add_noise(self, x0, t):                                         # :143-186
        a = self.alphas_cumprod[int(t)].to(x0.device)
        return a ** 0.5 * x0 + (1 - a) ** 0.5 * self.noise_fn(x0.shape)


# ------------------------------------------------------------------------------------------------
# sd/pipeline.py
def get_time_embedding(timestep):
    """sd/pipeline.py:310-349 (cos block first)."""
    freqs = torch.pow(10000, -torch.arange(start=0, end=160, dtype=torch.float32) / 160)
    x = torch.tensor([timestep], dtype=torch.float32)[:, None] * freqs[None]
    return torch.cat([torch.cos(x), torch.sin(x)], dim=-1)


def generate(weights, cond_tokens, uncond_tokens, *, seed=42, cfg_scale=7.5, n_inference_steps=50,
             do_cfg=True, input_image=None, strength=0.8, latent_hw=(64, 64), max_steps=None,
             trace=None, device="cpu"):
    """pipeline.generate — sd/pipeline.py:72-262 for one image, token ids instead of a tokenizer.

    weights: {'clip','encoder','decoder','diffusion'} state_dicts. The noise stream is a CPU
    torch.Generator seeded with `seed`, drawn in the reference's order (:177 encoder noise,
    sd/ddpm.py:184 add_noise, :196 initial latents, sd/ddpm.py:131 per step).
    `device` places the arithmetic (the noise is always drawn on the CPU and moved).
    `max_steps` stops the loop early (bounded CPU-baseline samples); `trace` (a list) receives
    (timestep, latents_in, unet_out) per step. Returns (uint8 HxWx3 image, final latents).
    """
    with torch.no_grad():
        gen = torch.Generator(device="cpu").manual_seed(seed)
        randn = lambda shape: torch.randn(tuple(shape), generator=gen).to(device)
        cond_tokens = cond_tokens.to(device)
        uncond_tokens = uncond_tokens.to(device) if uncond_tokens is not None else None
        if do_cfg:
            ctx = torch.cat([clip_forward(weights["clip"], cond_tokens.view(1, -1)),
                             clip_forward(weights["clip"], uncond_tokens.view(1, -1))])
        else:
            ctx = clip_forward(weights["clip"], cond_tokens.view(1, -1))
        sampler = OracleDDPM(randn)
        sampler.set_inference_timesteps(n_inference_steps)
        shape = (1, 4, latent_hw[0], latent_hw[1])
        if input_image is not None:
            img = torch.tensor(np.asarray(input_image), dtype=torch.float32, device=device)
            img = (img * (2.0 / 255.0) - 1.0).unsqueeze(0).permute(0, 3, 1, 2)
            latents = vae_encoder_forward(weights["encoder"], img, randn(shape))
            sampler.set_strength(strength)
            latents = sampler.add_noise(latents, sampler.timesteps[0])
        else:
            latents = randn(shape)
        for i, t in enumerate(sampler.timesteps):
            if max_steps is not None and i >= max_steps:
                break
            temb = get_time_embedding(int(t)).to(device)
            x = latents.repeat(2, 1, 1, 1) if do_cfg else latents
            out = diffusion_forward(weights["diffusion"], x, ctx, temb)
            if trace is not None:
                trace.append((int(t), latents.clone(), out.clone()))
            if do_cfg:
                c, u = out.chunk(2)
                out = cfg_scale * (c - u) + u
            latents = sampler.step(t, latents, out)
        img = vae_decoder_forward(weights["decoder"], latents.clone())
        img = ((img + 1.0) * 127.5).clamp(0, 255).permute(0, 2, 3, 1)
        return img.to("cpu", torch.uint8).numpy()[0], latents


# ------------------------------------------------------------------------------------------------
# metrics
def rel_err(got, ref):
    """max|got - ref| / max|ref| (north_star per-step metric)."""
    return float((got.float() - ref.float()).abs().max() / ref.float().abs().max().clamp_min(1e-20))


def psnr_u8(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mse = np.mean((a - b) ** 2)
    return float("inf") if mse == 0 else 10.0 * math.log10(255.0 ** 2 / mse)
