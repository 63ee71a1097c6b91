"""Diagnostic (GPU box): where does the bf16-vs-fp32 error of one UNet evaluation come from?
Runs the fp32 oracle with selected tensors rounded to bf16 and prints max|y-y_ref|/max|y_ref|.
Test infrastructure only (imports oracle/)."""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sd_oracle as o  # noqa: E402
from pytorch_stable_diffusion_b200 import synthetic  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
r = lambda t: t.bfloat16().float()
FLAGS = {"w": False, "a": False, "out": False, "attn": False}
orig_lin, orig_conv = o._lin, o._conv
orig_sa, orig_ca = o.self_attention, o.cross_attention


def lin(sd, name, x):
    w = sd[name + ".weight"]
    y = F.linear(r(x) if FLAGS["a"] else x, r(w) if FLAGS["w"] else w, sd.get(name + ".bias"))
    return r(y) if FLAGS["out"] and name in ("in_proj", "q_proj", "k_proj", "v_proj", "linear_geglu_1") else y


def conv(sd, name, x, stride=1, padding=0):
    w = sd[name + ".weight"]
    y = F.conv2d(r(x) if FLAGS["a"] else x, r(w) if FLAGS["w"] else w, sd.get(name + ".bias"), stride=stride,
                 padding=padding)
    return r(y) if FLAGS["out"] and name == "conv_feature" else y


def sa(sd, x, n_heads, causal=False):
    if not FLAGS["attn"]:
        return orig_sa(sd, x, n_heads, causal)
    b, s, e = x.shape
    dh = e // n_heads
    q, k, v = o._lin(sd, "in_proj", x).chunk(3, dim=-1)
    q, k, v = (r(t).view(b, s, n_heads, dh).transpose(1, 2) for t in (q, k, v))
    w = F.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    out = r((r(w) @ v).transpose(1, 2).reshape(b, s, e))
    return o._lin(sd, "out_proj", out)


o._lin, o._conv, o.self_attention = lin, conv, sa
models = synthetic.build_models(dev, which=("diffusion",))
sd = synthetic.state_dicts(models)["diffusion"]
g = torch.Generator().manual_seed(21)
lat = torch.randn(2, 4, 64, 64, generator=g).to(dev)
ctx = torch.randn(2, 77, 768, generator=g).to(dev)
temb = o.get_time_embedding(980).to(dev)
with torch.no_grad():
    ref = o.diffusion_forward(sd, lat, ctx, temb)
    for name, fl in [("weights bf16", dict(w=True)), ("A operands bf16", dict(a=True)),
                     ("weights + A", dict(w=True, a=True)),
                     ("weights + A + branch outputs (hid, qkv, geglu hidden)", dict(w=True, a=True, out=True)),
                     ("all + attention P/V/O", dict(w=True, a=True, out=True, attn=True))]:
        FLAGS.update({"w": False, "a": False, "out": False, "attn": False})
        FLAGS.update(fl)
        y = o.diffusion_forward(sd, lat, ctx, temb)
        print(f"{name:60s} rel_err = {o.rel_err(y, ref):.3e}   rms = "
              f"{float((y - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.3e}", flush=True)
    got = models["diffusion"](lat, ctx, temb)
    print(f"{'kernels':60s} rel_err = {o.rel_err(got, ref):.3e}   rms = "
          f"{float((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.3e}", flush=True)
