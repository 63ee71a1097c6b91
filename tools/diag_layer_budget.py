"""Diagnostic (GPU box): WHICH layers carry the bf16-vs-fp32 error of one UNet evaluation?

Runs the fp32 oracle with the tensor-core operands (weights and A operands of every conv / linear) rounded to a
16-bit type, then repeats the run with the rounding switched off for one top-level block (encoders.0 .. final) or
one operator kind at a time and prints how far the max-norm / RMS error drops. Also: what would fp16 operands buy.
Test infrastructure only (imports oracle/).   usage: python tools/diag_layer_budget.py [--hw 96] [--seeds 31,32]"""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sd_oracle as o  # noqa: E402
from pytorch_stable_diffusion_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--hw", type=int, default=96)
ap.add_argument("--seeds", default="31,32")
ap.add_argument("--t", type=int, default=980)
args = ap.parse_args()

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
BLOCKS = [f"enc{i}" for i in range(12)] + ["mid"] + [f"dec{i}" for i in range(12)] + ["final", "time"]
STATE = {"block": "time", "calls": 0}
CFG = {"w": torch.bfloat16, "a": torch.bfloat16, "skip_block": None, "skip_op": None, "w_only_block": None}


def rnd(t, kind, op):
    dt = CFG[kind]
    if dt is None or STATE["block"] == CFG["skip_block"] or op == CFG["skip_op"]:
        return t
    if kind == "a" and STATE["block"] == CFG["w_only_block"]:
        return t
    return t.to(dt).float()


def lin(sd, name, x):
    return F.linear(rnd(x, "a", name), rnd(sd[name + ".weight"], "w", name), sd.get(name + ".bias"))


def conv(sd, name, x, stride=1, padding=0):
    return F.conv2d(rnd(x, "a", name), rnd(sd[name + ".weight"], "w", name), sd.get(name + ".bias"), stride=stride,
                    padding=padding)


def switch_sequential(sd, x, context, time):
    STATE["block"] = BLOCKS[STATE["calls"]]
    STATE["calls"] += 1
    n_entries = 1 + max(int(k.split(".")[0]) for k in sd)
    for i in range(n_entries):
        m = o._sub(sd, str(i))
        if "groupnorm_feature.weight" in m:
            x = o.unet_residual_block(m, x, time)
        elif "attention_1.in_proj.weight" in m:
            x = o.unet_attention_block(m, x, context)
        elif "conv.weight" in m:
            x = conv(m, "conv", F.interpolate(x, scale_factor=2, mode="nearest"), padding=1)
        else:
            # the stem (4 -> 320) runs in fp32 on the CUDA cores in the product: not rounded
            if m["weight"].shape[1] <= 8:
                x = F.conv2d(x, m["weight"], m["bias"], stride=m["_stride"], padding=1)
            else:
                x = F.conv2d(rnd(x, "a", "down"), rnd(m["weight"], "w", "down"), m["bias"], stride=m["_stride"], padding=1)
    if STATE["calls"] == 25:
        STATE["block"] = "final"
    return x


o._lin, o._conv, o._switch_sequential = lin, conv, switch_sequential
models = synthetic.build_models(dev, which=("diffusion",))
sd = synthetic.state_dicts(models)["diffusion"]
OPS = ["conv_feature", "conv_merged", "residual_layer", "conv_input", "in_proj", "out_proj", "q_proj", "k_proj",
       "v_proj", "linear_geglu_1", "linear_geglu_2", "conv_output", "down", "conv", "linear_time"]


def run(lat, ctx, temb, **cfg):
    CFG.update({"w": torch.bfloat16, "a": torch.bfloat16, "skip_block": None, "skip_op": None, "w_only_block": None})
    CFG.update(cfg)
    STATE.update({"block": "time", "calls": 0})
    return o.diffusion_forward(sd, lat, ctx, temb)


def errs(y, ref):
    e = (y - ref).float()
    return float(e.abs().max() / ref.abs().max()), float(e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())


with torch.no_grad():
    for seed in [int(s) for s in args.seeds.split(",")]:
        g = torch.Generator().manual_seed(seed)
        lat = torch.randn(1, 4, args.hw, args.hw, generator=g).to(dev).repeat(2, 1, 1, 1)
        ctx = torch.randn(2, 77, 768, generator=g).to(dev)
        temb = o.get_time_embedding(args.t).to(dev)
        ref = run(lat, ctx, temb, w=None, a=None)
        base = errs(run(lat, ctx, temb), ref)
        got = errs(models["diffusion"](lat, ctx, temb), ref)
        print(f"== seed {seed} latent {args.hw} t={args.t}: emulated bf16 operands max {base[0]:.3e} rms {base[1]:.3e}; "
              f"kernels max {got[0]:.3e} rms {got[1]:.3e}", flush=True)
        for name, cfg in (("weights fp16, A bf16", dict(w=torch.float16)), ("weights bf16, A fp16", dict(a=torch.float16)),
                          ("weights fp16, A fp16", dict(w=torch.float16, a=torch.float16)),
                          ("weights exact, A bf16", dict(w=None)), ("weights bf16, A exact", dict(a=None))):
            e = errs(run(lat, ctx, temb, **cfg), ref)
            print(f"   {name:28s} max {e[0]:.3e} rms {e[1]:.3e}", flush=True)
        rows = []
        for b in BLOCKS[:-1]:
            e = errs(run(lat, ctx, temb, skip_block=b), ref)
            rows.append((base[1] ** 2 - e[1] ** 2, b, e))
        print("   exact operands in ONE block -> error (share of the squared RMS error it removes):")
        for d, b, e in sorted(rows, reverse=True):
            print(f"     {b:6s} max {e[0]:.3e} rms {e[1]:.3e}  share {100 * d / base[1] ** 2:5.1f} %", flush=True)
        rows = []
        for op in OPS:
            e = errs(run(lat, ctx, temb, skip_op=op), ref)
            rows.append((base[1] ** 2 - e[1] ** 2, op, e))
        print("   exact operands in ONE operator kind everywhere:")
        for d, op, e in sorted(rows, reverse=True):
            print(f"     {op:16s} max {e[0]:.3e} rms {e[1]:.3e}  share {100 * d / base[1] ** 2:5.1f} %", flush=True)
