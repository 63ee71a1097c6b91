"""Generates tests/golden/*.pt by EXECUTING THE UNMODIFIED REFERENCE (/root/reference/sd) on CPU.

Run in the build container only (the reference checkout does not travel to the GPU box):
    python oracle/make_golden.py [--full]
The fixtures pin oracle/sd_oracle.py (tests/test_oracle_golden.py) and give the GPU parity tests
reference outputs to compare against. Canonical inputs:
  weights  torch.manual_seed(0); VAE_Encoder(), VAE_Decoder(), Diffusion(), CLIP() in loader order
  tokens   torch.Generator().manual_seed(7): cond = randint(0, 49408, (77,)), uncond = next draw
  noise    CPU torch.Generator seeded 42, drawn in the reference's own order
"""
import argparse
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/sd"


def load_reference():
    """Imports the reference's flat modules under their own names in an isolated sys.path entry."""
    sys.path.insert(0, REF)
    mods = {}
    for name in ("attention", "clip", "ddpm", "decoder", "diffusion", "encoder", "pipeline"):
        mods[name] = importlib.import_module(name)
    sys.path.remove(REF)
    return mods


class StubTokenizer:
    """pipeline.generate only calls tokenizer.batch_encode_plus(...).input_ids (sd/pipeline.py:109)."""

    def __init__(self, table):
        self.table = table

    def batch_encode_plus(self, prompts, padding=None, max_length=None):
        class R:
            pass
        r = R()
        r.input_ids = [self.table[p] for p in prompts]
        return r


def canonical_tokens():
    g = torch.Generator().manual_seed(7)
    cond = torch.randint(0, 49408, (77,), generator=g)
    uncond = torch.randint(0, 49408, (77,), generator=g)
    return cond, uncond


def build_reference_models(ref):
    torch.manual_seed(0)
    enc = ref["encoder"].VAE_Encoder()
    dec = ref["decoder"].VAE_Decoder()
    dif = ref["diffusion"].Diffusion()
    clp = ref["clip"].CLIP()
    return {"encoder": enc.eval(), "decoder": dec.eval(), "diffusion": dif.eval(), "clip": clp.eval()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also run the 50-step txt2img and img2img references")
    ap.add_argument("--img2img50", action="store_true",
                    help="only config 3 as written: img2img on dog.jpg, strength 0.8, 50 nominal steps (40 UNet "
                         "evaluations) -> tests/golden/img2img_50.pt")
    args = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_grad_enabled(False)
    ref = load_reference()
    t0 = time.time()
    models = build_reference_models(ref)
    print(f"reference models built in {time.time() - t0:.1f}s", flush=True)

    # ---- the package's own classes must draw the identical seed-0 weights
    sys.path.insert(0, ROOT)
    from pytorch_stable_diffusion_b200.clip import CLIP
    from pytorch_stable_diffusion_b200.decoder import VAE_Decoder
    from pytorch_stable_diffusion_b200.diffusion import Diffusion
    from pytorch_stable_diffusion_b200.encoder import VAE_Encoder
    torch.manual_seed(0)
    mine = {"encoder": VAE_Encoder(), "decoder": VAE_Decoder(), "diffusion": Diffusion(), "clip": CLIP()}
    digest = {}
    for k in mine:
        a, b = mine[k].state_dict(), models[k].state_dict()
        assert list(a.keys()) == list(b.keys()), f"{k}: key order differs"
        for name in a:
            assert torch.equal(a[name], b[name]), f"{k}.{name}: seed-0 weights differ"
        digest[k] = {"n_keys": len(a), "n_params": sum(v.numel() for v in a.values()),
                     "sum": float(sum(v.double().sum() for v in a.values())),
                     "abs_sum": float(sum(v.double().abs().sum() for v in a.values()))}
    print("seed-0 weights identical between reference and package classes:", digest, flush=True)
    del mine

    cond, uncond = canonical_tokens()
    out = {"weights_digest": digest, "cond_tokens": cond, "uncond_tokens": uncond}

    if args.img2img50:
        # ---- config 3 as BASELINE.json words it: strength 0.8, 50 nominal DDPM steps = 40 UNet evaluations (t = 780 .. 0)
        from PIL import Image
        tok = StubTokenizer({"a": cond.tolist(), "b": uncond.tolist()})
        dog = Image.open("/root/reference/images/dog.jpg")
        dec_in = {}
        def grab(mod, inp):                 # must return None: a returned tensor would replace the input
            dec_in.setdefault("z", inp[0].clone())

        hd = models["decoder"].register_forward_pre_hook(grab)
        t0 = time.time()
        image = ref["pipeline"].generate(prompt="a", uncond_prompt="b", input_image=dog, strength=0.8, do_cfg=True,
                                         cfg_scale=7.5, sampler_name="ddpm", n_inference_steps=50, models=models,
                                         seed=42, device="cpu", idle_device=None, tokenizer=tok)
        dt = time.time() - t0
        hd.remove()
        print(f"reference img2img 50 nominal steps on CPU: {dt:.1f}s", flush=True)
        torch.save({"image": torch.from_numpy(image.copy()), "final_latents": dec_in["z"], "cpu_seconds": dt,
                    "threads": torch.get_num_threads(), "input_from": "img2img_5.pt['input'] (dog.jpg resized to 512x512)"},
                   os.path.join(GOLDEN, "img2img_50.pt"))
        print("img2img_50.pt written", flush=True)
        return

    # ---- small-block goldens (fast to re-check in the CPU suite)
    g = torch.Generator().manual_seed(123)
    torch.manual_seed(1)
    blk = {}
    rb = ref["diffusion"].UNET_ResidualBlock(64, 128).eval()
    x = torch.randn(2, 64, 8, 8, generator=g)
    t = torch.randn(1, 1280, generator=g)
    blk["unet_res"] = {"sd": rb.state_dict(), "x": x, "t": t, "y": rb(x, t)}
    ab = ref["diffusion"].UNET_AttentionBlock(2, 32).eval()
    x = torch.randn(2, 64, 8, 8, generator=g)
    c = torch.randn(2, 77, 768, generator=g)
    blk["unet_attn"] = {"sd": ab.state_dict(), "x": x, "ctx": c, "y": ab(x.clone(), c)}
    vr = ref["decoder"].VAE_ResidualBlock(64, 128).eval()
    x = torch.randn(1, 64, 8, 8, generator=g)
    blk["vae_res"] = {"sd": vr.state_dict(), "x": x, "y": vr(x.clone())}
    va = ref["decoder"].VAE_AttentionBlock(64).eval()
    x = torch.randn(2, 64, 8, 8, generator=g)
    blk["vae_attn"] = {"sd": va.state_dict(), "x": x, "y": va(x.clone())}
    cl = ref["clip"].CLIPLayer(4, 64).eval()
    x = torch.randn(2, 77, 64, generator=g)
    blk["clip_layer"] = {"sd": cl.state_dict(), "x": x, "y": cl(x.clone())}
    te = ref["pipeline"].get_time_embedding(980)
    blk["time_embedding_980"] = te
    smp = ref["ddpm"].DDPMSampler(torch.Generator().manual_seed(0))
    smp.set_inference_timesteps(50)
    blk["ddpm"] = {"timesteps": smp.timesteps.clone(), "alphas_cumprod": smp.alphas_cumprod.clone(),
                   "var_980": smp._get_variance(980).clone(), "var_0": smp._get_variance(0).clone()}
    lat = torch.randn(1, 4, 8, 8, generator=g)
    mo = torch.randn(1, 4, 8, 8, generator=g)
    smp.generator = torch.Generator().manual_seed(5)
    blk["ddpm"]["step_in"] = (lat, mo)
    blk["ddpm"]["step_980"] = smp.step(980, lat, mo)
    blk["ddpm"]["step_0"] = smp.step(0, lat, mo)
    smp.set_strength(0.8)
    blk["ddpm"]["strength08"] = (smp.start_step, smp.timesteps.clone())
    torch.save(blk, os.path.join(GOLDEN, "blocks.pt"))
    print("blocks.pt written", flush=True)

    # ---- full-size single evaluations with the canonical weights
    ctx = torch.cat([models["clip"](cond.view(1, -1)), models["clip"](uncond.view(1, -1))])
    out["context"] = ctx.clone()
    g = torch.Generator().manual_seed(11)
    lat = torch.randn(2, 4, 64, 64, generator=g)
    temb = ref["pipeline"].get_time_embedding(500)
    t0 = time.time()
    out["unet_eval"] = {"latent": lat, "t": 500, "y": models["diffusion"](lat, ctx, temb)}
    print(f"reference Diffusion.forward {time.time() - t0:.1f}s", flush=True)
    z = torch.randn(1, 4, 16, 16, generator=g)
    out["vae_decode_16"] = {"z": z, "y": models["decoder"](z.clone())}
    img = torch.rand(1, 3, 128, 128, generator=g) * 2 - 1
    nz = torch.randn(1, 4, 16, 16, generator=g)
    out["vae_encode_128"] = {"x": img, "noise": nz, "y": models["encoder"](img.clone(), nz)}
    torch.save(out, os.path.join(GOLDEN, "canonical.pt"))
    print("canonical.pt written", flush=True)

    if not args.full:
        return
    # ---- config 1: 50-step txt2img through the reference's own pipeline.generate on CPU
    tok = StubTokenizer({"a": cond.tolist(), "b": uncond.tolist()})
    trace = []
    keep = {0, 1, 24, 49}
    counter = {"i": 0}

    def hook(mod, inp, outp):
        i = counter["i"]
        if i in keep:
            trace.append({"step": i, "latent": inp[0].clone(), "time": inp[2].clone(), "y": outp.clone()})
        counter["i"] += 1

    h = models["diffusion"].register_forward_hook(hook)
    dec_in = {}
    def grab_decoder_input(mod, inp):     # must return None: a returned tensor would replace the input
        dec_in.setdefault("z", inp[0].clone())

    hd = models["decoder"].register_forward_pre_hook(grab_decoder_input)
    t0 = time.time()
    image = ref["pipeline"].generate(prompt="a", uncond_prompt="b", input_image=None, strength=0.8, do_cfg=True,
                                     cfg_scale=7.5, sampler_name="ddpm", n_inference_steps=50, models=models,
                                     seed=42, device="cpu", idle_device=None, tokenizer=tok)
    dt = time.time() - t0
    h.remove()
    hd.remove()
    print(f"reference txt2img 50 steps on CPU: {dt:.1f}s ({torch.get_num_threads()} threads)", flush=True)
    torch.save({"image": torch.from_numpy(image.copy()), "final_latents": dec_in["z"], "trace": trace,
                "cpu_seconds": dt, "threads": torch.get_num_threads()},
               os.path.join(GOLDEN, "txt2img_50.pt"))
    print("txt2img_50.pt written", flush=True)

    # ---- config 3: img2img on images/dog.jpg, strength 0.8 — 5 nominal steps (4 UNet evaluations)
    from PIL import Image
    dog = Image.open("/root/reference/images/dog.jpg")
    counter["i"] = 0
    dec_in.clear()
    hd = models["decoder"].register_forward_pre_hook(grab_decoder_input)
    image = ref["pipeline"].generate(prompt="a", uncond_prompt="b", input_image=dog, strength=0.8, do_cfg=True,
                                     cfg_scale=7.5, sampler_name="ddpm", n_inference_steps=5, models=models,
                                     seed=42, device="cpu", idle_device=None, tokenizer=tok)
    hd.remove()
    torch.save({"image": torch.from_numpy(image.copy()), "final_latents": dec_in["z"],
                "input": torch.from_numpy(np.array(dog.resize((512, 512))).copy())},
               os.path.join(GOLDEN, "img2img_5.pt"))
    print("img2img_5.pt written", flush=True)


if __name__ == "__main__":
    main()
