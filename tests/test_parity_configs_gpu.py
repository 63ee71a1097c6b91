"""Parity AT THE CONFIGURATIONS bench.py MEASURES (BASELINE.json configs[1]..[4]), against the oracle and the
reference's own golden outputs.

  configs[1]/[3]  one Diffusion.forward at N = 16 (8 CFG pairs, 64x64 latents: the wide-tile / split-K / multi-tile
                  choices of ops._choose_tiling at rows = 65 536), and generate(batch_size=8, seeds=42..49), 50 steps,
                  every image against the oracle's own 50-step run with the same seed
  configs[2]      img2img on dog.jpg, strength 0.8, 50 nominal steps (40 UNet evaluations) vs the reference's CPU image
  configs[4]      768x768: four noise samples x t in {980, 500, 20} at N = 2 and one N = 16 evaluation

Tolerance (north_star): per UNet evaluation max|y - y_ref| / max|y_ref| <= 1e-2; images PSNR >= 35 dB. The 768^2 cases
must hold the per-evaluation bound with >= 10 % margin (MARGIN_768).
"""
import os
import sys

import pytest
import torch

from canon import StubTokenizer, build_models, canonical_tokens, golden, state_dicts
from gpu_util import rel_err, setup_exact_fp32

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-2
MARGIN_768 = 0.9e-2


@pytest.fixture(scope="module")
def oracle():
    import sd_oracle
    setup_exact_fp32()
    return sd_oracle


@pytest.fixture(scope="module")
def models():
    return build_models(DEV)


@pytest.fixture(scope="module")
def weights(models):
    return state_dicts(models, DEV)


@pytest.fixture(scope="module")
def context(models):
    """(2, 77, 768): the canonical cond / uncond prompts through CLIP (reference golden when present)."""
    g = golden("canonical.pt")
    if g is not None:
        return g["context"].to(DEV)
    cond, uncond = canonical_tokens()
    with torch.no_grad():
        return models["clip"](torch.stack([cond, uncond]).to(DEV))


def _eval_pairs(models, weights, oracle, context, lat, t, name, tol):
    """One Diffusion.forward over B CFG pairs in bench.py's batch layout ([cond x B ; uncond x B]) against the oracle
    evaluated pair by pair in fp32 on the same device. Returns the per-pair errors."""
    from pytorch_stable_diffusion_b200 import _ext
    from pytorch_stable_diffusion_b200.pipeline import get_time_embedding
    B = lat.shape[0]
    temb = get_time_embedding(t).to(DEV)
    x = lat.repeat(2, 1, 1, 1)
    ctx = torch.cat([context[0:1].expand(B, -1, -1), context[1:2].expand(B, -1, -1)]).contiguous()
    errs = []
    with torch.no_grad():
        got = models["diffusion"](x, ctx, temb)
        for i in range(B):
            ref = oracle.diffusion_forward(weights["diffusion"], torch.stack([lat[i], lat[i]]), context, temb)
            errs.append(rel_err(torch.stack([got[i], got[B + i]]), ref))
            del ref
    assert _ext.read_fault() == 0
    print(f"[{name}] t={t} N={2 * B} per-pair rel_err: " + " ".join(f"{e:.2e}" for e in errs) +
          f"  max={max(errs):.3e} (tol {tol:.1e})", flush=True)
    return errs


# ------------------------------------------------------------------------------- configs[1] / configs[3]: N = 16
@pytest.mark.parametrize("t", [980, 500, 20])
def test_unet_eval_at_benched_batch(models, weights, oracle, context, t):
    gens = [torch.Generator().manual_seed(42 + i) for i in range(8)]
    lat = torch.cat([torch.randn((1, 4, 64, 64), generator=g) for g in gens]).to(DEV)
    errs = _eval_pairs(models, weights, oracle, context, lat, t, "Diffusion 64x64, benched batch", TOL)
    assert max(errs) <= TOL


def test_generate_batch8_50_steps_vs_oracle(models, weights, oracle):
    """The benchmarked call itself: generate(batch_size=8), 50 steps, CUDA-graph loop. Every image against the
    oracle's 50-step run with the same seed (fp32, same device); sample 0 also against the reference's own CPU image."""
    from pytorch_stable_diffusion_b200 import pipeline
    cond, uncond = canonical_tokens()
    seeds = list(range(42, 50))
    imgs = pipeline.generate("a", "b", models=models, seeds=seeds, batch_size=8, n_inference_steps=50, device=DEV,
                             tokenizer=StubTokenizer(), return_all=True)
    assert imgs.shape == (8, 512, 512, 3) and imgs.dtype.name == "uint8"
    psnrs = []
    for i, s in enumerate(seeds):
        ref_img, _ = oracle.generate(weights, cond, uncond, seed=s, n_inference_steps=50, device=DEV)
        psnrs.append(oracle.psnr_u8(imgs[i], ref_img))
    print("[generate B=8, 50 steps] PSNR vs oracle per seed: " + " ".join(f"{p:.1f}" for p in psnrs), flush=True)
    assert min(psnrs) >= 35.0
    g = golden("txt2img_50.pt")
    if g is not None:
        p = oracle.psnr_u8(imgs[0], g["image"].numpy())
        print(f"[generate B=8, 50 steps] sample 0 vs the reference's CPU image: {p:.2f} dB", flush=True)
        assert p >= 35.0


def test_graph_replay_equals_eager(models):
    """The captured loop launches the same kernels with the same arguments as the eager loop: byte-identical images."""
    from pytorch_stable_diffusion_b200 import pipeline
    kw = dict(models=models, seeds=[42, 43], batch_size=2, n_inference_steps=4, device=DEV, tokenizer=StubTokenizer(),
              return_all=True)
    a = pipeline.generate("a", "b", **kw)
    b = pipeline.generate("a", "b", use_cuda_graph=False, **kw)
    c = pipeline.generate("a", "b", **kw)
    diff = int(abs(a.astype(int) - b.astype(int)).max())
    print(f"[graph vs eager] max |diff| = {diff}; replay twice: {int(abs(a.astype(int) - c.astype(int)).max())}", flush=True)
    assert (a == c).all(), "two replays of the same graph differ"
    assert (a == b).all(), "graph replay differs from eager execution"


# ------------------------------------------------------------------------------- configs[2]: img2img, 50 nominal steps
def test_img2img_50_steps_vs_reference_golden(models, oracle):
    g = golden("img2img_50.pt")
    g5 = golden("img2img_5.pt")
    if g is None or g5 is None:
        pytest.skip("img2img_50 golden not generated (python oracle/make_golden.py --img2img50)")
    from PIL import Image
    from pytorch_stable_diffusion_b200 import pipeline
    dog = Image.fromarray(g5["input"].numpy())
    img = pipeline.generate("a", "b", input_image=dog, strength=0.8, models=models, seeds=[42],
                            n_inference_steps=50, device=DEV, tokenizer=StubTokenizer())
    p = oracle.psnr_u8(img, g["image"].numpy())
    print(f"[img2img strength 0.8, 50 nominal steps = 40 UNet evaluations] PSNR vs reference = {p:.2f} dB", flush=True)
    assert p >= 35.0


# ------------------------------------------------------------------------------- configs[4]: 768x768
@pytest.mark.parametrize("seed", [31, 32, 33, 34])
def test_768_unet_eval_margin(models, weights, oracle, context, seed):
    gen = torch.Generator().manual_seed(seed)
    lat = torch.randn(1, 4, 96, 96, generator=gen).to(DEV)
    worst = 0.0
    for t in (980, 500, 20):
        worst = max(worst, max(_eval_pairs(models, weights, oracle, context, lat, t, f"Diffusion 96x96 seed {seed}",
                                           MARGIN_768)))
    assert worst <= MARGIN_768


def test_768_unet_eval_at_benched_batch(models, weights, oracle, context):
    gens = [torch.Generator().manual_seed(142 + i) for i in range(8)]
    lat = torch.cat([torch.randn((1, 4, 96, 96), generator=g) for g in gens]).to(DEV)
    errs = _eval_pairs(models, weights, oracle, context, lat, 640, "Diffusion 96x96, benched batch", MARGIN_768)
    assert max(errs) <= MARGIN_768


def test_768_generate_batch8_vs_oracle(models, weights, oracle):
    """configs[4] through generate(): batch 8 at 768x768, 6 steps (the oracle's 9216-token fp32 attention makes 50
    steps x 8 seeds minutes of GPU time); two of the eight images against the oracle's run with the same seed."""
    from pytorch_stable_diffusion_b200 import pipeline
    cond, uncond = canonical_tokens()
    seeds = list(range(42, 50))
    imgs = pipeline.generate("a", "b", models=models, seeds=seeds, batch_size=8, n_inference_steps=6, device=DEV,
                             tokenizer=StubTokenizer(), return_all=True, height=768, width=768)
    assert imgs.shape == (8, 768, 768, 3)
    for i in (0, 7):
        ref_img, _ = oracle.generate(weights, cond, uncond, seed=seeds[i], n_inference_steps=6, latent_hw=(96, 96),
                                     device=DEV)
        p = oracle.psnr_u8(imgs[i], ref_img)
        print(f"[768^2 generate B=8, 6 steps] seed {seeds[i]} PSNR vs oracle = {p:.2f} dB", flush=True)
        assert p >= 35.0
