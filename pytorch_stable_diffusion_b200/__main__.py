"""Command-line front end replacing the reference's demo notebook (sd/inference_demo.ipynb:99-118).

    python -m pytorch_stable_diffusion_b200 --ckpt v1-5-pruned-emaonly.ckpt --vocab vocab.json --merges merges.txt \
        --prompt "a dog wearing a hat" --out dog.png [--image in.jpg --strength 0.8] [--batch 4 --seed 42]
    python -m pytorch_stable_diffusion_b200 --synthetic --prompt a --uncond b --steps 4 --out noise.png   # no files

--synthetic runs the canonical random-init weights and the stub tokenizer (prompts "a" / "b"): a wiring check for
machines without the SD-1.5 checkpoint. Images are written as PNG (or .npy when Pillow cannot encode the suffix)."""
import argparse
import sys

import numpy as np
import torch


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m pytorch_stable_diffusion_b200", description=__doc__.split("\n\n")[0])
    ap.add_argument("--prompt", required=True)
    ap.add_argument("--uncond", default="")
    ap.add_argument("--ckpt")
    ap.add_argument("--vocab")
    ap.add_argument("--merges")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--image", help="input image for img2img")
    ap.add_argument("--strength", type=float, default=0.8)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--cfg", type=float, default=7.5)
    ap.add_argument("--no-cfg", action="store_true")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=512)
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--allow-unsafe-pickle", action="store_true", help="load a legacy checkpoint with weights_only=False")
    ap.add_argument("--out", required=True, help="output path; with --batch N > 1 an index is inserted before the suffix")
    a = ap.parse_args(argv)

    from . import model_converter, model_loader, pipeline, synthetic
    from .tokenizer import CLIPTokenizer
    if a.synthetic:
        models = synthetic.build_models(a.device)
        tok = synthetic.StubTokenizer()
    else:
        if not (a.ckpt and a.vocab and a.merges):
            ap.error("--ckpt, --vocab and --merges are required (or --synthetic)")
        tok = CLIPTokenizer(a.vocab, a.merges)
        if a.allow_unsafe_pickle:
            sd = model_converter.load_from_standard_weights(a.ckpt, a.device, allow_unsafe_pickle=True)
            models = model_loader.models_from_state_dicts(sd, a.device)
        else:
            models = model_loader.preload_models_from_standard_weights(a.ckpt, a.device)
    image = None
    if a.image:
        from PIL import Image
        image = Image.open(a.image).convert("RGB")
    imgs = pipeline.generate(a.prompt, a.uncond, input_image=image, strength=a.strength, do_cfg=not a.no_cfg,
                             cfg_scale=a.cfg, n_inference_steps=a.steps, models=models, device=a.device, tokenizer=tok,
                             batch_size=a.batch, seeds=[a.seed + i for i in range(a.batch)], height=a.height,
                             width=a.width, return_all=True)
    stem, dot, suffix = a.out.rpartition(".")
    for i, im in enumerate(imgs):
        path = a.out if a.batch == 1 else (f"{stem}_{i}.{suffix}" if dot else f"{a.out}_{i}")
        try:
            from PIL import Image
            Image.fromarray(im).save(path)
        except (ImportError, ValueError, OSError):
            path = path + ".npy"
            np.save(path, im)
        print(path)
    return 0


if __name__ == "__main__":
    sys.exit(main())
