"""Max-norm and RMS relative error of one UNet evaluation against the fp32 oracle on the same device, at several
latent sizes - to see whether a change moves the error systematically or only the single worst element.
usage: [SDB_NO_FOLD_FF_OUT=1 | SDB_NO_SUM_ROW=1 ...] python tools/diag_fold_error.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sd_oracle  # noqa: E402
from pytorch_stable_diffusion_b200 import pipeline, synthetic  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
models = synthetic.build_models(dev, which=("diffusion",))
sd = synthetic.state_dicts(models)["diffusion"]
g = torch.Generator().manual_seed(11)
ctx = torch.randn(2, 77, 768, generator=g).to(dev)
CASES = ((96, 980), (96, 500), (96, 20), (96, 740), (48, 980)) if "--96" in sys.argv else \
    ((32, 980), (64, 980), (64, 500), (64, 20), (96, 980))
for hw, t in CASES:
    lat = torch.randn(1, 4, hw, hw, generator=g).to(dev).repeat(2, 1, 1, 1)
    temb = pipeline.get_time_embedding(t).to(dev)
    with torch.no_grad():
        got = models["diffusion"](lat, ctx, temb)
        ref = sd_oracle.diffusion_forward(sd, lat, ctx, temb)
    err = (got - ref).float()
    print(f"latent {hw}x{hw} t={t}: max|err|/max|ref| = {float(err.abs().max() / ref.abs().max()):.3e}   "
          f"rms(err)/rms(ref) = {float(err.pow(2).mean().sqrt() / ref.float().pow(2).mean().sqrt()):.3e}", flush=True)
