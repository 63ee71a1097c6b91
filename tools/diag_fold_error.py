"""Max-norm and RMS relative error of one UNet evaluation against the fp32 oracle on the same device, at several
latent sizes - to see whether a change moves the error systematically or only the single worst element.
usage: [SDB_NO_FOLD_FF_OUT=1 | SDB_NO_SUM_ROW=1 ...] python tools/diag_fold_error.py [--96] [--final-fp32]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sd_oracle  # noqa: E402
from pytorch_stable_diffusion_b200 import pipeline, synthetic  # noqa: E402

if "--final-fp32" in sys.argv:
    # diagnostic: the LAST layer (GroupNorm + SiLU + conv 320->4) evaluated by PyTorch in fp32 on the stream the
    # kernels produced - how much of the output error is the last layer's operand rounding? (test infrastructure:
    # the package itself has no such path)
    import torch.nn.functional as F
    from pytorch_stable_diffusion_b200 import engine, ops

    def _forward_fp32_tail(self, x, tvec, kvs):
        kv_iter = iter(kvs)
        skips = []
        for prog in self.encoders:
            x = self._run_seq(prog, x, None, tvec, kv_iter)
            skips.append(x)
        x = self._run_seq(self.bottleneck, x, None, tvec, kv_iter)
        for prog in self.decoders:
            x = self._run_seq(prog, x, skips.pop(), tvec, kv_iter)
        gw, gb, cw, cb = self._tail
        a32 = F.silu(F.group_norm(x.f.permute(0, 3, 1, 2), 32, gw, gb, 1e-5))
        return F.conv2d(a32, cw, cb, padding=1).permute(0, 2, 3, 1).contiguous()

    engine.UNetEngine.forward_nhwc = _forward_fp32_tail

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
models = synthetic.build_models(dev, which=("diffusion",))
sd = synthetic.state_dicts(models)["diffusion"]
if "--final-fp32" in sys.argv:
    fin = models["diffusion"].final
    models["diffusion"]._engine()._tail = tuple(t.detach().float().to(dev) for t in (
        fin.groupnorm.weight, fin.groupnorm.bias, fin.conv.weight, fin.conv.bias))
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
