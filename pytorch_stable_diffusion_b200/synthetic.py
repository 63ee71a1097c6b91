"""Canonical synthetic inputs: SD-1.5-architecture random-init weights, synthetic CLIP
token ids and a stub tokenizer. Checkpoints and the CLIP vocabulary are not available offline, so
tests, smoke() and bench.py all run on these.

  weights  torch.manual_seed(0); VAE_Encoder(), VAE_Decoder(), Diffusion(), CLIP() constructed in the
           reference loader's order (sd/model_loader.py:28-41) with PyTorch default init on the CPU
  tokens   torch.Generator().manual_seed(7): cond = randint(0, 49408, (77,)), uncond = the next draw
"""
import torch


def canonical_tokens():
    g = torch.Generator().manual_seed(7)
    cond = torch.randint(0, 49408, (77,), generator=g)
    uncond = torch.randint(0, 49408, (77,), generator=g)
    return cond, uncond


class StubTokenizer:
    """The one method pipeline.generate calls (sd/pipeline.py:109): fixed ids per prompt string."""

    def __init__(self, table=None):
        if table is None:
            cond, uncond = canonical_tokens()
            table = {"a": cond.tolist(), "b": uncond.tolist()}
        self.table = table

    def batch_encode_plus(self, prompts, padding=None, max_length=None):
        class _R:
            pass
        r = _R()
        r.input_ids = [self.table[p] for p in prompts]
        return r


def build_models(device="cpu", which=("encoder", "decoder", "diffusion", "clip")):
    """Seed-0 random-init models in the reference loader's construction order. `which` limits what is
    kept (all four are still constructed so that every kept model has its canonical weights)."""
    from .clip import CLIP
    from .decoder import VAE_Decoder
    from .diffusion import Diffusion
    from .encoder import VAE_Encoder
    torch.manual_seed(0)
    models = {}
    for name, ctor in (("encoder", VAE_Encoder), ("decoder", VAE_Decoder), ("diffusion", Diffusion),
                       ("clip", CLIP)):
        m = ctor()
        if name in which:
            m.eval()
            models[name] = m.to(device) if device != "cpu" else m
        else:
            del m
    return models


def state_dicts(models, device=None):
    return {k: {n: (t.detach().to(device) if device else t.detach()) for n, t in m.state_dict().items()}
            for k, m in models.items()}
