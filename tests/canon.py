"""Canonical synthetic inputs shared by tests, smoke() and bench.py."""
import os

import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


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


def build_models(device="cpu"):
    """torch.manual_seed(0); VAE_Encoder(), VAE_Decoder(), Diffusion(), CLIP() in the reference
    loader's order (sd/model_loader.py:28-41) with PyTorch default init."""
    from pytorch_stable_diffusion_b200.clip import CLIP
    from pytorch_stable_diffusion_b200.decoder import VAE_Decoder
    from pytorch_stable_diffusion_b200.diffusion import Diffusion
    from pytorch_stable_diffusion_b200.encoder import VAE_Encoder
    torch.manual_seed(0)
    models = {"encoder": VAE_Encoder(), "decoder": VAE_Decoder(), "diffusion": Diffusion(), "clip": CLIP()}
    for m in models.values():
        m.eval()
        if device != "cpu":
            m.to(device)
    return models


def state_dicts(models, device=None):
    return {k: {n: (t.detach().to(device) if device else t.detach()) for n, t in m.state_dict().items()}
            for k, m in models.items()}


def golden(name):
    path = os.path.join(GOLDEN_DIR, name)
    if not os.path.exists(path):
        return None
    return torch.load(path, map_location="cpu", weights_only=False)
