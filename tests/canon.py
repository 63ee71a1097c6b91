"""Canonical synthetic inputs shared by the tests + golden-fixture loader."""
import os

import torch

from pytorch_stable_diffusion_b200.synthetic import (StubTokenizer, build_models, canonical_tokens,  # noqa: F401
                                                     state_dicts)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    path = os.path.join(GOLDEN_DIR, name)
    if not os.path.exists(path):
        return None
    return torch.load(path, map_location="cpu", weights_only=False)


def compvis_checkpoint_from(sds):
    """The inverse of model_converter.convert_state_dict: a CompVis-layout {'state_dict': ...} holding the given
    {'diffusion','encoder','decoder','clip'} state_dicts (fused q/k/v matrices split back into their three sources,
    the two VAE attention matrices back to 1x1-conv shape) - a full-size synthetic v1-5-pruned-emaonly.ckpt."""
    from pytorch_stable_diffusion_b200 import model_converter
    out = {}
    for group, rules in model_converter.conversion_rules().items():
        for dst, (op, srcs) in rules:
            t = sds[group][dst].detach().cpu()
            if op == "copy":
                out[srcs[0]] = t.clone()
            elif op == "matrix":
                out[srcs[0]] = t.reshape(t.shape[0], -1, 1, 1).clone()
            elif op in ("cat", "cat_matrix"):
                for s, part in zip(srcs, t.chunk(len(srcs), dim=0)):
                    out[s] = (part.reshape(part.shape[0], -1, 1, 1) if op == "cat_matrix" else part).clone()
            else:
                raise ValueError(op)
    return {"state_dict": out}
