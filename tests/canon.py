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
