"""Builds the four models from a standard checkpoint (sd/model_loader.py:9-50)."""
from . import model_converter
from .clip import CLIP
from .decoder import VAE_Decoder
from .diffusion import Diffusion
from .encoder import VAE_Encoder


def preload_models_from_standard_weights(ckpt_path, device):
    state_dict = model_converter.load_from_standard_weights(ckpt_path, device)
    return models_from_state_dicts(state_dict, device)


def models_from_state_dicts(state_dict, device):
    """{'encoder','decoder','diffusion','clip'} state_dicts -> the four modules (strict load), as the reference's
    preload_models_from_standard_weights builds them (sd/model_loader.py:28-50)."""

    encoder = VAE_Encoder().to(device)
    encoder.load_state_dict(state_dict['encoder'], strict=True)

    decoder = VAE_Decoder().to(device)
    decoder.load_state_dict(state_dict['decoder'], strict=True)

    diffusion = Diffusion().to(device)
    diffusion.load_state_dict(state_dict['diffusion'], strict=True)

    clip = CLIP().to(device)
    clip.load_state_dict(state_dict['clip'], strict=True)

    return {
        'clip': clip,
        'encoder': encoder,
        'decoder': decoder,
        'diffusion': diffusion,
    }
