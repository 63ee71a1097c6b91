"""CompVis Stable Diffusion v1 checkpoint -> the four state_dicts of this package.

Same entry point and result as the reference's sd/model_converter.py:3-1056, but expressed as rules
derived from the module structure instead of 1042 literal assignments: every destination key is
produced by walking the target modules (UNET, VAE_Encoder, VAE_Decoder, CLIP) and naming the
CompVis tensor(s) it comes from. q/k/v projections are concatenated along dim 0 for the fused
in_proj weights (sd/model_converter.py:1009-1054), the VAE attention 1x1-conv kernels are reshaped
to matrices (:1025-1030).
"""
import torch
from torch import nn

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."
CLIP_PREFIX = "cond_stage_model.transformer.text_model."


def _wb(dst, src):
    return [(f"{dst}.weight", ("copy", [f"{src}.weight"])), (f"{dst}.bias", ("copy", [f"{src}.bias"]))]


def _unet_resblock(dst, src):
    r = []
    r += _wb(f"{dst}.groupnorm_feature", f"{src}.in_layers.0")
    r += _wb(f"{dst}.conv_feature", f"{src}.in_layers.2")
    r += _wb(f"{dst}.linear_time", f"{src}.emb_layers.1")
    r += _wb(f"{dst}.groupnorm_merged", f"{src}.out_layers.0")
    r += _wb(f"{dst}.conv_merged", f"{src}.out_layers.3")
    return r


def _unet_attnblock(dst, src):
    tb = f"{src}.transformer_blocks.0"
    r = []
    r += _wb(f"{dst}.groupnorm", f"{src}.norm")
    r += _wb(f"{dst}.conv_input", f"{src}.proj_in")
    for i in (1, 2, 3):
        r += _wb(f"{dst}.layernorm_{i}", f"{tb}.norm{i}")
    r.append((f"{dst}.attention_1.in_proj.weight",
              ("cat", [f"{tb}.attn1.to_{x}.weight" for x in "qkv"])))
    r += _wb(f"{dst}.attention_1.out_proj", f"{tb}.attn1.to_out.0")
    for x in "qkv":
        r.append((f"{dst}.attention_2.{x}_proj.weight", ("copy", [f"{tb}.attn2.to_{x}.weight"])))
    r += _wb(f"{dst}.attention_2.out_proj", f"{tb}.attn2.to_out.0")
    r += _wb(f"{dst}.linear_geglu_1", f"{tb}.ff.net.0.proj")
    r += _wb(f"{dst}.linear_geglu_2", f"{tb}.ff.net.2")
    r += _wb(f"{dst}.conv_output", f"{src}.proj_out")
    return r


def _unet_rules(diffusion):
    from .diffusion import UNET_AttentionBlock, UNET_ResidualBlock, Upsample
    P = UNET_PREFIX
    rules = []
    rules += _wb("time_embedding.linear_1", P + "time_embed.0")
    rules += _wb("time_embedding.linear_2", P + "time_embed.2")

    def walk(seq, dst_prefix, src_prefix):
        out = []
        for j, layer in enumerate(seq):
            dst, src = f"{dst_prefix}.{j}", f"{src_prefix}.{j}"
            if isinstance(layer, UNET_ResidualBlock):
                out += _unet_resblock(dst, src)
                if isinstance(layer.residual_layer, nn.Conv2d):
                    out += _wb(f"{dst}.residual_layer", f"{src}.skip_connection")
            elif isinstance(layer, UNET_AttentionBlock):
                out += _unet_attnblock(dst, src)
            elif isinstance(layer, Upsample):
                out += _wb(f"{dst}.conv", f"{src}.conv")
            elif isinstance(layer, nn.Conv2d):
                out += _wb(dst, f"{src}.op" if layer.stride[0] == 2 else src)
        return out

    for i, seq in enumerate(diffusion.unet.encoders):
        rules += walk(seq, f"unet.encoders.{i}", f"{P}input_blocks.{i}")
    rules += walk(diffusion.unet.bottleneck, "unet.bottleneck", f"{P}middle_block")
    for i, seq in enumerate(diffusion.unet.decoders):
        rules += walk(seq, f"unet.decoders.{i}", f"{P}output_blocks.{i}")
    rules += _wb("final.groupnorm", P + "out.0")
    rules += _wb("final.conv", P + "out.2")
    return rules


def _vae_resblock(dst, src, has_shortcut):
    r = []
    for i in (1, 2):
        r += _wb(f"{dst}.groupnorm_{i}", f"{src}.norm{i}")
        r += _wb(f"{dst}.conv_{i}", f"{src}.conv{i}")
    if has_shortcut:
        r += _wb(f"{dst}.residual_layer", f"{src}.nin_shortcut")
    return r


def _vae_attn(dst, src):
    return [
        (f"{dst}.groupnorm.weight", ("copy", [f"{src}.norm.weight"])),
        (f"{dst}.groupnorm.bias", ("copy", [f"{src}.norm.bias"])),
        (f"{dst}.attention.in_proj.weight", ("cat_matrix", [f"{src}.{x}.weight" for x in "qkv"])),
        (f"{dst}.attention.in_proj.bias", ("cat", [f"{src}.{x}.bias" for x in "qkv"])),
        (f"{dst}.attention.out_proj.weight", ("matrix", [f"{src}.proj_out.weight"])),
        (f"{dst}.attention.out_proj.bias", ("copy", [f"{src}.proj_out.bias"])),
    ]


def _vae_rules(module, side):
    """Walks the nn.Sequential of VAE_Encoder / VAE_Decoder and names the CompVis source of every
    entry (encoder: conv_in, down.{0..3}.block.{0,1} (+downsample), mid, norm_out, conv_out,
    quant_conv; decoder: post_quant_conv, conv_in, mid, up.{3..0}.block.{0..2} (+upsample), ...)."""
    from .decoder import VAE_AttentionBlock, VAE_ResidualBlock
    P = VAE_PREFIX + side
    entries = list(module)
    rules = []
    if side == "encoder":
        names, level, blk = [], 0, 0
        names.append(f"{P}.conv_in")
        i = 1
        while not isinstance(entries[i], VAE_AttentionBlock) and i < len(entries):
            e = entries[i]
            if isinstance(e, VAE_ResidualBlock):
                names.append(f"{P}.down.{level}.block.{blk}")
                blk += 1
            elif isinstance(e, nn.Conv2d):
                names.append(f"{P}.down.{level}.downsample.conv")
                level, blk = level + 1, 0
            i += 1
            if level == 3 and blk == 2:
                break
        names += [f"{P}.mid.block_1", f"{P}.mid.attn_1", f"{P}.mid.block_2", f"{P}.norm_out", None,
                  f"{P}.conv_out", VAE_PREFIX + "quant_conv"]
    else:
        names = [VAE_PREFIX + "post_quant_conv", f"{P}.conv_in", f"{P}.mid.block_1", f"{P}.mid.attn_1",
                 f"{P}.mid.block_2"]
        level, blk = 3, 0
        for e in entries[5:-3]:
            if isinstance(e, VAE_ResidualBlock):
                names.append(f"{P}.up.{level}.block.{blk}")
                blk += 1
            elif isinstance(e, nn.Upsample):
                names.append(None)
            elif isinstance(e, nn.Conv2d):
                names.append(f"{P}.up.{level}.upsample.conv")
                level, blk = level - 1, 0
        names += [f"{P}.norm_out", None, f"{P}.conv_out"]
    assert len(names) == len(entries), (len(names), len(entries))
    for idx, (e, src) in enumerate(zip(entries, names)):
        if src is None:
            continue
        if isinstance(e, VAE_ResidualBlock):
            rules += _vae_resblock(str(idx), src, isinstance(e.residual_layer, nn.Conv2d))
        elif isinstance(e, VAE_AttentionBlock):
            rules += _vae_attn(str(idx), src)
        else:
            rules += _wb(str(idx), src)
    return rules


def _clip_rules(clip):
    P = CLIP_PREFIX
    rules = [("embedding.token_embedding.weight", ("copy", [P + "embeddings.token_embedding.weight"])),
             ("embedding.position_embedding", ("copy", [P + "embeddings.position_embedding.weight"]))]
    for i in range(len(clip.layers)):
        s, d = f"{P}encoder.layers.{i}", f"layers.{i}"
        rules.append((f"{d}.attention.in_proj.weight", ("cat", [f"{s}.self_attn.{x}_proj.weight" for x in "qkv"])))
        rules.append((f"{d}.attention.in_proj.bias", ("cat", [f"{s}.self_attn.{x}_proj.bias" for x in "qkv"])))
        rules += _wb(f"{d}.attention.out_proj", f"{s}.self_attn.out_proj")
        rules += _wb(f"{d}.layernorm_1", f"{s}.layer_norm1")
        rules += _wb(f"{d}.layernorm_2", f"{s}.layer_norm2")
        rules += _wb(f"{d}.linear_1", f"{s}.mlp.fc1")
        rules += _wb(f"{d}.linear_2", f"{s}.mlp.fc2")
    rules += _wb("layernorm", P + "final_layer_norm")
    return rules


_RULES = None


def conversion_rules():
    """{'diffusion'|'encoder'|'decoder'|'clip': [(dst_key, (op, [src_keys]))]}; op in copy, cat,
    cat_matrix (cat then reshape to 2-D), matrix (reshape to 2-D)."""
    global _RULES
    if _RULES is None:
        from .clip import CLIP
        from .decoder import VAE_Decoder
        from .diffusion import Diffusion
        from .encoder import VAE_Encoder
        with torch.device("meta"):
            diffusion, enc, dec, clip = Diffusion(), VAE_Encoder(), VAE_Decoder(), CLIP()
        _RULES = {"diffusion": _unet_rules(diffusion), "encoder": _vae_rules(enc, "encoder"),
                  "decoder": _vae_rules(dec, "decoder"), "clip": _clip_rules(clip)}
    return _RULES


def convert_state_dict(original_model):
    converted = {}
    for group, rules in conversion_rules().items():
        out = {}
        for dst, (op, srcs) in rules:
            if op == "copy":
                t = original_model[srcs[0]]
            elif op == "cat":
                t = torch.cat([original_model[s] for s in srcs], 0)
            elif op == "cat_matrix":
                t = torch.cat([original_model[s] for s in srcs], 0)
                t = t.reshape(t.shape[0], -1)
            elif op == "matrix":
                t = original_model[srcs[0]]
                t = t.reshape(t.shape[0], -1)
            else:
                raise ValueError(op)
            out[dst] = t
        converted[group] = out
    return converted


def load_from_standard_weights(input_file: str, device: str, allow_unsafe_pickle: bool = None) -> dict:
    """Reads a CompVis .ckpt and returns {'diffusion','encoder','decoder','clip'} state_dicts
    (sd/model_converter.py:3).

    The file is unpickled with weights_only=True. The reference loads with weights_only=False
    (sd/model_converter.py:5), which executes whatever the pickle contains; that behaviour is opt-in here:
    allow_unsafe_pickle=True, or SDB_ALLOW_UNSAFE_PICKLE=1 in the environment, for legacy checkpoints that pickle
    arbitrary classes - only for files you trust."""
    import os
    import pickle
    if allow_unsafe_pickle is None:
        allow_unsafe_pickle = os.environ.get("SDB_ALLOW_UNSAFE_PICKLE") == "1"
    try:
        ckpt = torch.load(input_file, map_location=device, weights_only=True)
    except pickle.UnpicklingError as e:
        if not allow_unsafe_pickle:
            raise pickle.UnpicklingError(
                f"{input_file} cannot be read with weights_only=True ({e}). If you trust the file, pass "
                "allow_unsafe_pickle=True (or set SDB_ALLOW_UNSAFE_PICKLE=1) to unpickle it the way the reference "
                "does - this executes code stored in the checkpoint.") from e
        ckpt = torch.load(input_file, map_location=device, weights_only=False)
    return convert_state_dict(ckpt["state_dict"])
