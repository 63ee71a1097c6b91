"""CLIP text encoder with the reference's classes and state_dict keys (sd/clip.py:7-261)."""
import torch
from torch import nn

from . import engine
from .attention import SelfAttention, _require_cuda


class CLIPEmbedding(nn.Module):
    def __init__(self, vocab_size: int, embedding_dim: int, max_seq_length: int):
        super().__init__()
        self.token_embedding = nn.Embedding(vocab_size, embedding_dim)
        self.position_embedding = nn.Parameter(torch.zeros((max_seq_length, embedding_dim)))

    def forward(self, tokens: torch.LongTensor) -> torch.FloatTensor:
        """token gather + learned positions (sd/clip.py:38-66) -> (B, T, D) fp32."""
        _require_cuda(tokens, "CLIPEmbedding")
        from . import ops
        n, t = tokens.shape
        out = ops.clip_embed(tokens.contiguous(), self.token_embedding.weight.detach().float().contiguous(),
                             self.position_embedding.detach().float().contiguous(), t)
        return out.float()


class CLIPLayer(nn.Module):
    def __init__(self, n_heads: int, n_embed: int):
        super().__init__()
        self.layernorm_1 = nn.LayerNorm(n_embed)
        self.attention = SelfAttention(n_heads, n_embed)
        self.layernorm_2 = nn.LayerNorm(n_embed)
        self.linear_1 = nn.Linear(n_embed, 4 * n_embed)
        self.linear_2 = nn.Linear(4 * n_embed, n_embed)

    def forward(self, x):
        """pre-LN causal self-attention + quick-GELU MLP (sd/clip.py:123-176); x (B, T, D) fp32."""
        _require_cuda(x, "CLIPLayer")
        from types import SimpleNamespace as NS
        from . import ops
        dev = x.device
        pk = NS()
        pk.ln1 = engine.pack_norm(self.layernorm_1, dev)
        pk.att = engine.pack_self_attention(self.attention, dev)
        pk.ln2 = engine.pack_norm(self.layernorm_2, dev)
        pk.w1, pk.b1 = engine.pack_linear(self.linear_1, dev)
        pk.w2, pk.b2 = engine.pack_linear(self.linear_2, dev)
        n, t, d = x.shape
        t_pad = (t + 7) // 8 * 8
        xp = torch.zeros((n, t_pad, d), device=dev, dtype=torch.float32)
        xp[:, :t] = x
        y = engine.run_clip_layer(pk, xp.view(n * t_pad, d), n, t_pad)
        return y.view(n, t_pad, d)[:, :t].contiguous()


class CLIP(nn.Module, engine.EngineCache):
    _engine_cls = engine.CLIPEngine

    def __init__(self):
        super().__init__()
        self.embedding = CLIPEmbedding(49408, 768, 77)
        self.layers = nn.ModuleList([CLIPLayer(12, 768) for i in range(12)])
        self.layernorm = nn.LayerNorm(768)

    def forward(self, tokens: torch.LongTensor) -> torch.FloatTensor:
        """tokens (B, 77) int64 -> (B, 77, 768) fp32 (sd/clip.py:227-261)."""
        _require_cuda(tokens, "CLIP")
        return self._engine().forward(tokens.type(torch.long))
