"""SelfAttention / CrossAttention with the reference's constructor signatures and parameter names
(sd/attention.py:6-253). The parameters live in ordinary nn.Linear containers so state_dicts load
unchanged; forward() runs the tcgen05 projection GEMMs and the flash-attention kernel."""
import torch
from torch import nn

from . import engine, ops


def _require_cuda(x, who):
    if not x.is_cuda:
        raise RuntimeError(f"{who} runs on hand-written CUDA kernels only (no CPU fallback); got a {x.device} tensor")


class _Packed:
    """Per-module cache of packed parameters, rebuilt when any parameter changes."""

    def _packed(self, builder):
        fp = engine.fingerprint(self)
        c = self.__dict__.get("_sdb_pack")
        if c is None or c[0] != fp:
            dev = next(self.parameters()).device
            c = (fp, builder(self, dev))
            self.__dict__["_sdb_pack"] = c
        return c[1]


class SelfAttention(nn.Module, _Packed):
    def __init__(self, n_heads: int, d_embed: int, in_proj_bias: bool = True, out_proj_bias: bool = True):
        super().__init__()
        self.in_proj = nn.Linear(d_embed, d_embed * 3, bias=in_proj_bias)
        self.out_proj = nn.Linear(d_embed, d_embed, bias=out_proj_bias)
        self.n_heads = n_heads
        self.d_head = d_embed // n_heads

    def forward(self, x, causal_mask: bool = False):
        """x: (B, S, E) fp32 -> (B, S, E) fp32 (sd/attention.py:27-93)."""
        _require_cuda(x, "SelfAttention")
        pk = self._packed(engine.pack_self_attention)
        b, s, e = x.shape
        if self.d_head > 160:
            raise ValueError("SelfAttention kernel path needs d_head <= 160; the d=512 VAE attention "
                             "runs through VAE_AttentionBlock")
        xb = x.to(torch.bfloat16).contiguous().view(b * s, e)
        qk = ops.linear(xb, pk.wqk, bias=pk.bqk)
        vt, vt_ld = engine.project_vt(pk.wv, pk.bv, xb, b, s)
        o = torch.empty_like(xb)
        ops.attention(qk, qk[:, e:], vt, o, NB=b, heads=self.n_heads, d=self.d_head, S=s, Skv=s,
                      Skv_pad=s, vt_ld=vt_ld, ldq=2 * e, ldk=2 * e, ldo=e, causal=causal_mask)
        out = ops.linear(o, pk.wo, bias=pk.bo, out_fp32=True)
        return out.view(b, s, e)


class CrossAttention(nn.Module, _Packed):
    def __init__(self, n_heads: int, d_embed: int, d_cross: int, in_proj_bias: bool = True,
                 out_proj_bias: bool = True):
        super().__init__()
        self.q_proj = nn.Linear(d_embed, d_embed, bias=in_proj_bias)
        self.k_proj = nn.Linear(d_cross, d_embed, bias=in_proj_bias)
        self.v_proj = nn.Linear(d_cross, d_embed, bias=in_proj_bias)
        self.out_proj = nn.Linear(d_embed, d_embed, bias=out_proj_bias)
        self.n_heads = n_heads
        self.d_head = d_embed // n_heads

    @staticmethod
    def _pack(self, dev):
        from types import SimpleNamespace as NS
        pk = NS()
        pk.wq, pk.bq = engine.pack_linear(self.q_proj, dev)
        pk.wk, pk.bk = engine.pack_linear(self.k_proj, dev)
        pk.wv, pk.bv = engine.pack_linear(self.v_proj, dev)
        pk.wo, pk.bo = engine.pack_linear(self.out_proj, dev)
        return pk

    def forward(self, x, y):
        """x: (B, S, E), y: (B, T, d_cross) fp32 -> (B, S, E) fp32 (sd/attention.py:161-253)."""
        _require_cuda(x, "CrossAttention")
        pk = self._packed(CrossAttention._pack)
        b, s, e = x.shape
        t = y.shape[1]
        t_pad = (t + 7) // 8 * 8
        yb = torch.zeros((b, t_pad, y.shape[2]), device=x.device, dtype=torch.bfloat16)
        yb[:, :t] = y.to(torch.bfloat16)
        yb = yb.view(b * t_pad, -1)
        xb = x.to(torch.bfloat16).contiguous().view(b * s, e)
        q = ops.linear(xb, pk.wq, bias=pk.bq)
        k = ops.linear(yb, pk.wk, bias=pk.bk)
        vt = ops.gemm(pk.wv, yb, b * t_pad, M=e, c0=yb.shape[1], bias=pk.bv, bias_per_row=True)
        o = torch.empty_like(xb)
        ops.attention(q, k, vt, o, NB=b, heads=self.n_heads, d=self.d_head, S=s, Skv=t, Skv_pad=t_pad,
                      ldq=e, ldk=e, ldo=e)
        out = ops.linear(o, pk.wo, bias=pk.bo, out_fp32=True)
        return out.view(b, s, e)
