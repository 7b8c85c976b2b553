"""Timestep embedding and condition projection (reference models/embedding.py:19-55).

TimestepEmbedding / SinusoidalPosEmb are parameter containers: inside the UNet the
embedding, its MLP and all FiLM tables are produced by lm2a_time_mlp + lm2a_film.
CondProjection runs once per clip as two bf16 tcgen05 GEMMs (lm2a_conv1d_bf16, k=1)."""
import torch
import torch.nn as nn

from .. import ops


class SinusoidalPosEmb(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, t):
        raise RuntimeError("SinusoidalPosEmb is fused into lm2a_time_mlp; no standalone path")


class TimestepEmbedding(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(dim), nn.Linear(dim, dim), nn.SiLU())

    def forward(self, t):
        raise RuntimeError("TimestepEmbedding is fused into lm2a_time_mlp; no standalone path")


class CondProjection(nn.Module):
    """motion (B, T, 234) -> (B, T, out_dim); lyrics (B, T, 768) -> (B, T, out_dim)."""

    def __init__(self, motion_dim=78 * 3, text_dim=768, out_dim=128):
        super().__init__()
        self.motion_proj = nn.Linear(motion_dim, out_dim)
        self.text_proj = nn.Linear(text_dim, out_dim)
        self.out_dim = out_dim

    @torch.no_grad()
    def forward(self, motion, lyrics):
        return _project(self.motion_proj, motion), _project(self.text_proj, lyrics)


def _pad_to(n, m):
    return (n + m - 1) // m * m


def _project(lin, x):
    """fp32 [B, T, D] -> fp32 [B, T, out]: one k=1 implicit-GEMM launch on a bf16 slab."""
    ops.require_device(x)
    if x.dim() != 3:
        raise RuntimeError(f"CondProjection expects (B, T, D); got {tuple(x.shape)}")
    b, t, d = x.shape
    out_dim = lin.out_features
    if out_dim % 32 != 0:
        raise RuntimeError("CondProjection out_dim must be a multiple of 32 on the B200 path")
    kp = _pad_to(d, 64)
    np_ = _pad_to(out_dim, 128)
    dev = x.device
    w = torch.zeros(np_, kp, dtype=torch.bfloat16, device=dev)
    w[:out_dim, :d] = lin.weight.detach().to(torch.bfloat16)
    bias = torch.zeros(np_, dtype=torch.float32, device=dev)
    bias[:out_dim] = lin.bias.detach().float()
    slab = torch.empty(b * t, kp, dtype=torch.bfloat16, device=dev)
    ops.ingest_seq(x.contiguous().float(), slab, b, t, d, t, kp)
    out = torch.empty(b * t, out_dim, dtype=torch.bfloat16, device=dev)
    desc = ops.make_conv_desc([ops.Seg(slab, kp, kp, ops.TAPS_K1, b * t)], w, bias, out_dim,
                              b * t, t, t, out, out_dim)
    ops.conv1d(desc)
    return out.float().view(b, t, out_dim)
