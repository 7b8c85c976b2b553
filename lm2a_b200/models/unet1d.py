"""Legacy UNet1D — drop-in for the reference's first-generation noise predictor
(reference models/unet1d.py:15-60 ResBlock1D, :64-154 UNet1D).

Differences from UNet1D_ultimate that the launch plan honours (engine.LegacyPacked /
UNetPlan._build_legacy):
  * one constant-width ResBlock per level, cross-attention in every block (7 blocks,
    4 heads -> head dims 64..384), attention output replaces h, residual is `x + h`;
  * the timestep enters additively: h + time_proj(t_emb) (no FiLM scale, no SiLU in
    front of the projection);
  * down: Conv1d k4 s2 p1 that also changes the width; up: ConvTranspose1d k4 s2 p1,
    run as two tcgen05 GEMMs (even / odd output slots) straight into the concat slab;
  * the decoder width grows with every concat (dim + skip: 1536 / 768 / 512 at
    base_dim = 256); out_proj is a bare 1x1 conv.

As in unet1d_ultimate.py the module tree only holds parameters under the reference's
names (`input_proj`, `downs.{i}.0.*` block / `downs.{i}.1` conv, `mid.*`, `ups.{i}.0`
transposed conv / `ups.{i}.1.*` block, `out_proj`); forward hands device pointers to
the sm_100a kernels.
"""
import torch
import torch.nn as nn

from .cross_attention import CrossAttentionFusion
from .embedding import TimestepEmbedding


class ResBlock1D(nn.Module):
    """norm1-SiLU-conv1 -> + time_proj(t) -> norm2-SiLU-conv2 -> cross-attn -> x + h
    (reference models/unet1d.py:15-60)."""

    def __init__(self, channels, time_emb_dim, cond_dim=128, mel_dim=80, num_heads=4):
        super().__init__()
        self.conv1 = nn.Conv1d(channels, channels, kernel_size=3, padding=1)
        self.conv2 = nn.Conv1d(channels, channels, kernel_size=3, padding=1)
        self.time_proj = nn.Linear(time_emb_dim, channels)
        self.cross_attn = CrossAttentionFusion(mel_dim=channels, cond_dim=cond_dim,
                                               num_heads=num_heads)
        self.norm1 = nn.GroupNorm(8, channels)
        self.norm2 = nn.GroupNorm(8, channels)
        self.act = nn.SiLU()
        # what the weight packer reads (same vocabulary as the ultimate block)
        self.in_channels = self.out_channels = channels
        self.num_heads = num_heads

    def forward(self, x, t_emb, motion_f, text_f):
        raise RuntimeError("legacy ResBlock1D runs only inside UNet1D.forward on the fused "
                           "sm_100a path (lm2a_b200 has no per-module PyTorch fallback)")


class UNet1D(nn.Module):
    def __init__(self, in_dim=80, base_dim=128, dim_mults=(1, 2, 4), cond_dim=128,
                 time_emb_dim=256):
        super().__init__()
        self.in_dim, self.base_dim, self.dim_mults = in_dim, base_dim, tuple(dim_mults)
        self.cond_dim, self.time_emb_dim = cond_dim, time_emb_dim
        self.time_embedding = TimestepEmbedding(time_emb_dim)
        self.input_proj = nn.Conv1d(in_dim, base_dim, kernel_size=1)

        widths = [base_dim * m for m in dim_mults]
        self.downs = nn.ModuleList()
        cur, skips = base_dim, []
        for w in widths:
            self.downs.append(nn.ModuleList([
                ResBlock1D(cur, time_emb_dim, cond_dim=cond_dim, mel_dim=in_dim),
                nn.Conv1d(cur, w, kernel_size=4, stride=2, padding=1)]))
            skips.append(cur)
            cur = w
        self.mid = ResBlock1D(cur, time_emb_dim, cond_dim=cond_dim, mel_dim=in_dim)
        self.ups = nn.ModuleList()
        for w, skip in zip(reversed(widths), reversed(skips)):
            self.ups.append(nn.ModuleList([
                nn.ConvTranspose1d(cur, w, kernel_size=4, stride=2, padding=1),
                ResBlock1D(w + skip, time_emb_dim, cond_dim=cond_dim, mel_dim=in_dim)]))
            cur = w + skip
        self.out_proj = nn.Conv1d(cur, in_dim, kernel_size=1)
        self._engine = None

    def engine(self):
        """See UNet1D_ultimate.engine: re-packed when a parameter's (data_ptr, version) changes."""
        from ..engine import UNetEngine, params_fingerprint
        if self._engine is not None and self._engine.fingerprint != params_fingerprint(self):
            self._engine = None
        if self._engine is None:
            self._engine = UNetEngine(self)
        return self._engine

    def refresh(self):
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        self._engine = None
        return super().load_state_dict(state_dict, strict=strict, **kw)

    @torch.no_grad()
    def forward(self, x, t, motion_f, text_f):
        """x (B, in_dim, T), t (B,) long, motion_f / text_f (B, Lk, cond_dim) ->
        predicted noise (B, in_dim, T); reference models/unet1d.py:113-154."""
        if motion_f is None or text_f is None:
            raise RuntimeError("legacy UNet1D.forward needs motion_f and text_f (every block "
                               "cross-attends; reference models/unet1d.py:52-55)")
        return self.engine().forward(x, t, motion_f, text_f)
