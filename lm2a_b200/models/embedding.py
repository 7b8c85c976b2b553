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

    @torch.no_grad()
    def project_raw(self, motion, motion_lens, lyrics, lyrics_lens, t_len, dst_motion=None,
                    dst_text=None, want_resampled=True):
        """Per-clip prologue on the GPU: match_len(..., 'interp') of the raw
        condition sequences (datasetcode/dataset.py:49-87, sample.py:124-125) fused with the
        bf16 staging of the projection GEMM, then CondProjection (embedding.py:45-55).

        motion (B, Lm_max, 234) / lyrics (B, Ll_max, 768): padded fp32 device tensors with
        per-clip valid lengths `*_lens` (int32 device tensors or None = full length).
        dst_*: optional bf16 [B * t_len, out_dim] slabs to project into (the sampler's
        condition slabs, so the K/V cache build reads them without another copy).
        Returns (motion_f, text_f, motion_rs, lyrics_rs): projected bf16 slabs and — if
        `want_resampled` — the resampled fp32 sequences (B, t_len, D) the output npz stores."""
        outs, resampled = [], []
        for lin, x, lens, dst in ((self.motion_proj, motion, motion_lens, dst_motion),
                                  (self.text_proj, lyrics, lyrics_lens, dst_text)):
            ops.require_device(x)
            if x.dim() != 3 or x.dtype != torch.float32:
                raise RuntimeError(f"project_raw expects fp32 (B, L, D); got {tuple(x.shape)} "
                                   f"{x.dtype}")
            b, lmax, d = x.shape
            w, bias, kp, out_dim = _packed(lin)
            rs = torch.empty(b, t_len, d, dtype=torch.float32, device=x.device) \
                if want_resampled else None
            slab = torch.empty(b * t_len, kp, dtype=torch.bfloat16, device=x.device)
            ops.resample_seq(x.contiguous(), lens, rs, slab, b, lmax, d, t_len, t_len, kp)
            if dst is None:
                dst = torch.empty(b * t_len, out_dim, dtype=torch.bfloat16, device=x.device)
            elif tuple(dst.shape) != (b * t_len, out_dim) or dst.dtype != torch.bfloat16:
                raise RuntimeError("project_raw: dst must be a bf16 (B * t_len, out_dim) slab")
            ops.conv1d(ops.make_conv_desc([ops.Seg(slab, kp, kp, ops.TAPS_K1, b * t_len)], w,
                                          bias, out_dim, b * t_len, t_len, t_len, dst,
                                          dst.stride(0)))
            outs.append(dst)
            resampled.append(rs)
        return outs[0], outs[1], resampled[0], resampled[1]


def _pad_to(n, m):
    return (n + m - 1) // m * m


def _packed(lin):
    """bf16 [n_pad, k_pad] GEMM operand + fp32 bias of a Linear, cached on the module until
    its parameters change (in-place update, load_state_dict or a device move)."""
    out_dim, d = lin.out_features, lin.in_features
    if out_dim % 32 != 0:
        raise RuntimeError("CondProjection out_dim must be a multiple of 32 on the B200 path")
    key = (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version)
    cached = getattr(lin, "_lm2a_packed", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    kp, np_ = _pad_to(d, 64), _pad_to(out_dim, 128)
    dev = lin.weight.device
    w = torch.zeros(np_, kp, dtype=torch.bfloat16, device=dev)
    w[:out_dim, :d] = lin.weight.detach().to(torch.bfloat16)
    bias = torch.zeros(np_, dtype=torch.float32, device=dev)
    bias[:out_dim] = lin.bias.detach().float()
    lin._lm2a_packed = (key, (w, bias, kp, out_dim))
    return lin._lm2a_packed[1]


def _project(lin, x):
    """fp32 [B, T, D] -> fp32 [B, T, out]: one k=1 implicit-GEMM launch on a bf16 slab."""
    ops.require_device(x)
    if x.dim() != 3:
        raise RuntimeError(f"CondProjection expects (B, T, D); got {tuple(x.shape)}")
    b, t, d = x.shape
    w, bias, kp, out_dim = _packed(lin)
    dev = x.device
    slab = torch.empty(b * t, kp, dtype=torch.bfloat16, device=dev)
    ops.ingest_seq(x.contiguous().float(), slab, b, t, d, t, kp)
    out = torch.empty(b * t, out_dim, dtype=torch.bfloat16, device=dev)
    desc = ops.make_conv_desc([ops.Seg(slab, kp, kp, ops.TAPS_K1, b * t)], w, bias, out_dim,
                              b * t, t, t, out, out_dim)
    ops.conv1d(desc)
    return out.float().view(b, t, out_dim)
