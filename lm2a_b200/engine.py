"""Launch planner for UNet1D_ultimate (and the legacy UNet1D) on B200.

Turns the parameter tree of models/unet1d_ultimate.py into
  (1) packed device weights: bf16 K-major GEMM operands with the algebraic folds
      (q-scale*log2e into W_q, kv_proj into the MHA K/V in-projection, out_proj+concat+
      fuse_proj into one matrix, ResBlock skip conv appended as a second K segment),
  (2) a static plan: pre-allocated bf16 slabs + a flat list of C-ABI launches that
      computes reference UNet1D_ultimate.forward (unet1d_ultimate.py:367-426).

Slab geometry: level l has T_l valid slots per clip-row and pitch Tp_l, with
Tp_last = T_last + 1 and Tp_l = 2 * Tp_{l+1}; slots t >= T_l are zero. The zero slot is
the conv padding, and Tp_l = 2 Tp_{l+1} makes the stride-2 conv a plain 4-tap GEMM over
slot *pairs* of the flattened slab.
"""
import math
import os

import torch

from . import ops
from .ops import OUT_BF16_SLAB, OUT_F32_NCT, TAPS_K1, TAPS_K3, TAPS_K4S2, Seg

LOG2E = 1.4426950408889634
BF16 = torch.bfloat16


def _pad_to(n, m):
    return (n + m - 1) // m * m


class Geometry:
    def __init__(self, rows, t, n_down):
        self.rows = rows
        self.T = [int(t)]
        for _ in range(n_down):
            self.T.append(self.T[-1] // 2)
        if self.T[-1] < 1 or any(tl < 4 for tl in self.T[:-1]):
            raise RuntimeError(f"sequence length {t} too short for {n_down} stride-2 stages")
        self.Tp = [0] * (n_down + 1)
        self.Tp[-1] = self.T[-1] + 1
        for lvl in reversed(range(n_down)):
            self.Tp[lvl] = 2 * self.Tp[lvl + 1]
        for lvl in range(n_down + 1):
            assert self.Tp[lvl] >= self.T[lvl] + 1
        self.M = [rows * tp for tp in self.Tp]


# ------------------------------------------------------------------------ weight packing
def _f64(p):
    """Parameters are folded on the HOST in fp64 (exact algebra, a few seconds once per
    checkpoint at production width) and only the packed bf16 / fp32 operands go to the GPU: no
    library kernel runs on the device on behalf of the weight pack."""
    return p.detach().to("cpu", torch.float64)


def _conv_w(w, cin_pad=None):
    """Conv1d weight [Cout, Cin, k] -> [Cout, k*Cin_pad] (tap-major K), fp64."""
    w = _f64(w)
    cout, cin, k = w.shape
    cin_pad = cin_pad or cin
    out = torch.zeros(cout, k, cin_pad, dtype=torch.float64, device=w.device)
    out[:, :, :cin] = w.permute(0, 2, 1)
    return out.reshape(cout, k * cin_pad)


_PACK_DTYPE = [BF16]   # weight dtype of the pack in progress (fp32 for the validation path)


def _finish(w, b, dev):
    """Pad rows to a multiple of 128, cast: W -> bf16 (fp32 on the validation path), bias -> fp32."""
    n = w.shape[0]
    n_pad = _pad_to(n, 128)
    wp = torch.zeros(n_pad, w.shape[1], dtype=torch.float64, device=w.device)
    wp[:n] = w
    bp = torch.zeros(n_pad, dtype=torch.float64, device=w.device)
    bp[:n] = b
    return wp.to(dev, _PACK_DTYPE[0]).contiguous(), bp.to(dev, torch.float32).contiguous()


def _gn(gn, dev):
    if gn.num_channels % gn.num_groups != 0 or (gn.num_channels // gn.num_groups) % 8 != 0:
        raise RuntimeError(f"GroupNorm over {gn.num_channels} channels in {gn.num_groups} groups: "
                           "channels per group must be a multiple of 8 on the sm_100a path")
    return (gn.weight.detach().to(dev, torch.float32).contiguous(),
            gn.bias.detach().to(dev, torch.float32).contiguous(), gn.num_groups, float(gn.eps))


class PackedBlock:
    pass


def _check_channels(c, what):
    if c % 64 != 0:
        raise RuntimeError(f"{what}={c}: the sm_100a path needs channel counts that are "
                           "multiples of 64 (64-wide K blocks, 8-channel GroupNorm vectors)")


HEAD_DIMS = (32, 64, 96, 128, 192, 256, 384)


def pack_block(blk, heads, dev, film_col):
    """Packs one ResBlock: UNet1D_ultimate's (gn1/gn2, FiLM, optional skip conv and attention)
    or the legacy UNet1D's (norm1/norm2, constant width, always attention, identity residual —
    reference models/unet1d.py:15-60)."""
    cin, cout = blk.in_channels, blk.out_channels
    _check_channels(cin, "in_channels")
    _check_channels(cout, "out_channels")
    p = PackedBlock()
    p.cin, p.cout, p.attn = cin, cout, bool(getattr(blk, "use_attn", True))
    gn1 = blk.gn1 if hasattr(blk, "gn1") else blk.norm1
    gn2 = blk.gn2 if hasattr(blk, "gn2") else blk.norm2
    p.gn1, p.gn2 = _gn(gn1, dev), _gn(gn2, dev)
    p.film_col = film_col
    p.w1, p.b1 = _finish(_conv_w(blk.conv1.weight), _f64(blk.conv1.bias), dev)
    has_skip = hasattr(blk, "skip") and not isinstance(blk.skip, torch.nn.Identity)
    p.has_skip = has_skip
    w2, b2 = _conv_w(blk.conv2.weight), _f64(blk.conv2.bias)
    wskip = _f64(blk.skip.weight)[:, :, 0] if has_skip else None
    bskip = _f64(blk.skip.bias) if has_skip else None
    # plain variant (also used when forward is called without conditions)
    if has_skip:
        p.w2s, p.b2s = _finish(torch.cat([w2, wskip], dim=1), b2 + bskip, dev)
    else:
        p.w2s, p.b2s = _finish(w2, b2, dev)
    if p.attn:
        ca = blk.cross_attn
        e = cout
        if e % heads != 0 or (e // heads) not in HEAD_DIMS:
            raise RuntimeError(f"attention head dim {e}/{heads} unsupported on the sm_100a path "
                               f"(needs one of {HEAD_DIMS})")
        p.e, p.heads = e, heads
        qs = LOG2E / math.sqrt(e // heads)
        wq, bq, wkv, bkv, wo, bo = [], [], [], [], [], []
        for mha, kvp in ((ca.attn_motion, ca.motion_kv_proj), (ca.attn_text, ca.text_kv_proj)):
            ipw, ipb = _f64(mha.in_proj_weight), _f64(mha.in_proj_bias)
            wq.append(ipw[:e] * qs)
            bq.append(ipb[:e] * qs)
            wp_, bp_ = _f64(kvp.weight), _f64(kvp.bias)  # cond_dim -> e
            wk, wv = ipw[e:2 * e], ipw[2 * e:]
            # K = (c Wp^T + bp) Wk^T + bk  ==  c (Wk Wp)^T + (Wk bp + bk); same for V
            wkv.append(torch.cat([wk @ wp_, wv @ wp_], dim=0))
            bkv.append(torch.cat([wk @ bp_ + ipb[e:2 * e], wv @ bp_ + ipb[2 * e:]], dim=0))
            wo.append(_f64(mha.out_proj.weight))
            bo.append(_f64(mha.out_proj.bias))
        # conv2 -> Q in-projection composed (h2 is only ever the attention query: the attention
        # output replaces it, unet1d_ultimate.py:152-159): q = Wq (W2 * n + b2) + bq
        wq_all, bq_all = torch.cat(wq, dim=0), torch.cat(bq, dim=0)
        p.wq2, p.bq2 = _finish(wq_all @ w2, wq_all @ b2 + bq_all, dev)
        # motion stream alone (the lyrics stream is constant in time: see UNetPlan.const_text)
        p.wq2_m, p.bq2_m = _finish(wq[0] @ w2, wq[0] @ b2 + bq[0], dev)
        p.wkv_m, p.bkv_m = _finish(wkv[0], bkv[0], dev)
        p.wkv_t, p.bkv_t = _finish(wkv[1], bkv[1], dev)
        wf, bf = _f64(ca.fuse_proj.weight), _f64(ca.fuse_proj.bias)
        # fuse(cat(Om Wo_m^T + bo_m, Ot Wo_t^T + bo_t)) = [Om Ot] [Wf1 Wo_m, Wf2 Wo_t]^T + b
        wof = torch.cat([wf[:, :e] @ wo[0], wf[:, e:] @ wo[1]], dim=1)
        bof = wf[:, :e] @ bo[0] + wf[:, e:] @ bo[1] + bf
        # all-zero conditions: V rows of stream s are the constant bv'_s, softmax is uniform, so
        # the block's attention output is the constant c (used by the uncond shortcut)
        c_unc = wf[:, :e] @ (wo[0] @ bkv[0][e:] + bo[0]) + wf[:, e:] @ (wo[1] @ bkv[1][e:] + bo[1]) + bf
        if has_skip:
            wof = torch.cat([wof, wskip], dim=1)
            bof = bof + bskip
            p.wsk, p.bsk_c = _finish(wskip, bskip + c_unc, dev)
        p.c_uncond = c_unc.to(dev, torch.float32).contiguous()
        p.wof, p.bof = _finish(wof, bof, dev)
        # constant lyrics stream: out = [O_m (| x)] [Wf1 Wo_m (| Wskip)]^T + b + (Wf2 Wo_t) v_t,
        # the last term a per-clip vector computed once per batch from the clip's V row
        wof_m = wf[:, :e] @ wo[0]
        if has_skip:
            wof_m = torch.cat([wof_m, wskip], dim=1)
        p.wof_m, p.bof_m = _finish(wof_m, bof, dev)
        p.wof_t, p.zero_b = _finish(wf[:, e:] @ wo[1], torch.zeros_like(bof), dev)
        # Head dim == condition width: K_h = C Wk_h^T + bk_h and V_h = C Wv_h^T + bv_h are images
        # of the raw condition sequence C, so head h can attend to C itself (lm2a_cross_attn_cond):
        #   scores  Q_h K_h^T = (Q_h Wk_h) C^T + (Q_h bk_h) 1^T   (second term: constant per row,
        #                                                          cancels in the softmax)
        #   output  P V_h = (P C) Wv_h^T + bv_h                    (rows of P sum to 1)
        # Wk_h is folded into the (composed) query projection, Wv_h and bv_h into the output
        # projection: no per-head K / V is read in the loop.
        dh = e // heads
        p.cond_fold = dh == 128 and dh == wkv[0].shape[1]
        if p.cond_fold:
            hs = [slice(h * dh, (h + 1) * dh) for h in range(heads)]
            wqc = [torch.cat([wkv[s][:e][b].t() @ wq[s][b] for b in hs], dim=0) for s in (0, 1)]
            bqc = [torch.cat([wkv[s][:e][b].t() @ bq[s][b] for b in hs], dim=0) for s in (0, 1)]
            wqc_all, bqc_all = torch.cat(wqc, dim=0), torch.cat(bqc, dim=0)
            p.wq2c, p.bq2c = _finish(wqc_all @ w2, wqc_all @ b2 + bqc_all, dev)
            p.wq2c_m, p.bq2c_m = _finish(wqc[0] @ w2, wqc[0] @ b2 + bqc[0], dev)
            wos = [wf[:, :e] @ wo[0], wf[:, e:] @ wo[1]]      # [cout, e] per stream
            wofc = [torch.cat([wos[s][:, b] @ wkv[s][e:][b] for b in hs], dim=1) for s in (0, 1)]
            bvc = [sum(wos[s][:, b] @ bkv[s][e:][b] for b in hs) for s in (0, 1)]
            tail = [wskip] if has_skip else []
            p.wofc, p.bofc = _finish(torch.cat(wofc + tail, dim=1), bof + bvc[0] + bvc[1], dev)
            p.wofc_m, p.bofc_m = _finish(torch.cat([wofc[0]] + tail, dim=1), bof + bvc[0], dev)
    return p


class PackedModel:
    kind = "ultimate"

    def scratch_width(self, lvl):
        """Widest bf16 scratch slab (in channels) any launch of level `lvl` needs."""
        n = len(self.dims)
        width = 2 * self.dims[min(lvl, n - 1)]
        if lvl < n:
            width = max(width, self.dims[min(lvl + 1, n - 1)])  # upsampled input
        return max(width, self.base)

    def cat_width(self, lvl):
        return 2 * self.dims[lvl]

    def __init__(self, model, dev):
        self.in_dim = model.in_dim
        self.in_pad = _pad_to(model.in_dim, 64)
        self.base = model.base_dim
        self.dims = [model.base_dim * m for m in model.dim_mults]
        self.cond_dim = model.cond_dim
        _check_channels(model.cond_dim, "cond_dim")
        self.time_dim = model.time_emb_dim
        heads = model.attn_heads
        lin = model.time_embedding.time_mlp[1]
        self.time_w = lin.weight.detach().to(dev, torch.float32).contiguous()
        self.time_b = lin.bias.detach().to(dev, torch.float32).contiguous()
        self.w_in, self.b_in = _finish(_conv_w(model.in_proj.weight, self.in_pad),
                                       _f64(model.in_proj.bias), dev)
        film_w, film_b = [], []
        col = [0]

        def pack(blk):
            pb = pack_block(blk, heads, dev, col[0])
            film_w.append(blk.film.net[1].weight.detach().float())
            film_b.append(blk.film.net[1].bias.detach().float())
            col[0] += 2 * blk.out_channels
            return pb

        self.downs = []
        for stage in model.downs:
            blocks = [pack(b) for b in stage["blocks"]]
            wd, bd = _finish(_conv_w(stage["down"].conv.weight), _f64(stage["down"].conv.bias), dev)
            self.downs.append((blocks, wd, bd))
        self.mid = [pack(b) for b in model.mid.blocks]
        self.ups = []
        for stage in model.ups:
            wu, bu = _finish(_conv_w(stage["up"].conv.weight), _f64(stage["up"].conv.bias), dev)
            blocks = [pack(b) for b in stage["blocks"]]
            self.ups.append((wu, bu, blocks))
        self.film_cols = col[0]
        self.film_w = torch.cat(film_w, dim=0).to(dev).contiguous()
        self.film_b = torch.cat(film_b, dim=0).to(dev).contiguous()
        self.gn_out = _gn(model.out_proj[0], dev)
        self.w_out, self.b_out = _finish(_conv_w(model.out_proj[2].weight),
                                         _f64(model.out_proj[2].bias), dev)
        self.attn_blocks = [b for blocks, _, _ in self.downs for b in blocks if b.attn]
        self.attn_blocks += [b for b in self.mid if b.attn]
        self.attn_blocks += [b for _, _, blocks in self.ups for b in blocks if b.attn]
        gns = [m for m in model.modules() if isinstance(m, torch.nn.GroupNorm)]
        self.n_groupnorms, self.max_groups = len(gns), max(m.num_groups for m in gns)


def _convT_w(w):
    """ConvTranspose1d k4 s2 p1 weight [Cin, Cout, 4] -> two k3 GEMM operands [Cout, 3*Cin]
    (tap-major K over input slots m-1, m, m+1) for the even / odd output slots:
    y[2m] = W1 x[m] + W3 x[m-1], y[2m+1] = W2 x[m] + W0 x[m+1] (t' = 2t - 1 + k)."""
    w = _f64(w)
    cin, cout, k = w.shape
    assert k == 4
    z = torch.zeros(cout, cin, dtype=torch.float64, device=w.device)
    tap = lambda i: w[:, :, i].t()  # noqa: E731
    even = torch.cat([tap(3), tap(1), z], dim=1)
    odd = torch.cat([z, tap(2), tap(0)], dim=1)
    return even, odd


class PackedLegacy:
    """Packed weights of the legacy UNet1D (reference models/unet1d.py:64-154): level l has one
    block of width c_l (c_0 = base, c_l = dims[l-1]); the decoder block of level l is
    dims[l] + c_l wide (transposed-conv output | skip)."""
    kind = "legacy"

    def scratch_width(self, lvl):
        n = len(self.dims)
        return 2 * (self.dims[lvl] + self.c[lvl]) if lvl < n else 2 * self.dims[n - 1]

    def cat_width(self, lvl):
        return self.dims[lvl] + self.c[lvl]

    def __init__(self, model, dev):
        self.in_dim = model.in_dim
        self.in_pad = _pad_to(model.in_dim, 64)
        self.base = model.base_dim
        self.dims = [model.base_dim * m for m in model.dim_mults]
        self.c = [self.base] + self.dims[:-1]
        self.cond_dim = model.cond_dim
        _check_channels(model.cond_dim, "cond_dim")
        self.time_dim = td = model.time_emb_dim
        lin = model.time_embedding.time_mlp[1]
        self.time_w = lin.weight.detach().to(dev, torch.float32).contiguous()
        self.time_b = lin.bias.detach().to(dev, torch.float32).contiguous()
        self.w_in, self.b_in = _finish(_conv_w(model.input_proj.weight, self.in_pad),
                                       _f64(model.input_proj.bias), dev)
        film_w, film_b = [], []
        col = [0]

        def pack(blk):
            pb = pack_block(blk, blk.num_heads, dev, col[0])
            ch = blk.out_channels
            # additive timestep term as a FiLM table with scale == 0: h * (1 + 0) + time_proj(t)
            tw, tb = blk.time_proj.weight.detach().float(), blk.time_proj.bias.detach().float()
            film_w.append(torch.cat([torch.zeros_like(tw), tw], dim=0))
            film_b.append(torch.cat([torch.zeros_like(tb), tb], dim=0))
            col[0] += 2 * ch
            return pb

        self.downs = []
        for blk, conv in model.downs:
            wd, bd = _finish(_conv_w(conv.weight), _f64(conv.bias), dev)
            self.downs.append((pack(blk), wd, bd))
        self.mid = pack(model.mid)
        self.ups = []
        for convT, blk in model.ups:
            even, odd = _convT_w(convT.weight)
            we, bu = _finish(even, _f64(convT.bias), dev)
            wo, _ = _finish(odd, _f64(convT.bias), dev)
            self.ups.append((we, wo, bu, pack(blk)))
        self.film_cols = col[0]
        self.film_w = torch.cat(film_w, dim=0).to(dev).contiguous()
        self.film_b = torch.cat(film_b, dim=0).to(dev).contiguous()
        assert self.film_w.shape == (self.film_cols, td)
        self.w_out, self.b_out = _finish(_conv_w(model.out_proj.weight), _f64(model.out_proj.bias),
                                         dev)
        self.attn_blocks = ([b for b, _, _ in self.downs] + [self.mid]
                            + [b for _, _, _, b in self.ups])
        gns = [m for m in model.modules() if isinstance(m, torch.nn.GroupNorm)]
        self.n_groupnorms, self.max_groups = len(gns), max(m.num_groups for m in gns)


# ------------------------------------------------------------------------------ the plan
class UNetPlan:
    """Static launch list for `rows` clip-rows of length T attending to `lk` condition
    frames held in `nslots` K/V cache slots. `copies` rows share each input clip
    (CFG: copies=2, rows = 2B, row k*B+b reads clip b)."""

    def __init__(self, pm, rows, t, lk, nslots, copies, use_cond, dev, uniform_t=False,
                 uncond_rows=0):
        self.pm, self.rows, self.t, self.lk, self.nslots = pm, rows, t, lk, nslots
        # uniform_t: every clip-row is at the same timestep (sampling) -> one FiLM table row
        self.uniform_t = uniform_t
        self.t_rows = 1 if uniform_t else rows
        # leading rows with all-zero conditions whose attention blocks take the constant shortcut
        self.uncond_rows = uncond_rows if use_cond else 0
        assert 0 <= self.uncond_rows < rows
        self.copies, self.use_cond, self.dev = copies, use_cond, dev
        # fp32 validation path: fp32 slabs, CUDA-core kernels, every transform inside the conv
        self.fp32 = getattr(pm, "wdt", BF16) == torch.float32
        self.adt = torch.float32 if self.fp32 else BF16
        assert rows % copies == 0
        self.batch = rows // copies
        n_down = len(pm.dims)
        g = self.geo = Geometry(rows, t, n_down)
        z = lambda m, c, dt=None: torch.zeros(m, c, dtype=dt or self.adt, device=dev)  # noqa: E731

        # static I/O
        self.x_in = torch.zeros(self.batch, pm.in_dim, t, dtype=torch.float32, device=dev)
        self.t_in = torch.zeros(rows, dtype=torch.int64, device=dev)
        self.eps = torch.zeros(rows, pm.in_dim, t, dtype=torch.float32, device=dev)
        self.kv_slot = torch.zeros(rows, dtype=torch.int32, device=dev)
        self.silu_temb = torch.zeros(self.t_rows, pm.time_dim, dtype=torch.float32, device=dev)
        self.film = torch.zeros(self.t_rows, pm.film_cols, dtype=torch.float32, device=dev)

        # scratch slabs shared by all blocks (sized for the largest level)
        cmax = max(g.M[lvl] * pm.scratch_width(lvl) for lvl in range(n_down + 1))
        flat = lambda: torch.zeros(cmax, dtype=self.adt, device=dev)  # noqa: E731
        self._h1, self._q, self._o = flat(), flat(), flat()
        self._flat = flat
        self._norm = None   # normalised operands of the launches that do not transform in-kernel
        # widest operand (channels) normalised / upsampled inside the consuming conv
        self.xf_max_c = int(os.environ.get("LM2A_XF_MAX_C", "256"))
        self.up_xf_max_c = int(os.environ.get("LM2A_UP_XF_MAX_C", "512"))
        # launches of at most this many slots (M) take the in-kernel transform at any width: they
        # fill half of the SMs and are bound by operand ingest, not by the MMA (0 = off)
        self.xf_max_m = int(os.environ.get("LM2A_XF_MAX_M", "0"))
        # leftover query rows of the attention launches on the CUDA cores (parallel branch)
        # (opt-in: measured slower in the step - 26 / 46 / 25 us per launch at levels 0 / 1 / 2 for
        # work that saves the tensor-core launches 11 / 16 / 12 us, and its CTAs take registers
        # the tensor-core CTAs need; the condition-slab kernel gives its tail tile a CTA of its
        # own instead, attention_res.cu split_tail)
        self.attn_tail = os.environ.get("LM2A_ATTN_TAIL", "0") == "1"
        if self.fp32:
            self.xf_max_c = self.up_xf_max_c = 1 << 30
        # every GroupNorm statistics buffer of the plan lives in one arena that the step's first
        # kernel (ingest_x) clears; the epilogues accumulate exact integer sums into it
        self.arena = ops.StatsArena(dev, (2 * pm.n_groupnorms + 16) * rows * pm.max_groups * 2)
        self._pp = [flat(), flat()]  # block outputs ping-pong
        self.x_slab = z(g.M[0], pm.in_pad)
        self.cat = [z(g.M[lvl], pm.cat_width(lvl)) for lvl in range(n_down)]

        # K/V caches per attention block and stream: [nslots*lk, 2E] (K | V) written by the
        # projection GEMM, plus V^T [nslots*E, lk_pad] (keys contiguous: the K-major B operand
        # of the tcgen05 P.V product)
        self.kv, self.vt = [], []
        self.lk_pad = _pad_to(lk, 8) if use_cond else 0
        if use_cond:
            self.cond_m = z(nslots * lk, pm.cond_dim)
            self.cond_t = z(nslots * lk, pm.cond_dim)
            for b in pm.attn_blocks:
                self.kv.append((z(nslots * lk, 2 * b.e), z(nslots * lk, 2 * b.e)))
                self.vt.append((None, None) if self.fp32 else
                               (z(nslots * b.e, self.lk_pad), z(nslots * b.e, self.lk_pad)))
        # two launch lists over the same buffers: the full one, and the variant used when the
        # lyrics stream is constant in time for every clip of the batch (what the reference's
        # preprocessing produces: ONE sentence embedding tiled over all frames,
        # preprocess.py:64-71). Every key of that stream is then identical, its softmax uniform
        # and its attention output the stream's single V row: its image under the folded output
        # projection is computed once per batch as a per-clip shift of the block's output GEMM,
        # the per-step attention launch computes the motion stream only and the Q projection
        # produces motion queries only.
        self._ops_by_mode = {False: [], True: []}
        self.const_text = False
        self._building_ct = False
        self.allow_const_text = (os.environ.get("LM2A_CONST_STREAM", "1") != "0"
                                 and not self.fp32)
        # blocks whose head dim equals the condition width attend to the raw condition slabs
        self.cond_attn = (use_cond and not self.fp32
                          and os.environ.get("LM2A_COND_ATTN", "1") != "0"
                          and ops.cond_attn_supported(lk))
        # per attention block: fp32 [rows, 2 * cout] table (zero scale | per-clip shift), built by
        # ct_ops once per batch
        self.ct_tab = [None] * (len(pm.attn_blocks) if use_cond else 0)
        self.ct_ops = []
        self.kv_ops = []
        self.use_side_stream = True
        self._side = None
        self._side_op = self._side_pending = self._partial_rows = False
        self._stats_pool = {}
        for mode in ((False, True) if use_cond else (False,)):
            self._building_ct = mode
            self._stats_seq = []
            self.kv_ops = []
            self._attn_i = 0
            self._shared_rows = 0
            self._side_op = self._side_pending = self._partial_rows = False
            self._hold_join = self._force_join = False
            if pm.kind == "legacy":
                self._build_legacy()
            else:
                self._build()
        self._building_ct = False

    @property
    def ops(self):
        """The launch list in effect (full, or the constant-lyrics-stream variant)."""
        return self._ops_by_mode[self.const_text]

    # -- helpers -------------------------------------------------------------------------
    def _view(self, flat, m, c):
        return flat[: m * c].view(m, c)

    def _add(self, fn, *args, meta=None):
        meta = meta or {"kind": fn.__name__, "flops": 0}
        if self._side_op:
            meta["side"] = True
            self._side_pending = True
        elif self._force_join:
            # first launch of the first ResBlock: the FiLM tables of the side branch must be there
            meta["join"] = True
            self._force_join = self._side_pending = False
        elif self._side_pending and not self._partial_rows and not self._hold_join:
            meta["join"] = True   # first main op that reads rows written on the side stream
            self._side_pending = False
        self._ops_by_mode[self._building_ct].append((fn, args, meta))

    def _conv(self, segs, w, bias, n_valid, m, tp, t_valid, *a, **k):
        """Queues one implicit-GEMM launch; meta carries its ALGORITHMIC flops
        (valid slots x real output channels x K, no padding counted)."""
        ntaps = {TAPS_K1: 1, TAPS_K3: 3, TAPS_K4S2: 4}
        k_total = k.pop("k_real", None) or sum(ntaps[s.taps] * s.cin for s in segs)
        executed = 2 * (m // tp) * t_valid * n_valid * k_total
        flops = min(executed, k.pop("flops_alg", None) or executed)
        meta = {"kind": "conv_gemm", "flops": flops, "flops_executed": executed,
                "m": m, "n": n_valid, "k": k_total, "in_gn": k.get("in_gn") is not None,
                "up2x": k.get("up2x") is not None}
        self._add(ops.conv1d_f32 if self.fp32 else ops.conv1d,
                  ops.make_conv_desc(segs, w, bias, n_valid, m, tp, t_valid, *a, **k), meta=meta)

    def _stats(self, rows, lvl, c, groups):
        """Statistics buffer of a [rows, Tp_lvl, c] slab whose consumer normalises it in `groups`
        groups. Both launch lists (full / constant lyrics) are built over the same buffers, so
        the n-th request of either build returns the same object."""
        key = len(self._stats_seq)
        self._stats_seq.append(key)
        if key not in self._stats_pool:
            self._stats_pool[key] = ops.Stats(rows, c, groups, self.dev, arena=self.arena)
        st = self._stats_pool[key]
        assert (st.rows, st.c, st.groups) == (rows, c, groups)
        return st

    def _gn_operand(self, x, x_ld, x_off, x_st, gn, nr, tp, tv, c):
        """Input of a conv that follows GroupNorm + SiLU: (slab, ld, element offset, in_gn).
        Narrow operands (c <= xf_max_c = 256: level 0) are normalised inside the conv (operand
        transform: the launch is load- / latency-bound there and the transform is free); wider
        operands feed MMA-bound GEMMs, where the transform's shared-memory traffic costs more
        than one streaming gn_apply pass in front of a plain launch (measured per level on B200:
        1.938 ms/step with the threshold at 256, 1.957 at 512, 2.19 with every operand
        transformed in-kernel). Both paths evaluate the same arithmetic from the same exact sums: the
        choice does not change a bit of the result. Clips too short for the transform also take
        the stand-alone pass."""
        gm, bt, groups, eps = gn
        if ((c <= self.xf_max_c or nr * tp <= self.xf_max_m)
                and (self.fp32 or ops.in_gn_supported(tp, groups))):
            return x, x_ld, x_off, (x_st, gm, bt, eps, True)
        if self._norm is None:
            self._norm = [self._flat(), self._flat()]
        self._norm_i = getattr(self, "_norm_i", 0) ^ 1
        norm = self._view(self._norm[self._norm_i], nr * tp, c)
        self._add(ops.gn_apply, x, x_ld, norm, c, x_st, gm, bt, nr, tp, tv, c, groups, eps, True,
                  x_off, 0, meta={"kind": "gn_apply", "flops": 0})
        return norm, c, 0, None

    def _resblock(self, p, lvl, xin, xin_ld, xin_off, xin_st, out, out_ld, out_off, out_st, kv):
        """xin/out: (tensor, ld, channel offset, Stats) views of slabs at level `lvl`; the block
        consumes the partial GroupNorm sums of xin and produces those of out. With the uncond
        shortcut, attention blocks run the full pipeline on the cond rows only; the leading
        `uncond_rows` rows get `skip(x) + const` (their attention output is Q-independent)."""
        u = self.uncond_rows if (p.attn and self.use_cond) else 0
        if u > 0:
            # rows [0, u) only need skip(x) + const: one small launch, independent of the cond
            # rows' pipeline below -> side stream; the block's own main ops touch rows >= u only.
            # While the CFG copies are still identical (_shared_rows: nothing has attended yet)
            # the uncond rows' input is read from the cond copy of the same clip.
            self._side_op = True
            self._uncond_block(p, lvl, xin, xin_ld, xin_off, out, out_ld, out_off, out_st, u,
                               self._shared_rows)
            self._side_op = False
            self._partial_rows = True
            self._shared_rows = 0
        elif self._shared_rows:
            # no attention yet: uncond row b == cond row B + b, compute the cond copy only
            self._resblock_rows(p, lvl, xin, xin_ld, xin_off, xin_st, out, out_ld, out_off, out_st,
                                kv, self._shared_rows, self.rows - self._shared_rows)
            return
        self._resblock_rows(p, lvl, xin, xin_ld, xin_off, xin_st, out, out_ld, out_off, out_st, kv,
                            u, self.rows - u)
        self._partial_rows = False

    def _uncond_block(self, p, lvl, xin, xin_ld, xin_off, out, out_ld, out_off, out_st, u,
                      src_row0=0):
        """Rows whose conditions are all-zero (sample.py:155-157): every key of a stream is the
        same vector, softmax is exactly uniform, the attention output is the constant
        fuse(out_proj(v_row)) and replaces h (unet1d_ultimate.py:152-159) -> out = skip(x) + c."""
        g = self.geo
        tp, tv = g.Tp[lvl], g.T[lvl]
        m = u * tp
        st = out_st.view(0, out_off)
        xin_off = xin_off + src_row0 * tp * xin_ld   # element offset of the source row range
        if p.has_skip:
            self._conv([Seg(xin, xin_ld, p.cin, TAPS_K1, m, xin_off)], p.wsk, p.bsk_c, p.cout, m,
                       tp, tv, out, out_ld, out_chan_off=out_off, stats=st)
        else:
            self._add(ops.bias_add_f32 if self.fp32 else ops.bias_add, xin, xin_ld, xin_off, out,
                      out_ld, out_off, p.c_uncond, m, tp, tv, p.cout, st,
                      meta={"kind": "bias_add", "flops": 0})

    def _make_ct_table(self, ai, p, kv_t):
        """Per-batch setup of attention block `ai` for the constant-lyrics launch list: the clip's
        V row (all Lk rows of the text V cache are identical) -> c = (Wf2 Wo_t) v as one small
        implicit-GEMM launch with fp32 output -> shift half of the block's per-row table."""
        e, cout, n = p.e, p.cout, self.nslots
        tab = torch.zeros(self.rows, 2 * cout, dtype=torch.float32, device=self.dev)
        v_rows = torch.zeros(n, e, dtype=BF16, device=self.dev)
        c_t = torch.zeros(1, cout, n, dtype=torch.float32, device=self.dev)
        desc = ops.make_conv_desc([Seg(v_rows, e, e, TAPS_K1, n)], p.wof_t, p.zero_b, cout, n, n, n,
                                  c_t, 0, out_mode=OUT_F32_NCT, block_n=128)
        self.ct_tab[ai] = tab

        def fill():
            v_rows.copy_(kv_t.view(n, self.lk, 2 * e)[:, 0, e:])
            ops.conv1d(desc)
            tab[:, cout:] = c_t[0].t()[self.kv_slot.long()]
        self.ct_ops.append(fill)

    def _resblock_rows(self, p, lvl, xin, xin_ld, xin_off, xin_st, out, out_ld, out_off, out_st,
                       kv, r0, nr):
        g = self.geo
        tp, tv = g.Tp[lvl], g.T[lvl]
        m = nr * tp
        xo = r0 * tp * xin_ld + xin_off      # element offsets of the row range inside the slabs
        oo = r0 * tp * out_ld + out_off
        cin, cout = p.cin, p.cout
        assert xin_off == 0, "a ResBlock input starts at channel 0 of its slab"
        h1 = self._view(self._h1, m, cout)
        # conv1 reads x through gn1 + SiLU (operand transform), FiLM in its epilogue, and emits
        # the exact sums gn2 needs; conv2 (or the composed conv2 . Q projection) reads h1 through
        # gn2 + SiLU the same way: no normalised tensor is ever written
        a, a_ld, a_off, in_gn = self._gn_operand(xin, xin_ld, xo, xin_st.view(r0, 0), p.gn1, nr,
                                                 tp, tv, cin)
        h1_st = self._stats(nr, lvl, cout, p.gn2[2])
        self._conv([Seg(a, a_ld, cin, TAPS_K3, m, a_off)], p.w1, p.b1, cout, m, tp, tv, h1, cout,
                   film=self.film, film_col=p.film_col, film_shift_off=cout,
                   film_bcast=self.uniform_t, film_row=r0, stats=h1_st, in_gn=in_gn)
        n2, n2_ld, n2_off, in_gn2 = self._gn_operand(h1, cout, 0, h1_st, p.gn2, nr, tp, tv, cout)
        main = [Seg(n2, n2_ld, cout, TAPS_K3, m, n2_off)]
        skip_seg = [Seg(xin, xin_ld, cin, TAPS_K1, m, xo)] if p.has_skip else []
        res = {} if p.has_skip else dict(residual=xin, res_ld=xin_ld, res_chan_off=xo)
        ost = out_st.view(r0, out_off)
        if not (p.attn and self.use_cond):
            self._conv(main + skip_seg, p.w2s, p.b2s, cout, m, tp, tv, out, out_ld,
                       out_chan_off=oo, stats=ost, in_gn=in_gn2, **res)
            return
        e = p.e
        (kv_m, kv_t), (vt_m, vt_t) = kv
        ai = self._attn_i
        self._attn_i += 1
        rows_valid = nr * tv
        cond = self.cond_attn and p.cond_fold
        cdim = self.pm.cond_dim

        def attend(q, o, width, n_streams):
            flops = n_streams * 4 * nr * tv * self.lk * e
            # T mod 128 query rows (4 / 2 / 1 at T = 516 / 258 / 129) on the CUDA cores, on the
            # parallel branch next to the tensor-core launch, which then covers whole tiles only
            # (in that kernel the leftover rows cost a CTA slot - on the condition slab a serial
            # round - of their own); joined in front of the output GEMM
            tail = tv % 128
            if (self.attn_tail and not self.fp32 and self.use_side_stream and 0 < tail <= 8 < tv
                    and e // p.heads in (32, 64, 128)):
                # keys / values row-major: the condition slabs, or the K | V projection output
                if cond:
                    km = vm = ops._ptr(self.cond_m)
                    kt = vt = ops._ptr(self.cond_t)
                    ld = cdim
                else:
                    km, vm, kt, vt = (ops._ptr(kv_m), ops._ptr(kv_m, e), ops._ptr(kv_t),
                                      ops._ptr(kv_t, e))
                    ld = 2 * e
                self._side_op = True
                self._add(ops.cross_attn_tail, q, width, o, width, km, vm, kt, vt, ld, ld,
                          ops._ptr(self.kv_slot, r0), self.nslots, nr, tp, tv - tail, tail, self.lk,
                          e, p.heads, n_streams, cond,
                          meta={"kind": "cross_attn_tail", "flops": 0})
                self._side_op = False
                hold, self._hold_join = self._hold_join, True
                attend_rows(q, o, width, n_streams, tv - tail, flops)
                self._hold_join = hold
                self._force_join = True    # the next main launch reads the tail rows of o
                return
            attend_rows(q, o, width, n_streams, tv, flops)

        def attend_rows(q, o, width, n_streams, tv, flops):
            if self.fp32:
                self._add(ops.cross_attn_f32, q, width, o, width, kv_m, kv_t, 2 * e,
                          ops._ptr(self.kv_slot, r0), self.nslots, nr, tp, tv, self.lk, e, p.heads,
                          n_streams, meta={"kind": "cross_attn", "flops": flops})
            elif cond:
                self._add(ops.cross_attn_cond, q, width, o, width, ops._ptr(self.cond_m),
                          ops._ptr(self.cond_t), cdim, ops._ptr(self.kv_slot, r0), self.nslots, nr,
                          tp, tv, self.lk, p.heads, n_streams,
                          meta={"kind": "cross_attn", "flops": flops, "cond": True})
            else:
                self._add(ops.cross_attn, q, width, o, width, ops._ptr(kv_m), ops._ptr(vt_m),
                          ops._ptr(kv_t), ops._ptr(vt_t), 2 * e, self.lk_pad,
                          ops._ptr(self.kv_slot, r0), self.nslots, nr, tp, tv, self.lk, e, p.heads,
                          n_streams, meta={"kind": "cross_attn", "flops": flops})
        if self._building_ct:
            # motion stream only; the lyrics stream's contribution (Wf2 Wo_t) v_t is a per-clip
            # vector: a per-row epilogue shift of the output GEMM, read from the block's table
            if self.ct_tab[ai] is None:
                self._make_ct_table(ai, p, kv_t)
            q = self._view(self._q, m, e)
            o = self._view(self._o, m, e)
            # algorithmic work of conv2 (6 T C^2) + one Q projection (2 T C^2) exceeds what the
            # composed GEMM executes (6 T C^2): credit the executed FLOPs
            wq_, bq_ = (p.wq2c_m, p.bq2c_m) if cond else (p.wq2_m, p.bq2_m)
            wo_, bo_ = (p.wofc_m, p.bofc_m) if cond else (p.wof_m, p.bof_m)
            self._conv(main, wq_, bq_, e, m, tp, tv, q, e, in_gn=in_gn2)
            attend(q, o, e, 1)
            self._conv([Seg(o, e, e, TAPS_K1, m)] + skip_seg, wo_, bo_, cout, m, tp, tv,
                       out, out_ld, out_chan_off=oo, stats=ost, film=self.ct_tab[ai], film_col=0,
                       film_shift_off=cout, film_bcast=False, film_row=r0, **res)
            return
        q = self._view(self._q, m, 2 * e)
        o = self._view(self._o, m, 2 * e)
        # the composed conv2 . [Q_motion | Q_lyrics] GEMM executes 2 * 2E * 3C MACs per slot where
        # the reference's conv2 followed by two Q projections needs 3C * C + 2E * C: credit that
        wq_, bq_ = (p.wq2c, p.bq2c) if cond else (p.wq2, p.bq2)
        wo_, bo_ = (p.wofc, p.bofc) if cond else (p.wof, p.bof)
        self._conv(main, wq_, bq_, 2 * e, m, tp, tv, q, 2 * e, in_gn=in_gn2,
                   flops_alg=2 * rows_valid * (3 * cout * cout + 2 * e * cout))
        attend(q, o, 2 * e, 2)
        self._conv([Seg(o, 2 * e, 2 * e, TAPS_K1, m)] + skip_seg, wo_, bo_, cout, m, tp, tv,
                   out, out_ld, out_chan_off=oo, stats=ost, **res)

    # -- plan construction ----------------------------------------------------------------
    def _build_kv_ops(self):
        """Per-clip K/V cache build: one projection GEMM + one V transpose per block and stream."""
        pm = self.pm
        if not self.use_cond:
            return
        for p, (kv_m, kv_t), (vt_m, vt_t) in zip(pm.attn_blocks, self.kv, self.vt):
            n = self.nslots * self.lk
            for cond, w, b, dst, vt in ((self.cond_m, p.wkv_m, p.bkv_m, kv_m, vt_m),
                                        (self.cond_t, p.wkv_t, p.bkv_t, kv_t, vt_t)):
                self.kv_ops.append((ops.conv1d_f32 if self.fp32 else ops.conv1d, (ops.make_conv_desc(
                    [Seg(cond, pm.cond_dim, pm.cond_dim, TAPS_K1, n)], w, b, 2 * p.e, n,
                    self.lk, self.lk, dst, 2 * p.e),)))
                if not self.fp32:   # (the fp32 attention reads V from the K | V slab itself)
                    self.kv_ops.append((ops.transpose_kv, (dst, 2 * p.e, p.e, vt, self.lk_pad,
                                                           self.nslots, self.lk, p.e)))

    def _build_legacy(self):
        """Launch list of the legacy UNet1D.forward (reference models/unet1d.py:113-154)."""
        pm, g, rows = self.pm, self.geo, self.rows
        if not self.use_cond:
            raise RuntimeError("legacy UNet1D needs motion_f / text_f (every block cross-attends)")
        n_down = len(pm.dims)
        kv_iter = iter(zip(self.kv, self.vt))
        self._build_kv_ops()
        # TimestepEmbedding output itself (one SiLU): time_proj has no activation in front
        self._side_op = True
        self._add(ops.time_embed, self.t_in, pm.time_w, pm.time_b, self.silu_temb, self.t_rows,
                  pm.time_dim, False, meta={"kind": "time_mlp", "flops": 0})
        self._add(ops.film, self.silu_temb, pm.film_w, pm.film_b, self.film, self.t_rows,
                  pm.time_dim, pm.film_cols)
        self._side_op = False
        self._hold_join = True
        self._add(ops.ingest_x_f32 if self.fp32 else ops.ingest_x, self.x_in, self.x_slab,
                  self.batch, self.copies, pm.in_dim, self.t, g.Tp[0], pm.in_pad, self.arena,
                  meta={"kind": "ingest_x", "flops": 0})
        cur = self._view(self._pp[0], g.M[0], pm.base)
        cur_st = self._stats(rows, 0, pm.base, 8)
        # CFG copies are identical up to the first (attention) block: input_proj runs on the cond
        # copies only, the first block's uncond rows read from there (see _build)
        if (self.copies == 2 and self.uncond_rows == self.batch and self.uniform_t
                and os.environ.get("LM2A_SHARE_CFG_ROWS", "1") != "0"):
            self._shared_rows = self.batch
        r0 = self._shared_rows
        m0 = (rows - r0) * g.Tp[0]
        self._conv([Seg(self.x_slab, pm.in_pad, pm.in_pad, TAPS_K1, m0, r0 * g.Tp[0] * pm.in_pad)],
                   pm.w_in, pm.b_in, pm.base, m0, g.Tp[0], g.T[0], cur, pm.base,
                   out_chan_off=r0 * g.Tp[0] * pm.base, k_real=pm.in_dim, stats=cur_st.view(r0, 0))
        cur_c, pp = pm.base, 1
        self._hold_join, self._force_join = False, True
        # concat slab of level l: [transposed-conv output (dims[l]) | skip (c_l)], normalised as
        # a whole by the decoder block. The transposed conv writes it as two launches (even / odd
        # slots) in the geometry of level l+1; both add into the same exact sums.
        self.cat_st = [self._stats(rows, lvl, pm.cat_width(lvl), 8) for lvl in range(n_down)]
        for lvl, (p, wd, bd) in enumerate(pm.downs):
            dim, wcat = pm.dims[lvl], pm.cat_width(lvl)
            self._resblock(p, lvl, cur, cur_c, 0, cur_st, self.cat[lvl], wcat, dim,
                           self.cat_st[lvl], next(kv_iter))
            nxt = self._view(self._pp[pp], g.M[lvl + 1], dim)
            pp ^= 1
            nxt_st = self._stats(rows, lvl + 1, dim, 8)
            self._conv([Seg(self.cat[lvl], wcat, p.cout, TAPS_K4S2, g.M[lvl], dim)], wd, bd, dim,
                       g.M[lvl + 1], g.Tp[lvl + 1], g.T[lvl + 1], nxt, dim, stats=nxt_st)
            cur, cur_c, cur_st = nxt, dim, nxt_st
        lvl = n_down
        out = self._view(self._pp[pp], g.M[lvl], cur_c)
        pp ^= 1
        self._resblock(pm.mid, lvl, cur, cur_c, 0, cur_st, out, cur_c, 0,
                       self._stats(rows, lvl, cur_c, 8), next(kv_iter))
        cur = out
        for i, (we, wo, bu, p) in enumerate(pm.ups):
            lvl = n_down - 1 - i
            dim, wcat = pm.dims[lvl], pm.cat_width(lvl)
            lo = lvl + 1
            assert 2 * g.T[lo] <= g.T[lvl] and g.Tp[lvl] == 2 * g.Tp[lo]
            # ConvTranspose1d k4 s2 p1: even slots 2m and odd slots 2m+1 of the level-l slab are
            # the two halves of a row of the [M_{l+1}, 2 * wcat] view of the concat slab; slots
            # past 2 * T_{l+1} are written as zero (= the reference's F.pad, unet1d.py:141-147)
            for half, w in enumerate((we, wo)):
                self._conv([Seg(cur, cur_c, cur_c, TAPS_K3, g.M[lo])], w, bu, dim, g.M[lo],
                           g.Tp[lo], g.T[lo], self.cat[lvl], 2 * wcat, out_chan_off=half * wcat,
                           k_real=2 * cur_c, stats=self.cat_st[lvl].view(0, 0))
            out = self._view(self._pp[pp], g.M[lvl], wcat)
            pp ^= 1
            self._resblock(p, lvl, self.cat[lvl], wcat, 0, self.cat_st[lvl], out, wcat, 0,
                           self._stats(rows, lvl, wcat, 8), next(kv_iter))
            cur, cur_c = out, wcat
        self._conv([Seg(cur, cur_c, cur_c, TAPS_K1, g.M[0])], pm.w_out, pm.b_out, pm.in_dim,
                   g.M[0], g.Tp[0], g.T[0], self.eps, 0, out_mode=OUT_F32_NCT, block_n=128)

    def _build(self):
        pm, g, rows = self.pm, self.geo, self.rows
        n_down = len(pm.dims)
        kv_iter = iter(zip(self.kv, self.vt)) if self.use_cond else iter(())
        next_kv = lambda p: next(kv_iter) if (p.attn and self.use_cond) else None  # noqa: E731

        self._build_kv_ops()

        # timestep embedding + all FiLM tables, once per step: a parallel branch (side stream)
        # next to x ingest + in_proj; the first ResBlock joins it
        self._side_op = True
        self._add(ops.time_mlp, self.t_in, pm.time_w, pm.time_b, self.silu_temb, self.t_rows,
                  pm.time_dim)
        self._add(ops.film, self.silu_temb, pm.film_w, pm.film_b, self.film, self.t_rows,
                  pm.time_dim, pm.film_cols)
        self._side_op = False
        self._hold_join = True
        # x -> bf16 slab (CFG row duplication happens here), in_proj
        self._add(ops.ingest_x_f32 if self.fp32 else ops.ingest_x, self.x_in, self.x_slab,
                  self.batch, self.copies, pm.in_dim, self.t, g.Tp[0], pm.in_pad, self.arena,
                  meta={"kind": "ingest_x", "flops": 0})
        groups_of = lambda blk_gn: blk_gn[2]  # noqa: E731
        cur = self._view(self._pp[0], g.M[0], pm.base)
        first = pm.downs[0][0][0]
        cur_st = self._stats(rows, 0, pm.base, groups_of(first.gn1))
        # CFG: the uncond and cond copies of a clip are identical until the first attention block
        # (nothing depends on the conditions before it). With the uncond shortcut on, in_proj and
        # the leading attention-free ResBlocks run on the cond copies (rows >= B) only and the
        # first attention block's uncond rows read their input from there.
        self._shared_rows = 0
        if (self.copies == 2 and self.use_cond and self.uncond_rows == self.batch
                and self.uniform_t and not first.attn
                and os.environ.get("LM2A_SHARE_CFG_ROWS", "1") != "0"):
            self._shared_rows = self.batch
        r0 = self._shared_rows
        m0 = (rows - r0) * g.Tp[0]
        self._conv([Seg(self.x_slab, pm.in_pad, pm.in_pad, TAPS_K1, m0, r0 * g.Tp[0] * pm.in_pad)],
                   pm.w_in, pm.b_in, pm.base, m0, g.Tp[0], g.T[0], cur, pm.base,
                   out_chan_off=r0 * g.Tp[0] * pm.base, k_real=pm.in_dim, stats=cur_st.view(r0, 0))
        cur_c, pp = pm.base, 1
        self._hold_join, self._force_join = False, True

        # the concat slab of level l is normalised as a whole by the first up block of that
        # level; its two halves are written by different kernels into one Stats buffer
        up_first = {n_down - 1 - i: blocks[0] for i, (_, _, blocks) in enumerate(pm.ups)}
        self.cat_st = [self._stats(rows, lvl, 2 * pm.dims[lvl], groups_of(up_first[lvl].gn1))
                       for lvl in range(n_down)]

        def consumer_stats(lvl, c, nxt_block):
            return self._stats(rows, lvl, c, groups_of(nxt_block.gn1))

        # down path: the last block of each stage writes straight into the second half of
        # the level's concat slab (= the skip connection), the stride-2 conv reads it there
        for lvl, (blocks, wd, bd) in enumerate(pm.downs):
            dim = pm.dims[lvl]
            cur_ld, cur_off = cur_c, 0
            for bi, p in enumerate(blocks):
                last = bi == len(blocks) - 1
                if last:
                    out, out_ld, out_off, out_st = self.cat[lvl], 2 * dim, dim, self.cat_st[lvl]
                else:
                    out = self._view(self._pp[pp], g.M[lvl], p.cout)
                    out_ld, out_off = p.cout, 0
                    out_st = consumer_stats(lvl, p.cout, blocks[bi + 1])
                    pp ^= 1
                self._resblock(p, lvl, cur, cur_ld, cur_off, cur_st, out, out_ld, out_off, out_st,
                               next_kv(p))
                cur, cur_ld, cur_off, cur_c, cur_st = out, out_ld, out_off, p.cout, out_st
            nxt = self._view(self._pp[pp], g.M[lvl + 1], dim)
            pp ^= 1
            nxt_block = pm.downs[lvl + 1][0][0] if lvl + 1 < n_down else pm.mid[0]
            nxt_st = self._stats(rows, lvl + 1, dim, groups_of(nxt_block.gn1))
            self._conv([Seg(cur, cur_ld, dim, TAPS_K4S2, g.M[lvl], cur_off)], wd, bd, dim,
                       g.M[lvl + 1], g.Tp[lvl + 1], g.T[lvl + 1], nxt, dim, stats=nxt_st)
            cur, cur_c, cur_st = nxt, dim, nxt_st

        lvl = n_down
        for i, p in enumerate(pm.mid):
            out = self._view(self._pp[pp], g.M[lvl], p.cout)
            pp ^= 1
            # (the last mid block feeds the up-conv, which has no GroupNorm: stats unused there)
            out_st = self._stats(rows, lvl, p.cout, groups_of(pm.mid[min(i + 1, len(pm.mid) - 1)].gn1))
            self._resblock(p, lvl, cur, cur_c, 0, cur_st, out, p.cout, 0, out_st, next_kv(p))
            cur, cur_c, cur_st = out, p.cout, out_st

        # up path: interp x2 -> conv k3 into the first half of the concat slab (slots past
        # 2*T_{l+1} stay zero = F.pad), then the ResBlocks read the concat slab directly
        for i, (wu, bu, blocks) in enumerate(pm.ups):
            lvl = n_down - 1 - i
            dim = pm.dims[lvl]
            t_up = 2 * g.T[lvl + 1]
            assert t_up <= g.T[lvl]
            # UpSampleConv: narrow levels interpolate x2 in the conv's operand path (the
            # transform warps build each operand block from the low-resolution slab); wide levels
            # (MMA-bound up-conv) take one streaming upsample pass and a plain launch. Same
            # arithmetic either way.
            if cur_c <= self.up_xf_max_c:
                self._conv([Seg(cur, cur_c, cur_c, TAPS_K3, g.M[lvl + 1])], wu, bu, dim, g.M[lvl],
                           g.Tp[lvl], t_up, self.cat[lvl], 2 * dim,
                           stats=self.cat_st[lvl].view(0, 0), up2x=(g.Tp[lvl + 1], g.T[lvl + 1]))
            else:
                xup = self._view(self._h1, g.M[lvl], cur_c)
                self._add(ops.upsample2x, cur, cur_c, xup, cur_c, rows, g.Tp[lvl + 1],
                          g.T[lvl + 1], g.Tp[lvl], cur_c, meta={"kind": "upsample2x", "flops": 0})
                self._conv([Seg(xup, cur_c, cur_c, TAPS_K3, g.M[lvl])], wu, bu, dim, g.M[lvl],
                           g.Tp[lvl], t_up, self.cat[lvl], 2 * dim,
                           stats=self.cat_st[lvl].view(0, 0))
            cur, cur_ld, cur_c, cur_st = self.cat[lvl], 2 * dim, 2 * dim, self.cat_st[lvl]
            for bi, p in enumerate(blocks):
                out = self._view(self._pp[pp], g.M[lvl], p.cout)
                pp ^= 1
                if bi + 1 < len(blocks):
                    out_groups = groups_of(blocks[bi + 1].gn1)
                elif i + 1 < len(pm.ups):
                    out_groups = groups_of(p.gn2)  # feeds the next up-conv: no GroupNorm reads it
                else:
                    out_groups = pm.gn_out[2]
                out_st = self._stats(rows, lvl, p.cout, out_groups)
                self._resblock(p, lvl, cur, cur_ld, 0, cur_st, out, p.cout, 0, out_st, next_kv(p))
                cur, cur_ld, cur_c, cur_st = out, p.cout, p.cout, out_st

        # out_proj: GN + SiLU + 1x1 conv, written as fp32 [rows, in_dim, T]
        a, a_ld, a_off, in_gn = self._gn_operand(cur, cur_c, 0, cur_st, pm.gn_out, rows, g.Tp[0],
                                                 g.T[0], cur_c)
        self._conv([Seg(a, a_ld, cur_c, TAPS_K1, g.M[0], a_off)], pm.w_out, pm.b_out, pm.in_dim,
                   g.M[0], g.Tp[0], g.T[0], self.eps, 0, out_mode=OUT_F32_NCT, block_n=128,
                   in_gn=in_gn)

    # -- execution --------------------------------------------------------------------
    def set_conditions(self, motion_f, text_f, kv_slot):
        """motion_f / text_f: fp32 [nslots, lk, cond_dim]; builds every layer's K/V cache
        (step-invariant: reference recomputes them every step, cross_attention.py:46-61)."""
        assert self.use_cond
        n, lk, c = self.nslots, self.lk, self.pm.cond_dim
        if tuple(motion_f.shape) != (n, lk, c) or tuple(text_f.shape) != (n, lk, c):
            raise RuntimeError(f"conditions must be ({n}, {lk}, {c}); got "
                               f"{tuple(motion_f.shape)} / {tuple(text_f.shape)}")
        ingest = ops.ingest_seq_f32 if self.fp32 else ops.ingest_seq
        ingest(motion_f.contiguous().float(), self.cond_m, n, lk, c, lk, c)
        ingest(text_f.contiguous().float(), self.cond_t, n, lk, c, lk, c)
        for fn, args in self.kv_ops:
            fn(*args)
        self.kv_slot.copy_(kv_slot.to(torch.int32))
        self._fill_const_text()

    def _fill_const_text(self):
        """Selects the launch list for this batch. The lyrics stream counts as constant in time
        when, for every cache slot, all Lk rows of the projected condition are bit-identical (a
        host-side decision, once per batch). Then each block's text-stream attention output is
        the V row of the clip (softmax over identical keys is uniform): its image under the folded
        output projection becomes a per-clip shift of the block's output GEMM (_make_ct_table),
        and the per-step launches skip that stream."""
        self.const_text = False
        if not (self.use_cond and self.allow_const_text and self.lk > 1):
            return
        c = self.cond_t.view(self.nslots, self.lk, -1)
        if not bool((c[:, 1:] == c[:, :1]).all()):
            return
        for fill in self.ct_ops:
            fill()
        self.const_text = True

    def cond_slabs(self, first_slot=0):
        """bf16 views [ (nslots - first_slot) * lk, cond_dim ] of the motion / lyrics condition
        slabs from cache slot `first_slot` on: CondProjection.project_raw writes the projected
        conditions straight into them; `build_kv` then builds the caches."""
        assert self.use_cond
        if self.fp32:
            raise RuntimeError("the fp32 validation path takes projected conditions through "
                               "set_conditions (CondProjection.project_raw writes bf16 slabs)")
        o = first_slot * self.lk
        return self.cond_m[o:], self.cond_t[o:]

    def build_kv(self, kv_slot):
        """K/V cache build from the condition slabs as they are (see cond_slabs)."""
        assert self.use_cond
        for fn, args in self.kv_ops:
            fn(*args)
        self.kv_slot.copy_(kv_slot.to(torch.int32))
        self._fill_const_text()

    def run(self, skip_ingest=False):
        """Launches the plan on torch's current stream (skip_ingest: the input slab and the
        cleared statistics arena were already produced by the previous step's lm2a_cfg_step). Ops tagged `side` (the uncond rows'
        `skip(x) + const` kernels and the timestep / FiLM tables: small launches that nothing
        on the main chain needs until the tagged `join` op) go to a second stream, forked from
        and joined back into the current one — under CUDA Graph capture they become a parallel
        branch, off the critical path of the cond rows' pipeline."""
        todo = [op for op in self.ops if not (skip_ingest and op[2]["kind"] == "ingest_x")]
        if not self.use_side_stream:
            for fn, args, _ in todo:
                fn(*args)
            return self.eps
        main = torch.cuda.current_stream(self.dev)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        side, pending = self._side, False
        for fn, args, meta in todo:
            if meta.get("side"):
                side.wait_stream(main)  # fork after everything enqueued so far
                with torch.cuda.stream(side):
                    fn(*args)
                pending = True
                continue
            if pending and meta.get("join"):
                main.wait_stream(side)
                pending = False
            fn(*args)
        if pending:
            main.wait_stream(side)
        return self.eps

    def _time_op(self, fn, args, iters, reps=3):
        """Device time of one launch: `iters` back-to-back launches captured in a CUDA Graph,
        replayed twice untimed (the capture leaves the GPU idle for milliseconds: without a warm
        replay a short burst is timed while the clocks ramp up - the per-launch sum then exceeded
        the measured step on some boxes) and `reps` times between CUDA events on the launching
        stream."""
        fn(*args)  # warm (first-launch attribute setup must not happen under capture)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn(*args)
        g.replay()
        g.replay()
        stream = torch.cuda.current_stream(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            g.replay()
        e1.record(stream)
        torch.cuda.synchronize(self.dev)
        del g
        return e0.elapsed_time(e1) * 1e-3 / (iters * reps)

    def autotune(self, iters=5, margin=0.97):
        """Picks the tile shape of every implicit-GEMM launch by measurement instead of the wave
        model: (block_n, cta_group) in {auto, 128x1, 256x1, 128x2 (CTA pair), 256x2}, each timed
        on this plan's own operands; a non-default shape is kept only if it is > 3 % faster. The
        K order of the accumulation and the GroupNorm partial-sum slices do not depend on the tile
        shape, so the results are unchanged. Returns [(launch index, block_n, cta_group, us)]."""
        if getattr(self, "_tuned", False):
            return []
        self._tuned = True
        picked = []
        for idx, (fn, args, meta) in enumerate(self.ops):
            if meta["kind"] != "conv_gemm":
                continue
            desc = args[0]
            best = None
            for bn, cg in ((0, 0), (128, 1), (256, 1), (128, 2), (256, 2)):
                if bn and desc.n_pad % bn != 0:
                    continue
                desc.block_n, desc.cta_group = bn, cg
                try:
                    t = self._time_op(fn, args, iters)
                except RuntimeError:   # shape not admissible for this launch (fused GroupNorm)
                    continue
                if best is None or t < best[0] * margin:
                    best = (t, bn, cg)
            desc.block_n, desc.cta_group = best[1], best[2]
            if best[1] or best[2]:
                picked.append((idx, best[1], best[2], best[0] * 1e6))
        return picked

    def flops(self):
        """Algorithmic FLOPs of one forward over all rows (K/V hoisted, out_proj.fuse folded):
        what the roofline credits."""
        return sum(meta["flops"] for _, _, meta in self.ops)

    def flops_executed(self):
        """FLOPs the launches execute on valid slots (>= flops(): the composed conv2 . Q GEMM)."""
        return sum(meta.get("flops_executed", meta["flops"]) for _, _, meta in self.ops)

    def profile(self, iters=10):
        """Per-launch device time for bench.py's roofline line and profiles/. Each launch is
        captured `iters` times back to back in a CUDA Graph and the replay is timed with CUDA
        events on the launching stream, so short kernels are not measured at the Python launch
        rate. Repeats of one launch re-read the same operands (L2-warm; the slabs of the big
        layers exceed what stays resident next to the weights). Returns [(kind, meta, seconds)]
        in launch order."""
        return [(meta["kind"], meta, self._time_op(fn, args, iters)) for fn, args, meta in self.ops]


def params_fingerprint(model):
    """(data_ptr, version) of every parameter: changes whenever a parameter is re-allocated or
    written in place through the tensor itself (load_state_dict, optimizer steps, EMA copy_,
    the fused Adan step, .to()). Writes through `.data` / raw pointers do not bump the version
    counter: call `model.refresh()` after those."""
    return tuple((p.data_ptr(), p._version) for p in model.parameters())


class UNetEngine:
    def __init__(self, model, precision="bf16"):
        p = next(model.parameters())
        ops.require_device(p)
        self.dev = p.device
        self.precision = precision
        self.fingerprint = params_fingerprint(model)
        legacy = hasattr(model, "input_proj")  # models/unet1d.py names vs unet1d_ultimate.py
        wdt = torch.float32 if precision == "fp32" else BF16
        _PACK_DTYPE[0] = wdt
        try:
            with torch.cuda.device(self.dev):
                self.pm = PackedLegacy(model, self.dev) if legacy else PackedModel(model, self.dev)
        finally:
            _PACK_DTYPE[0] = BF16
        self.pm.wdt = wdt
        self.plans = {}
        self._cond_key = None
        self._cond_ref = None

    def plan(self, rows, t, lk, nslots, copies=1, use_cond=True, uniform_t=False, uncond_rows=0):
        key = (rows, t, lk, nslots, copies, use_cond, uniform_t, uncond_rows)
        if key not in self.plans:
            with torch.cuda.device(self.dev):
                self.plans[key] = UNetPlan(self.pm, rows, t, lk, nslots, copies, use_cond,
                                           self.dev, uniform_t, uncond_rows)
        return self.plans[key]

    def forward(self, x, t, motion_f=None, text_f=None):
        ops.require_device(x)
        if x.dim() != 3 or x.shape[1] != self.pm.in_dim:
            raise RuntimeError(f"expected x of shape (B, {self.pm.in_dim}, T); got {tuple(x.shape)}")
        b, _, tlen = x.shape
        use_cond = motion_f is not None and text_f is not None
        lk = motion_f.shape[1] if use_cond else 0
        if use_cond and (motion_f.shape[0] != b or text_f.shape[0] != b
                         or text_f.shape[1] != lk):
            raise RuntimeError("motion_f / text_f must be (B, Lk, cond_dim) with matching B, Lk")
        plan = self.plan(b, tlen, lk, b if use_cond else 0, 1, use_cond)
        with torch.cuda.device(self.dev):   # launches go to THIS device's current stream
            if use_cond:
                # The K/V caches are step-invariant: rebuilt only when the conditions change.
                # The key holds the tensors themselves (strong references), so their storage
                # cannot be recycled for another clip's conditions while the key is alive.
                key = (id(plan), motion_f.data_ptr(), motion_f._version, text_f.data_ptr(),
                       text_f._version)
                same = (key == self._cond_key and self._cond_ref is not None
                        and self._cond_ref[0] is motion_f and self._cond_ref[1] is text_f)
                if not same:
                    plan.set_conditions(motion_f, text_f,
                                        torch.arange(b, device=self.dev, dtype=torch.int32))
                    self._cond_key, self._cond_ref = key, (motion_f, text_f)
            if not torch.is_tensor(t):
                t = torch.full((b,), int(t), device=self.dev, dtype=torch.int64)
            plan.x_in.copy_(x)
            plan.t_in.copy_(t.reshape(-1).expand(b) if t.numel() == 1 else t)
            return plan.run().clone()
