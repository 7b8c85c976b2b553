"""Python-side launchers for the C ABI (include/lm2a_b200.h).

Tensors are only carriers of device memory here: every function passes raw device
pointers, sizes and torch's *current* CUDA stream to liblm2a_b200.so, so the launches
are captured when called under torch.cuda.graph().
"""
import ctypes

import torch

from . import _lib
from ._lib import (OUT_BF16_SLAB, OUT_F32_NCT, TAPS_K1, TAPS_K3, TAPS_K4S2,  # noqa: F401
                   ConvDesc)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def on_device_of(t):
    """Context manager: launches go to the current stream of the tensor's device (a process may
    drive several GPUs)."""
    return torch.cuda.device(t.device)


def _ptr(t, elem_offset=0):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr() + elem_offset * t.element_size())


def require_device(t):
    if not t.is_cuda:
        raise RuntimeError("lm2a_b200 runs on CUDA sm_100a only (no CPU path); got a CPU tensor")
    _lib.check(_lib.load().lm2a_check_device(), "lm2a_check_device")


def launch_count():
    return int(_lib.load().lm2a_launch_count())


def reset_launch_count():
    _lib.load().lm2a_reset_launch_count()


class StatsArena:
    """One zero-initialised int64 region holding every GroupNorm statistics buffer of a launch
    plan; lm2a_ingest_x (the first kernel of a step) clears the used part, the epilogues
    accumulate into it. Fixed capacity: addresses are handed out while the plan is built."""

    def __init__(self, dev, capacity_words):
        self.dev, self.words = dev, 0
        self.buf = torch.zeros(max(2, (capacity_words + 1) // 2 * 2), dtype=torch.int64,
                               device=dev)

    def reserve(self, words):
        off = self.words
        self.words += (words + 1) // 2 * 2      # keep every buffer 16-byte aligned
        if self.words > self.buf.numel():
            raise RuntimeError("GroupNorm statistics arena exhausted")
        return off

    def tensor(self):
        return self.buf

    def used(self):
        return self.buf[: self.words]


class Stats:
    """Exact GroupNorm sums of a slab [rows, tp, c] normalised in `groups` groups: int64
    [rows, groups, 2] = {sum * 2^24, sum of squares * 2^20}, accumulated by the producing
    kernel's epilogue (integer adds: independent of tile shape and batch position), read by the
    consuming conv's operand transform or by gn_apply. `view(row0, chan0)` addresses a row /
    channel sub-range (a launch over part of the rows, or one half of a concat slab). The
    buffer must be zero before the producers of a step run: standalone objects are zeroed at
    construction (one use), plan-owned ones live in a StatsArena that ingest_x clears."""

    def __init__(self, rows, c, groups, dev, arena=None, _base=None, row0=0, chan0=0):
        self.rows, self.c, self.groups = rows, c, groups
        self.cg = c // groups
        assert c % groups == 0 and self.cg % 8 == 0
        self.row0, self.chan0 = row0, chan0
        if _base is not None:
            self._base = _base
        elif arena is not None:
            off = arena.reserve(rows * groups * 2)
            self._base = (arena, off)
        else:
            self._base = (torch.zeros(rows * groups * 2, dtype=torch.int64, device=dev), 0)

    @property
    def buf(self):
        owner, off = self._base
        t = owner.tensor() if isinstance(owner, StatsArena) else owner
        return t[off: off + self.rows * self.groups * 2]

    def zero_(self):
        self.buf.zero_()

    def view(self, row0=0, chan0=0):
        assert chan0 % 8 == 0
        return Stats(self.rows, self.c, self.groups, None, None, self._base, self.row0 + row0,
                     self.chan0 + chan0)

    def ptr(self):
        """Address of clip-row row0's sums; the channel offset travels separately (stats_c0):
        a producer whose channels start inside a group still adds into the right one."""
        return self.buf.data_ptr() + self.row0 * self.groups * 2 * 8

    def sums(self):
        """fp64 [rows, groups, 2] = (sum, sum of squares) decoded from the fixed-point words."""
        v = self.buf.view(self.rows, self.groups, 2).double()
        return torch.stack([v[..., 0] / 2.0 ** 24, v[..., 1] / 2.0 ** 20], dim=-1)


class Seg:
    """One K segment of the implicit GEMM: a bf16 slab view (tensor + channel offset)."""

    def __init__(self, slab, ld, cin, taps, rows, chan_off=0):
        self.slab, self.ld, self.cin, self.taps, self.rows, self.chan_off = (
            slab, ld, cin, taps, rows, chan_off)


def make_conv_desc(segs, w, bias, n_valid, m, tp, t_valid, out, out_ld, out_chan_off=0,
                   film=None, film_col=0, film_shift_off=0, film_bcast=False, film_row=0,
                   residual=None, res_ld=0,
                   res_chan_off=0, out_mode=OUT_BF16_SLAB, block_n=0, stats=None, cta_group=0,
                   in_gn=None, up2x=None, k_order=0):
    """Builds the (reusable) descriptor of one lm2a_conv1d_bf16 launch. Keeps the tensors
    alive by attaching them to the descriptor object.
    in_gn = (Stats view of segs[0]'s slab, gamma, beta, eps, silu): GroupNorm (+ SiLU) applied to
    the first segment's operand tiles on the fly.
    up2x = (tp_in, t_in): segs[0] is the LOW-resolution slab (rows = m / 2); its x2 linear
    upsampling (align_corners) is computed inside the conv's operand path."""
    d = ConvDesc()
    for i, s in enumerate(segs):
        d.seg[i].x = s.slab.data_ptr() + s.chan_off * s.slab.element_size()
        d.seg[i].rows = s.rows
        d.seg[i].ld = s.ld
        d.seg[i].cin = s.cin
        d.seg[i].taps = s.taps
    d.w = w.data_ptr()
    d.n_pad = w.shape[0]
    d.n_valid = n_valid
    d.m = m
    d.tp = tp
    d.t_valid = t_valid
    d.bias = bias.data_ptr()
    if film is not None:
        d.film = film.data_ptr() + (film_col + (0 if film_bcast else film_row * film.shape[1])) * 4
        d.film_ld = 0 if film_bcast else film.shape[1]  # 0: one table row for all clip-rows
        d.film_shift_off = film_shift_off
    if residual is not None:
        d.residual = residual.data_ptr() + res_chan_off * residual.element_size()
        d.res_ld = res_ld
    d.out_mode = out_mode
    d.out = out.data_ptr() + out_chan_off * out.element_size()
    d.out_ld = out_ld
    d.block_n = block_n
    d.cta_group = cta_group
    if stats is not None:  # Stats view: exact GroupNorm sums of the output
        d.stats = stats.ptr()
        d.stats_pitch, d.stats_cg, d.stats_c0 = stats.groups, stats.cg, stats.chan0
    if in_gn is not None:
        st, gamma, beta, eps, silu = in_gn
        assert st.c == segs[0].cin and st.chan0 == 0
        d.in_gn_stats = st.ptr()
        d.in_gn_gamma, d.in_gn_beta = gamma.data_ptr(), beta.data_ptr()
        d.in_gn_pitch, d.in_gn_groups = st.groups, st.groups
        d.in_gn_eps, d.in_gn_silu = eps, 1 if silu else 0
    if up2x is not None:
        d.in_up_tp, d.in_up_t = up2x
    d.k_order = k_order   # 1: accumulate in the operand-transform launches' order (bit-for-bit)
    d._keep = (segs, w, bias, film, residual, out, stats, in_gn)
    return d


def in_gn_supported(tp, groups):
    """Whether the conv's operand transform can normalise clips of `tp` slots in `groups`
    groups (a 130-slot tile must touch at most 512 (clip-row, group) pairs)."""
    return ((128 + 1) // tp + 2) * groups <= 512


def conv1d(desc):
    _lib.check(_lib.load().lm2a_conv1d_bf16(_stream(), ctypes.byref(desc)), "lm2a_conv1d_bf16")


# ---- fp32 validation path (csrc/ref_f32.cu): same descriptors, fp32 slabs and weights ----------
def conv1d_f32(desc):
    _lib.check(_lib.load().lm2a_conv1d_f32(_stream(), ctypes.byref(desc)), "lm2a_conv1d_f32")


def cross_attn_f32(q, q_ld, o, o_ld, kv_m, kv_t, kv_ld, kv_slot, slots, rows, tp, t_valid, lk, e,
                   heads, n_streams=2):
    _lib.check(_lib.load().lm2a_cross_attn_f32(
        _stream(), _ptr(q), q_ld, _ptr(o), o_ld, _ptr(kv_m), _ptr(kv_t), kv_ld,
        kv_slot if isinstance(kv_slot, ctypes.c_void_p) else _ptr(kv_slot), slots, rows, tp,
        t_valid, lk, e, heads, n_streams), "lm2a_cross_attn_f32")


def bias_add_f32(x, x_ld, x_off, y, y_ld, y_off, bias, slots, tp, t_valid, c, stats=None):
    _lib.check(_lib.load().lm2a_bias_add_f32(
        _stream(), _ptr(x, x_off), x_ld, _ptr(y, y_off), y_ld, _ptr(bias), slots, tp, t_valid, c,
        ctypes.c_void_p(stats.ptr()) if stats is not None else None,
        stats.groups if stats is not None else 0, stats.cg if stats is not None else 0,
        stats.chan0 if stats is not None else 0), "lm2a_bias_add_f32")


def ingest_x_f32(x, slab, batch, copies, c, t, tp, ld, zero=None):
    zt = zero.used() if isinstance(zero, StatsArena) else zero
    zbytes = 0 if zt is None else zt.numel() * zt.element_size()
    _lib.check(_lib.load().lm2a_ingest_x_f32(_stream(), _ptr(x), _ptr(slab), batch, copies, c, t,
                                             tp, ld, _ptr(zt), zbytes), "lm2a_ingest_x_f32")


def ingest_seq_f32(x, slab, rows, t, c, tp, ld):
    _lib.check(_lib.load().lm2a_ingest_seq_f32(_stream(), _ptr(x), _ptr(slab), rows, t, c, tp, ld),
               "lm2a_ingest_seq_f32")


def gn_silu(x, x_ld, y, y_ld, gamma, beta, rows, tp, t_valid, c, groups, eps=1e-5, silu=True,
            x_chan_off=0, y_chan_off=0):
    _lib.check(_lib.load().lm2a_gn_silu_bf16(
        _stream(), _ptr(x, x_chan_off), x_ld, _ptr(y, y_chan_off), y_ld, _ptr(gamma), _ptr(beta),
        rows, tp, t_valid, c, groups, eps, 1 if silu else 0), "lm2a_gn_silu_bf16")


def cross_attn(q, q_ld, o, o_ld, k_m, vt_m, k_t, vt_t, k_ld, vt_ld, kv_slot, slots, rows, tp,
               t_valid, lk, e, heads, n_streams=2, q_off=0, o_off=0):
    """q_off / o_off: element offsets into the q / o slabs (row sub-range of a shared slab)."""
    _lib.check(_lib.load().lm2a_cross_attn_streams_bf16(
        _stream(), _ptr(q, q_off), q_ld, _ptr(o, o_off), o_ld, k_m, vt_m, k_t, vt_t, k_ld, vt_ld,
        kv_slot if isinstance(kv_slot, ctypes.c_void_p) else _ptr(kv_slot), slots, rows,
        tp, t_valid, lk, e, heads, n_streams), "lm2a_cross_attn_bf16")


def cond_attn_supported(lk):
    """Whether lm2a_cross_attn_cond_bf16 can keep `lk` condition frames resident."""
    return bool(_lib.load().lm2a_cross_attn_cond_supported(int(lk)))


def cross_attn_cond(q, q_ld, o, o_ld, cond_m, cond_t, cond_ld, kv_slot, slots, rows, tp, t_valid,
                    lk, heads, n_streams=2, q_off=0, o_off=0):
    """softmax(q'_h C^T) C per head against the raw condition slabs (head dim = cond width = 128)."""
    _lib.check(_lib.load().lm2a_cross_attn_cond_bf16(
        _stream(), _ptr(q, q_off), q_ld, _ptr(o, o_off), o_ld, cond_m, cond_t, cond_ld,
        kv_slot if isinstance(kv_slot, ctypes.c_void_p) else _ptr(kv_slot), slots, rows, tp,
        t_valid, lk, heads, n_streams), "lm2a_cross_attn_cond_bf16")


def cross_attn_tail(q, q_ld, o, o_ld, k_m, v_m, k_t, v_t, k_ld, v_ld, kv_slot, slots, rows, tp,
                    t0, n_tail, lk, e, heads, n_streams=2, shared_kv=False):
    """The n_tail (<= 8) query rows from t0 on, per (clip-row, stream, head), on the CUDA cores:
    the rows T mod 128 leaves over. Keys and values row-major [slots*lk, ld] (pointers);
    shared_kv: every head reads channels [0, dh) (the raw condition slabs)."""
    _lib.check(_lib.load().lm2a_cross_attn_tail_bf16(
        _stream(), _ptr(q), q_ld, _ptr(o), o_ld, k_m, v_m, k_t, v_t, k_ld, v_ld,
        kv_slot if isinstance(kv_slot, ctypes.c_void_p) else _ptr(kv_slot), slots, rows, tp, t0,
        n_tail, lk, e, heads, n_streams, 1 if shared_kv else 0), "lm2a_cross_attn_tail_bf16")


def transpose_kv(src, src_ld, src_off, dst, dst_ld, slots, lk, c):
    _lib.check(_lib.load().lm2a_transpose_kv_bf16(_stream(), _ptr(src, src_off), src_ld, _ptr(dst),
                                                  dst_ld, slots, lk, c), "lm2a_transpose_kv_bf16")


def time_mlp(t, w, b, out, rows, dim):
    _lib.check(_lib.load().lm2a_time_mlp(_stream(), _ptr(t), _ptr(w), _ptr(b), _ptr(out), rows,
                                         dim), "lm2a_time_mlp")


def time_embed(t, w, b, out, rows, dim, fold_silu):
    _lib.check(_lib.load().lm2a_time_embed(_stream(), _ptr(t), _ptr(w), _ptr(b), _ptr(out), rows,
                                           dim, 1 if fold_silu else 0), "lm2a_time_embed")


def film(s, w, b, out, rows, dim, cols):
    _lib.check(_lib.load().lm2a_film(_stream(), _ptr(s), _ptr(w), _ptr(b), _ptr(out), rows, dim,
                                     cols), "lm2a_film")


def ingest_x(x, slab, batch, copies, c, t, tp, ld, zero=None):
    """zero: optional StatsArena (or int64 tensor) cleared by the same launch."""
    zt = zero.used() if isinstance(zero, StatsArena) else zero
    zbytes = 0 if zt is None else zt.numel() * zt.element_size()
    _lib.check(_lib.load().lm2a_ingest_x(_stream(), _ptr(x), _ptr(slab), batch, copies, c, t, tp,
                                         ld, _ptr(zt), zbytes), "lm2a_ingest_x")


def ingest_seq(x, slab, rows, t, c, tp, ld):
    _lib.check(_lib.load().lm2a_ingest_seq(_stream(), _ptr(x), _ptr(slab), rows, t, c, tp, ld),
               "lm2a_ingest_seq")


def resample_seq(x, lens, out_f32, out_slab, rows, t_in_max, c, t_out, tp, ld):
    _lib.check(_lib.load().lm2a_resample_seq(_stream(), _ptr(x), _ptr(lens), _ptr(out_f32),
                                             _ptr(out_slab), rows, t_in_max, c, t_out, tp, ld),
               "lm2a_resample_seq")


def upsample2x(x, x_ld, y, y_ld, rows, tp_in, t_in, tp_out, c):
    _lib.check(_lib.load().lm2a_upsample2x_bf16(_stream(), _ptr(x), x_ld, _ptr(y), y_ld, rows,
                                                tp_in, t_in, tp_out, c), "lm2a_upsample2x_bf16")


def cfg_posterior(x, eps, noise, sched, t_dev, ticket, batch, elems_per_clip, guidance, guided,
                  advance, eps_out=None):
    _lib.check(_lib.load().lm2a_cfg_posterior(
        _stream(), _ptr(x), _ptr(eps), _ptr(noise), _ptr(sched), _ptr(t_dev), t_dev.numel(),
        _ptr(ticket), batch, elems_per_clip, float(guidance), 1 if guided else 0,
        1 if advance else 0, _ptr(eps_out)), "lm2a_cfg_posterior")


def cfg_step(x, eps, noise, clip_seed, sched, t_dev, ticket, batch, c, t, guidance, guided,
             advance, slab=None, copies=1, tp=0, ld=0, zero=None, eps_out=None):
    """The whole per-step update in one launch (see lm2a_cfg_step). zero: StatsArena / tensor."""
    zt = zero.used() if isinstance(zero, StatsArena) else zero
    zbytes = 0 if zt is None else zt.numel() * zt.element_size()
    _lib.check(_lib.load().lm2a_cfg_step(
        _stream(), _ptr(x), _ptr(eps), _ptr(noise), _ptr(clip_seed), _ptr(sched), _ptr(t_dev),
        t_dev.numel(), _ptr(ticket), batch, c, t, float(guidance), 1 if guided else 0,
        1 if advance else 0, _ptr(slab), copies, tp, ld, _ptr(zt), zbytes, _ptr(eps_out)),
        "lm2a_cfg_step")


def philox_normal(out, clip_seed, step):
    """out fp32 [B, c, t] <- the per-clip Philox normals of counter word `step`."""
    b, c, t = out.shape
    _lib.check(_lib.load().lm2a_philox_normal(_stream(), _ptr(out), _ptr(clip_seed), b, c, t,
                                              int(step)), "lm2a_philox_normal")


def cfg_ddim(x, eps, noise, table, t_seq, step_idx, t_dev, ticket, batch, elems_per_clip,
             guidance, guided, advance, x0_out=None):
    _lib.check(_lib.load().lm2a_cfg_ddim(
        _stream(), _ptr(x), _ptr(eps), _ptr(noise), _ptr(table), _ptr(t_seq), _ptr(step_idx),
        _ptr(t_dev), t_dev.numel() if t_dev is not None else 0, _ptr(ticket), batch,
        elems_per_clip, float(guidance), 1 if guided else 0, 1 if advance else 0, _ptr(x0_out)),
        "lm2a_cfg_ddim")


def mel_metrics(gen, real, out, batch, n_mels, t, gen_scale=1.0, gen_shift=0.0):
    _lib.check(_lib.load().lm2a_mel_metrics(_stream(), _ptr(gen), _ptr(real), _ptr(out), batch,
                                            n_mels, t, float(gen_scale), float(gen_shift)),
               "lm2a_mel_metrics")


def bias_add(x, x_ld, x_off, y, y_ld, y_off, bias, slots, tp, t_valid, c, stats=None):
    _lib.check(_lib.load().lm2a_bias_add_bf16(
        _stream(), _ptr(x, x_off), x_ld, _ptr(y, y_off), y_ld, _ptr(bias), slots, tp, t_valid, c,
        ctypes.c_void_p(stats.ptr()) if stats is not None else None,
        stats.groups if stats is not None else 0, stats.cg if stats is not None else 0,
        stats.chan0 if stats is not None else 0), "lm2a_bias_add_bf16")


def gn_apply(x, x_ld, y, y_ld, stats, gamma, beta, rows, tp, t_valid, c, groups, eps=1e-5,
             silu=True, x_chan_off=0, y_chan_off=0):
    assert stats.cg * groups == c and stats.chan0 == 0
    _lib.check(_lib.load().lm2a_gn_apply_bf16(
        _stream(), _ptr(x, x_chan_off), x_ld, _ptr(y, y_chan_off), y_ld,
        ctypes.c_void_p(stats.ptr()), stats.groups, _ptr(gamma), _ptr(beta),
        rows, tp, t_valid, c, groups, eps, 1 if silu else 0), "lm2a_gn_apply_bf16")
