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


class Stats:
    """Partial GroupNorm statistics of a slab [rows, tp, c]: float2 [rows, c/gran, ns] written by
    the producing kernel's epilogue, read by gn_apply. `view(row0, chan0)` addresses a
    row / channel sub-range (a launch over part of the rows, or one half of a concat slab)."""

    def __init__(self, rows, tp, c, gran, dev, buf=None, row0=0, chan0=0, slice0=0, ns=None):
        self.rows, self.tp, self.c, self.gran = rows, tp, c, gran
        self.sub = c // gran
        # `ns` > tp/32 + 2: room for a second producer of the same (row, channels) whose slices
        # start at `slice0` (the odd-slot launch of a transposed conv)
        self.ns = ns if ns is not None else tp // 32 + 2
        self.buf = buf if buf is not None else torch.zeros(rows * self.sub * self.ns, 2,
                                                           dtype=torch.float32, device=dev)
        self.row0, self.chan0, self.slice0 = row0, chan0, slice0

    def view(self, row0=0, chan0=0, slice0=0):
        return Stats(self.rows, self.tp, self.c, self.gran, None, self.buf, self.row0 + row0,
                     self.chan0 + chan0, self.slice0 + slice0, self.ns)

    def ptr(self):
        off = (self.row0 * self.sub + self.chan0 // self.gran) * self.ns + self.slice0
        return self.buf.data_ptr() + off * 8


class Seg:
    """One K segment of the implicit GEMM: a bf16 slab view (tensor + channel offset)."""

    def __init__(self, slab, ld, cin, taps, rows, chan_off=0):
        self.slab, self.ld, self.cin, self.taps, self.rows, self.chan_off = (
            slab, ld, cin, taps, rows, chan_off)


def make_conv_desc(segs, w, bias, n_valid, m, tp, t_valid, out, out_ld, out_chan_off=0,
                   film=None, film_col=0, film_shift_off=0, film_bcast=False, film_row=0,
                   residual=None, res_ld=0,
                   res_chan_off=0, out_mode=OUT_BF16_SLAB, block_n=0, stats=None, cta_group=0,
                   gn=None):
    """Builds the (reusable) descriptor of one lm2a_conv1d_bf16 launch. Keeps the tensors
    alive by attaching them to the descriptor object."""
    d = ConvDesc()
    for i, s in enumerate(segs):
        d.seg[i].x = s.slab.data_ptr() + s.chan_off * 2
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
        d.residual = residual.data_ptr() + res_chan_off * 2
        d.res_ld = res_ld
    d.out_mode = out_mode
    if out is not None:
        d.out = out.data_ptr() + out_chan_off * out.element_size()
    d.out_ld = out_ld
    d.block_n = block_n
    d.cta_group = cta_group
    if stats is not None:  # Stats view: partial GroupNorm sums of the output
        d.stats = stats.ptr()
        d.stats_sub, d.stats_ns, d.stats_gran = stats.sub, stats.ns, stats.gran
    if gn is not None:  # (gamma, beta, groups, eps, gn_out slab, gn_out_ld, barrier words)
        gamma, beta, groups, eps, gn_out, gn_out_ld, barrier = gn
        d.gn_gamma, d.gn_beta = gamma.data_ptr(), beta.data_ptr()
        d.gn_groups, d.gn_eps = groups, eps
        d.gn_out, d.gn_out_ld = gn_out.data_ptr(), gn_out_ld
        d.gn_barrier = barrier.data_ptr()
    d._keep = (segs, w, bias, film, residual, out, stats, gn)
    return d


def conv_gn_fusable(m, n_pad):
    return bool(_lib.load().lm2a_conv_gn_fusable(m, n_pad))


def conv1d(desc):
    _lib.check(_lib.load().lm2a_conv1d_bf16(_stream(), ctypes.byref(desc)), "lm2a_conv1d_bf16")


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


def ingest_x(x, slab, batch, copies, c, t, tp, ld):
    _lib.check(_lib.load().lm2a_ingest_x(_stream(), _ptr(x), _ptr(slab), batch, copies, c, t, tp,
                                         ld), "lm2a_ingest_x")


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
        stats.sub if stats is not None else 0, stats.ns if stats is not None else 0,
        stats.gran if stats is not None else 0), "lm2a_bias_add_bf16")


def gn_apply(x, x_ld, y, y_ld, stats, gamma, beta, rows, tp, t_valid, c, groups, eps=1e-5,
             silu=True, x_chan_off=0, y_chan_off=0):
    _lib.check(_lib.load().lm2a_gn_apply_bf16(
        _stream(), _ptr(x, x_chan_off), x_ld, _ptr(y, y_chan_off), y_ld,
        ctypes.c_void_p(stats.ptr()), stats.sub, stats.ns, stats.gran, _ptr(gamma), _ptr(beta),
        rows, tp, t_valid, c, groups, eps, 1 if silu else 0), "lm2a_gn_apply_bf16")
