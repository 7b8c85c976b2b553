"""torch custom-op registration of the C ABI (`torch.ops.lm2a.*`): a thin dispatcher-visible
layer over lm2a_b200/ops.py for callers that want the kernels as torch operators (schemas with
mutated arguments declared, CUDA dispatch key only — there is no CPU implementation to fall
back to). Importing this module registers the ops once per process.

    import lm2a_b200.torch_ops  # noqa: F401
    torch.ops.lm2a.cfg_posterior(x, eps_cat, noise, sched, t_dev, ticket, guidance, True, True)

The UNet itself is not a single op: its launch plan (engine.UNetPlan) is a list of C-ABI calls
captured in a CUDA Graph (descriptors with fused GroupNorm / FiLM / statistics / virtual concat);
the ops registered here are tensor-in / tensor-out forms of the same kernels: the update kernels,
resampling, metrics, GroupNorm + SiLU, upsampling, a plain conv (bias + optional residual), the
three attention cores, the V transpose, the timestep MLP / FiLM tables and the Philox normals.
"""
import torch

from . import ops

_lib = torch.library.Library("lm2a", "DEF")

_lib.define("cfg_posterior(Tensor(a!) x, Tensor eps, Tensor? noise, Tensor sched, Tensor(b!) t_dev, "
            "Tensor(c!)? ticket, float guidance, bool guided, bool advance) -> ()")
_lib.define("cfg_ddim(Tensor(a!) x, Tensor eps, Tensor? noise, Tensor table, Tensor? t_seq, "
            "Tensor(b!) step_idx, Tensor(c!)? t_dev, Tensor(d!)? ticket, float guidance, "
            "bool guided, bool advance) -> ()")
_lib.define("resample_seq(Tensor x, Tensor? lens, int t_out) -> Tensor")
_lib.define("mel_metrics(Tensor gen, Tensor real, float gen_scale, float gen_shift) -> Tensor")
_lib.define("gn_silu(Tensor x, Tensor gamma, Tensor beta, int rows, int tp, int t_valid, "
            "int groups, float eps, bool silu) -> Tensor")
_lib.define("upsample2x(Tensor x, int rows, int tp_in, int t_in, int tp_out) -> Tensor")
_lib.define("conv1d(Tensor x, Tensor w, Tensor bias, int n_valid, int tp, int t_valid, int taps, "
            "Tensor? residual) -> Tensor")
_lib.define("cross_attn(Tensor q, Tensor kv_motion, Tensor kv_text, Tensor kv_slot, int tp, "
            "int t_valid, int lk, int heads, int n_streams) -> Tensor")
_lib.define("cross_attn_cond(Tensor q, Tensor cond_motion, Tensor cond_text, Tensor kv_slot, int tp, "
            "int t_valid, int lk, int heads, int n_streams) -> Tensor")
_lib.define("cross_attn_tail(Tensor(a!) o, Tensor q, Tensor kv_motion, Tensor kv_text, "
            "Tensor kv_slot, int tp, int t0, int n_tail, int lk, int heads, int n_streams, "
            "bool shared_kv) -> ()")
_lib.define("transpose_kv(Tensor kv, int slots, int lk, int e) -> Tensor")
_lib.define("time_mlp(Tensor t, Tensor w, Tensor b) -> Tensor")
_lib.define("film(Tensor s, Tensor w, Tensor b) -> Tensor")
_lib.define("philox_normal(Tensor(a!) out, Tensor clip_seed, int step) -> ()")


def _cfg_posterior(x, eps, noise, sched, t_dev, ticket, guidance, guided, advance):
    ops.require_device(x)
    ops.cfg_posterior(x, eps, noise, sched, t_dev, ticket, x.shape[0], x[0].numel(), guidance,
                      guided, advance)


def _cfg_ddim(x, eps, noise, table, t_seq, step_idx, t_dev, ticket, guidance, guided, advance):
    ops.require_device(x)
    ops.cfg_ddim(x, eps, noise, table, t_seq, step_idx, t_dev, ticket, x.shape[0], x[0].numel(),
                 guidance, guided, advance)


def _resample_seq(x, lens, t_out):
    """fp32 (B, L_max, D) [+ int32 lengths] -> fp32 (B, t_out, D): match_len(..., 'interp')."""
    ops.require_device(x)
    b, lmax, d = x.shape
    out = torch.empty(b, t_out, d, dtype=torch.float32, device=x.device)
    ops.resample_seq(x.contiguous(), lens, out, None, b, lmax, d, t_out, t_out, d)
    return out


def _mel_metrics(gen, real, gen_scale, gen_shift):
    """fp32 (B, n_mels, T) x 2 -> fp64 (B, 8): val.py:25-113, see lm2a_mel_metrics."""
    ops.require_device(gen)
    b, n_mels, t = gen.shape
    out = torch.empty(b, 8, dtype=torch.float64, device=gen.device)
    ops.mel_metrics(gen.contiguous(), real.contiguous(), out, b, n_mels, t, gen_scale, gen_shift)
    return out


def _gn_silu(x, gamma, beta, rows, tp, t_valid, groups, eps, silu):
    """bf16 slab [rows * tp, C] -> SiLU(GroupNorm(x)) as a new slab (pad slots zero)."""
    ops.require_device(x)
    c = x.shape[1]
    y = torch.zeros_like(x)
    ops.gn_silu(x, c, y, c, gamma, beta, rows, tp, t_valid, c, groups, eps, silu)
    return y


def _upsample2x(x, rows, tp_in, t_in, tp_out):
    """bf16 slab [rows * tp_in, C] -> linear x2 (align_corners=True) slab [rows * tp_out, C]."""
    ops.require_device(x)
    c = x.shape[1]
    y = torch.zeros(rows * tp_out, c, dtype=x.dtype, device=x.device)
    ops.upsample2x(x, c, y, c, rows, tp_in, t_in, tp_out, c)
    return y


def _conv1d(x, w, bias, n_valid, tp, t_valid, taps, residual):
    """nn.Conv1d on a bf16 slab [rows * tp, Cin] with packed weights [N_pad, taps * Cin] (K ordered
    tap-major, see engine._conv_w) and fp32 bias [N_pad]; taps: 0 = k1, 1 = k3 p1, 2 = k4 s2 p1
    (x is then the [rows * tp, 2 * Cin] slot-pair view and tp / t_valid those of the OUTPUT).
    residual: optional bf16 slab [rows * tp, n_valid] added in the epilogue. -> bf16 slab."""
    ops.require_device(x)
    m, ld = x.shape
    cin = ld // 2 if taps == ops.TAPS_K4S2 else ld
    out = torch.zeros(m, n_valid, dtype=x.dtype, device=x.device)
    res = {} if residual is None else dict(residual=residual, res_ld=residual.shape[1])
    ops.conv1d(ops.make_conv_desc([ops.Seg(x, ld, cin, taps, m)], w, bias, n_valid, m, tp, t_valid,
                                  out, n_valid, **res))
    return out


def _kv_views(kv, e):
    return ops._ptr(kv), ops._ptr(kv, e)


def _cross_attn(q, kv_motion, kv_text, kv_slot, tp, t_valid, lk, heads, n_streams):
    """softmax(q k^T) v per (clip-row, stream, head). q: bf16 slab [rows * tp, n_streams * E]
    pre-scaled by log2(e)/sqrt(d_h); kv_*: the K | V projection output bf16 [slots * lk, 2E];
    kv_slot int32 [rows]. Builds the V^T operand (lm2a_transpose_kv_bf16) and returns the O slab."""
    ops.require_device(q)
    e = kv_motion.shape[1] // 2
    rows, slots = q.shape[0] // tp, kv_motion.shape[0] // lk
    lk_pad = (lk + 7) // 8 * 8
    vts = []
    for kv in (kv_motion, kv_text):
        vt = torch.zeros(slots * e, lk_pad, dtype=q.dtype, device=q.device)
        ops.transpose_kv(kv, 2 * e, e, vt, lk_pad, slots, lk, e)
        vts.append(vt)
    o = torch.zeros_like(q)
    ops.cross_attn(q, q.shape[1], o, q.shape[1], ops._ptr(kv_motion), ops._ptr(vts[0]),
                   ops._ptr(kv_text), ops._ptr(vts[1]), 2 * e, lk_pad, kv_slot, slots, rows, tp,
                   t_valid, lk, e, heads, n_streams)
    return o


def _cross_attn_cond(q, cond_motion, cond_text, kv_slot, tp, t_valid, lk, heads, n_streams):
    """softmax(q'_h C^T) C per head against the raw condition slabs bf16 [slots * lk, 128]."""
    ops.require_device(q)
    rows, slots = q.shape[0] // tp, cond_motion.shape[0] // lk
    o = torch.zeros_like(q)
    ops.cross_attn_cond(q, q.shape[1], o, q.shape[1], ops._ptr(cond_motion), ops._ptr(cond_text),
                        cond_motion.shape[1], kv_slot, slots, rows, tp, t_valid, lk, heads, n_streams)
    return o


def _cross_attn_tail(o, q, kv_motion, kv_text, kv_slot, tp, t0, n_tail, lk, heads, n_streams,
                     shared_kv):
    """Rows t0 .. t0 + n_tail - 1 of every (clip-row, stream, head) into o (CUDA cores).
    shared_kv: kv_* are the condition slabs [slots * lk, 128]; else the K | V output [.., 2E]."""
    ops.require_device(q)
    rows, slots = q.shape[0] // tp, kv_motion.shape[0] // lk
    ld = kv_motion.shape[1]
    if shared_kv:
        e = heads * ld
        km, vm, kt, vt = (ops._ptr(kv_motion), ops._ptr(kv_motion), ops._ptr(kv_text),
                          ops._ptr(kv_text))
    else:
        e = ld // 2
        (km, vm), (kt, vt) = _kv_views(kv_motion, e), _kv_views(kv_text, e)
    ops.cross_attn_tail(q, q.shape[1], o, o.shape[1], km, vm, kt, vt, ld, ld, kv_slot, slots, rows,
                        tp, t0, n_tail, lk, e, heads, n_streams, shared_kv)


def _transpose_kv(kv, slots, lk, e):
    """V half of the K | V projection output [slots * lk, 2E] -> V^T [slots * E, lk_pad]."""
    ops.require_device(kv)
    lk_pad = (lk + 7) // 8 * 8
    vt = torch.zeros(slots * e, lk_pad, dtype=kv.dtype, device=kv.device)
    ops.transpose_kv(kv, kv.shape[1], e, vt, lk_pad, slots, lk, e)
    return vt


def _time_mlp(t, w, b):
    """SiLU(SiLU(W sinus(t) + b)) (embedding.py:19-43 + the SiLU in front of every FiLM Linear):
    int64 [rows] -> fp32 [rows, dim]."""
    ops.require_device(w)
    out = torch.empty(t.numel(), w.shape[0], dtype=torch.float32, device=w.device)
    ops.time_mlp(t, w, b, out, t.numel(), w.shape[0])
    return out


def _film(s, w, b):
    """All FiLM Linears at once (unet1d_ultimate.py:43-65): fp32 [rows, dim] x [cols, dim]^T + b."""
    ops.require_device(s)
    out = torch.empty(s.shape[0], w.shape[0], dtype=torch.float32, device=s.device)
    ops.film(s, w, b, out, s.shape[0], s.shape[1], w.shape[0])
    return out


def _philox_normal(out, clip_seed, step):
    ops.require_device(out)
    ops.philox_normal(out, clip_seed, step)


for _name, _fn in (("cfg_posterior", _cfg_posterior), ("cfg_ddim", _cfg_ddim),
                   ("resample_seq", _resample_seq), ("mel_metrics", _mel_metrics),
                   ("gn_silu", _gn_silu), ("upsample2x", _upsample2x), ("conv1d", _conv1d),
                   ("cross_attn", _cross_attn), ("cross_attn_cond", _cross_attn_cond),
                   ("cross_attn_tail", _cross_attn_tail), ("transpose_kv", _transpose_kv),
                   ("time_mlp", _time_mlp), ("film", _film), ("philox_normal", _philox_normal)):
    _lib.impl(_name, _fn, "CUDA")
