"""torch custom-op registration of the C ABI (`torch.ops.lm2a.*`): a thin dispatcher-visible
layer over lm2a_b200/ops.py for callers that want the kernels as torch operators (schemas with
mutated arguments declared, CUDA dispatch key only — there is no CPU implementation to fall
back to). Importing this module registers the ops once per process.

    import lm2a_b200.torch_ops  # noqa: F401
    torch.ops.lm2a.cfg_posterior(x, eps_cat, noise, sched, t_dev, ticket, guidance, True, True)

The UNet itself is not a single op: its launch plan (engine.UNetPlan) is a list of C-ABI calls
captured in a CUDA Graph; the ops registered here are the ones with tensor-only signatures.
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


for _name, _fn in (("cfg_posterior", _cfg_posterior), ("cfg_ddim", _cfg_ddim),
                   ("resample_seq", _resample_seq), ("mel_metrics", _mel_metrics),
                   ("gn_silu", _gn_silu), ("upsample2x", _upsample2x)):
    _lib.impl(_name, _fn, "CUDA")
