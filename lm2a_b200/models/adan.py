"""Adan optimizer — drop-in for reference models/adan.py (same constructor, param_groups and
per-parameter state: step, prev_grad, m, v, n), with `step()` executed as ONE fused sm_100a
launch over every parameter tensor (lm2a_adan_step) instead of ~20 elementwise launches per
tensor, and the EMA shadow-weight update of train.py:177-180 optionally folded into the same
pass (`step(ema=(ema_params, decay))`). First piece of SURVEY.md §8 f4 (training step); the
backward kernels are not built, so gradients still come from whoever calls `backward()`.

fp32 parameters on an sm_100a device only; `restart_cond` (a host callback on the state,
unused by the reference's train.py) is not supported on the fused path.
"""
import ctypes

import torch
from torch.optim import Optimizer

from .. import _lib, ops

CHUNK = 65536


class _AdanTensor(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("prev_g", ctypes.c_void_p),
                ("m", ctypes.c_void_p), ("v", ctypes.c_void_p), ("n", ctypes.c_void_p),
                ("ema", ctypes.c_void_p), ("numel", ctypes.c_int64)]


class Adan(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.02, 0.08, 0.01), eps=1e-8, weight_decay=0,
                 restart_cond: callable = None):
        assert len(betas) == 3
        if restart_cond is not None:
            raise RuntimeError("lm2a_b200 Adan: restart_cond is not supported on the fused path")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                        restart_cond=restart_cond)
        super().__init__(params, defaults)
        self._tables = {}

    def _chunk_tables(self, numels, dev):
        key = (tuple(numels), str(dev))
        if key not in self._tables:
            ct, ci = [], []
            for t, n in enumerate(numels):
                for c in range((n + CHUNK - 1) // CHUNK):
                    ct.append(t)
                    ci.append(c)
            self._tables[key] = (torch.tensor(ct, dtype=torch.int32, device=dev),
                                 torch.tensor(ci, dtype=torch.int32, device=dev))
        return self._tables[key]

    @torch.no_grad()
    def step(self, closure=None, ema=None):
        """One optimizer step (reference adan.py:34-114). ema = (iterable of shadow parameters
        in the order of this optimizer's parameters, decay) also applies
        shadow = shadow * decay + p * (1 - decay) (train.py:177-180) in the same launch."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ema_params, ema_decay = (list(ema[0]), float(ema[1])) if ema is not None else (None, 0.0)
        k = 0
        for group in self.param_groups:
            lr, eps, wd = group["lr"], group["eps"], group["weight_decay"]
            beta1, beta2, beta3 = group["betas"]
            # parameters of a group that share a step count go into one launch
            by_step = {}
            for p in group["params"]:
                shadow = ema_params[k] if ema_params is not None else None
                k += 1
                if p.grad is None:
                    continue
                ops.require_device(p)
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("lm2a_b200 Adan: dense fp32 parameters and gradients only")
                if not (p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("lm2a_b200 Adan: parameters / gradients must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    for name in ("prev_grad", "m", "v", "n"):
                        st[name] = torch.zeros_like(p.grad)
                by_step.setdefault(st["step"], []).append((p, st, shadow))
            for step0, items in by_step.items():
                step = step0 + 1
                cm, cv, cn = (1 / (1 - (1 - b) ** step) for b in (beta1, beta2, beta3))
                scal = (ctypes.c_float * 14)(1 - beta1, beta1, 1 - beta2, beta2, 1 - beta3, beta3,
                                             cm, cv, cn, lr, eps, 1 + wd * lr, ema_decay,
                                             1.0 - ema_decay)
                dev = items[0][0].device
                host = (_AdanTensor * len(items))()
                for i, (p, st, shadow) in enumerate(items):
                    host[i] = _AdanTensor(p.data_ptr(), p.grad.data_ptr(),
                                          st["prev_grad"].data_ptr(), st["m"].data_ptr(),
                                          st["v"].data_ptr(), st["n"].data_ptr(),
                                          shadow.data_ptr() if shadow is not None else None,
                                          p.numel())
                raw = torch.frombuffer(bytearray(bytes(host)), dtype=torch.uint8).to(dev)
                ct, ci = self._chunk_tables([p.numel() for p, _, _ in items], dev)
                _lib.check(_lib.load().lm2a_adan_step(
                    ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream),
                    ctypes.c_void_p(raw.data_ptr()), ctypes.c_void_p(ct.data_ptr()),
                    ctypes.c_void_p(ci.data_ptr()), ct.numel(), CHUNK, 1 if step0 == 0 else 0,
                    scal), "lm2a_adan_step")
                for p, st, shadow in items:
                    st["step"] = step
                    # the kernel wrote the parameter (and its EMA shadow) through raw pointers:
                    # bump the version counters so that caches keyed on them (packed GEMM
                    # operands of UNet1D_ultimate / CondProjection) see the update
                    torch._C._increment_version(p)
                    if shadow is not None:
                        torch._C._increment_version(shadow)
        return loss
