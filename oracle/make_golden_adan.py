"""Generates tests/golden/adan.npz by running the UNMODIFIED reference optimizer
(/root/reference/models/adan.py, build container only) for four steps on three small fp32
tensors with seeded gradients, plus the EMA shadow update of train.py:177-180.

    python oracle/make_golden_adan.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LM2A_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
SHAPES = [(37, 5), (1030,), (8, 16, 3)]
STEPS = 4
HP = dict(lr=2e-4, betas=(0.02, 0.08, 0.01), eps=1e-8, weight_decay=1e-4)
EMA_DECAY = 0.999


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found")
    sys.path.insert(0, REF)
    from models.adan import Adan  # noqa: E402
    g = torch.Generator().manual_seed(99)
    params = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in SHAPES]
    shadow = [p.detach().clone() for p in params]
    opt = Adan(params, **HP)
    out = {f"p0_{i}": p.detach().numpy().copy() for i, p in enumerate(params)}
    for step in range(STEPS):
        for i, p in enumerate(params):
            p.grad = torch.randn(p.shape, generator=g) * (0.1 + step)
            out[f"g{step}_{i}"] = p.grad.numpy().copy()
        opt.step()
        for sp, p in zip(shadow, params):                     # train.py:177-180
            sp.data.mul_(EMA_DECAY).add_(p.data * (1.0 - EMA_DECAY))
        for i, p in enumerate(params):
            out[f"p{step + 1}_{i}"] = p.detach().numpy().copy()
            out[f"ema{step + 1}_{i}"] = shadow[i].numpy().copy()
    for i, p in enumerate(params):
        st = opt.state[p]
        for k in ("m", "v", "n", "prev_grad"):
            out[f"{k}_{i}"] = st[k].numpy().copy()
    np.savez_compressed(os.path.join(OUT, "adan.npz"), **out)
    print("adan.npz", os.path.getsize(os.path.join(OUT, "adan.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
