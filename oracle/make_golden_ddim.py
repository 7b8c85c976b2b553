"""Generates tests/golden/ddim.npz by calling the UNMODIFIED reference
GaussianDiffusion.ddim_sample (/root/reference/models/diffusion.py:124-165; build container
only) on seeded inputs. tests/test_oracle_golden.py pins oracle.ddim_step against it.

    python oracle/make_golden_ddim.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LM2A_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
CASES = [(49, 40, 0.0), (49, 40, 0.7), (25, 24, 1.0), (10, 0, 0.5), (0, -1, 0.0), (0, -1, 1.0)]


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found")
    sys.path.insert(0, REF)
    from models.diffusion import GaussianDiffusion  # noqa: E402
    diff = GaussianDiffusion(None, timesteps=50, device="cpu")
    g = torch.Generator().manual_seed(77)
    x = torch.randn(2, 80, 40, generator=g)
    eps = torch.randn(2, 80, 40, generator=g)
    out = {"x": x.numpy(), "eps": eps.numpy(), "cases": np.array(CASES, dtype=np.float64)}
    for i, (t, tp, eta) in enumerate(CASES):
        torch.manual_seed(500 + i)
        xp, x0 = diff.ddim_sample(x, t, tp, eps, eta=eta)
        torch.manual_seed(500 + i)
        out[f"noise_{i}"] = torch.randn_like(x).numpy()   # the draw ddim_sample made (t_prev > 0)
        out[f"x_prev_{i}"], out[f"x0_{i}"] = xp.numpy(), x0.numpy()
    np.savez_compressed(os.path.join(OUT, "ddim.npz"), **out)
    print("ddim.npz", os.path.getsize(os.path.join(OUT, "ddim.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
