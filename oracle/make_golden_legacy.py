"""Generates tests/golden/legacy_*.npz by executing the UNMODIFIED reference legacy model
(/root/reference/models/unet1d.py, build container only) on seeded inputs.

    python oracle/make_golden_legacy.py

Weights are rebuilt on both sides from oracle.legacy_random_state_dict(cfg, seed); the
fixtures hold inputs + the reference's outputs. tests/test_oracle_golden.py pins
oracle.legacy_unet_forward against them."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LM2A_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import lm2a_oracle as orc  # noqa: E402


def case(UNet1D, name, cfg, seed, bsz, t_len, lk, timesteps):
    net = UNet1D(in_dim=cfg.in_dim, base_dim=cfg.base_dim, dim_mults=cfg.dim_mults,
                 cond_dim=cfg.cond_dim, time_emb_dim=cfg.time_emb_dim)
    sd = orc.legacy_random_state_dict(cfg, seed)
    own = net.state_dict()
    assert list(own.keys()) == list(sd.keys()), "legacy state_dict spec drifted from the reference"
    for k in own:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    net.load_state_dict(sd, strict=True)
    net.eval()
    g = torch.Generator().manual_seed(2000 + seed)
    x = torch.randn(bsz, cfg.in_dim, t_len, generator=g)
    t = torch.tensor(timesteps, dtype=torch.long)
    mf = torch.randn(bsz, lk, cfg.cond_dim, generator=g)
    tf = torch.randn(bsz, lk, cfg.cond_dim, generator=g)
    with torch.no_grad():
        eps = net(x, t, mf, tf)
        # CFG uncond rows: zeroed projected conditions (sample.py:155-157)
        eps_zero = net(x, t, mf * 0, tf * 0)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), x=x.numpy(), t=t.numpy(), motion_f=mf.numpy(),
        text_f=tf.numpy(), eps=eps.numpy(), eps_zero=eps_zero.numpy(), seed=seed,
        cfg=np.array([cfg.in_dim, cfg.base_dim, cfg.cond_dim, cfg.time_emb_dim]
                     + list(cfg.dim_mults)))
    print(name, "eps std", float(eps.std()), tuple(eps.shape))


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found")
    sys.path.insert(0, REF)
    from models.unet1d import UNet1D  # noqa: E402
    torch.set_num_threads(os.cpu_count() or 1)
    # base 128: head dims 32, 32, 64, 128, 192, 96, 64; odd length (pad path 2*T/8 -> T/4)
    case(UNet1D, "legacy_b128", orc.LegacyConfig(80, 128, (1, 2, 4), 128, 256), 21, 2, 100, 60,
         [800, 5])
    # the width the reference trains (base 256): head dims 64, 64, 128, 256, 384, 192, 128
    case(UNet1D, "legacy_b256", orc.LegacyConfig(80, 256, (1, 2, 4), 128, 256), 22, 1, 72, 72,
         [321])


if __name__ == "__main__":
    main()
