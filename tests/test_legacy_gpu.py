"""Legacy UNet1D (reference models/unet1d.py; SURVEY §8 a15, BASELINE config 5) on B200:
forward against the reference's own outputs (tests/golden/legacy_*.npz, written by
oracle/make_golden_legacy.py) and a short guided trajectory against the oracle.
Tolerance: 2e-2 relative (bf16), BASELINE.json north_star."""
import os

import numpy as np
import pytest
import torch

import lm2a_oracle as orc

pytestmark = pytest.mark.gpu
TOL_BF16 = 2e-2


def _model(cfg, sd):
    from lm2a_b200.models import UNet1D
    net = UNet1D(in_dim=cfg.in_dim, base_dim=cfg.base_dim, dim_mults=cfg.dim_mults,
                 cond_dim=cfg.cond_dim, time_emb_dim=cfg.time_emb_dim)
    assert list(net.state_dict().keys()) == list(sd.keys())
    net.load_state_dict(sd, strict=True)
    return net.to("cuda").eval()


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name", ["legacy_b128", "legacy_b256"])
def test_legacy_forward_matches_reference_golden(golden_dir, name):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    d = np.load(os.path.join(golden_dir, name + ".npz"))
    a = [int(v) for v in d["cfg"]]
    cfg = orc.LegacyConfig(a[0], a[1], tuple(a[4:]), a[2], a[3])
    net = _model(cfg, orc.legacy_random_state_dict(cfg, int(d["seed"])))
    x = torch.from_numpy(d["x"]).cuda()
    t = torch.from_numpy(d["t"]).cuda()
    mf, tf = torch.from_numpy(d["motion_f"]).cuda(), torch.from_numpy(d["text_f"]).cuda()
    eps = net(x, t, mf, tf)
    torch.cuda.synchronize()
    assert eps.shape == x.shape and eps.dtype == torch.float32
    err = _rel(eps, torch.from_numpy(d["eps"]))
    assert err < TOL_BF16, f"{name}: eps rel-L2 {err:.3e}"
    # zeroed conditions through the FULL attention path (no shortcut in plain forward)
    err0 = _rel(net(x, t, mf * 0, tf * 0), torch.from_numpy(d["eps_zero"]))
    assert err0 < TOL_BF16, f"{name}: zero-condition eps rel-L2 {err0:.3e}"
    assert torch.equal(net(x, t, mf, tf), eps)


def test_legacy_full_length_vs_oracle():
    """The width the reference trains (base 256: head dims 64..384, decoder widths 1536 / 768 /
    512) at the canonical clip length T = Lk = 516 with CFG-style rows (uncond = zeroed
    conditions) against the fp32 oracle: multi-tile attention at every head dim, fused
    GroupNorm epilogues, transposed convs at 64 -> 128 (+pad) / 129 -> 258 / 258 -> 516."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    cfg = orc.LegacyConfig(80, 256, (1, 2, 4), 128, 256)
    sd = orc.legacy_random_state_dict(cfg, 22)
    net = _model(cfg, sd)
    g = torch.Generator().manual_seed(177)
    x1 = torch.randn(1, 80, 516, generator=g)
    mf1 = torch.randn(1, 516, 128, generator=g)
    tf1 = torch.randn(1, 516, 128, generator=g)
    x = torch.cat([x1, x1])
    mf, tf = torch.cat([mf1 * 0, mf1]), torch.cat([tf1 * 0, tf1])
    t = torch.tensor([500, 500])
    with torch.no_grad():
        ref = orc.legacy_unet_forward(sd, cfg, x, t, mf, tf)
    eps = net(x.cuda(), t.cuda(), mf.cuda(), tf.cuda())
    assert torch.isfinite(eps).all()
    err = _rel(eps, ref)
    assert err < TOL_BF16, f"eps rel-L2 {err:.3e}"


@pytest.mark.parametrize("tiled_lyrics", [False, True])
def test_legacy_guided_trajectory_vs_oracle(tiled_lyrics):
    """CFG loop (sample.py:144-210) around the legacy model: [uncond, cond] rows, uncond rows on
    the attention-constant shortcut, posterior update; 6 steps with injected noise. With a lyrics
    condition tiled over time (the reference's real data) the one-stream launch list is used."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lm2a_b200.models import GaussianDiffusion
    cfg = orc.LegacyConfig(80, 128, (1, 2, 4), 128, 256)
    sd = orc.legacy_random_state_dict(cfg, 23)
    net = _model(cfg, sd)
    steps, bsz, t_len, lk, gw = 6, 2, 100, 60, 2.1
    g = torch.Generator().manual_seed(321)
    x0 = torch.randn(bsz, 80, t_len, generator=g)
    mf = torch.randn(bsz, lk, 128, generator=g)
    tf = torch.randn(bsz, lk, 128, generator=g)
    if tiled_lyrics:
        tf = tf[:, :1].expand(bsz, lk, 128).contiguous()
    noises = torch.randn(steps - 1, bsz, 80, t_len, generator=g)
    diff = GaussianDiffusion(net, timesteps=steps, device="cuda")
    got = diff.sample_cfg((bsz, 80, t_len), mf.cuda(), tf.cuda(), gw, x_init=x0.cuda(),
                          noises=noises.cuda())
    assert diff.sampler(bsz, t_len, lk, True).plan.const_text is tiled_lyrics
    with torch.no_grad():
        ref = orc.sample_loop(sd, cfg, mf, tf, (bsz, 80, t_len), steps, gw, x0, list(noises))
    assert torch.isfinite(got).all()
    for b in range(bsz):
        mse, cos = orc.mel_metrics(got[b].cpu().numpy(), ref[b].numpy())
        assert mse / float(ref[b].var()) < 2e-3, f"clip {b}: rel MSE {mse / float(ref[b].var()):.3e}"
        assert cos > 0.999, f"clip {b}: frame cosine {cos:.6f}"
    # graph replay path (noise drawn on device) runs and stays finite
    out = diff.sample_cfg((bsz, 80, t_len), mf.cuda(), tf.cuda(), gw)
    assert torch.isfinite(out).all()
