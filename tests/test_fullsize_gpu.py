"""Parity at the real sizes (BASELINE.json configs 2 and 3, the reference's 1000-step schedule).

The checker is the fp32 oracle (oracle/lm2a_oracle.py: a restatement of sample.py:144-210 and the
model files, pinned to reference outputs by tests/test_oracle_golden.py) executed as PyTorch eager
on the GPU with TF32 switched off, itself pinned here against its own CPU execution on one clip.
Stated tolerances (BASELINE.json north_star): single-step eps within 2e-2 relative (bf16 path);
full 1000-step trajectory with identical injected noise: relative MSE (MSE / variance of the
reference mel) < 2e-2 and mean frame cosine (val.py:81-87) > 0.99 per clip."""
import numpy as np
import pytest
import torch

import lm2a_oracle as orc

pytestmark = pytest.mark.gpu
GW = 2.1
T_MEL = 516


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


_CACHE = {}


def _production():
    if "net" not in _CACHE:
        from lm2a_b200.models import UNet1D_ultimate
        cfg = orc.UNetConfig.production()
        sd = orc.random_state_dict(cfg, 5)
        net = UNet1D_ultimate(80, cfg.base_dim, cfg.dim_mults, cfg.cond_dim, cfg.time_emb_dim,
                              cfg.num_res_blocks, cfg.mid_blocks, cfg.attn_heads)
        net.load_state_dict(sd)
        _CACHE["net"] = (cfg, sd, {k: v.cuda() for k, v in sd.items()}, net.cuda().eval())
    return _CACHE["net"]


def _conditions(batch, time_varying):
    """Projected conditions of `batch` synthetic npz-shaped clips (SURVEY 8d recipe): host
    match_len + the oracle's CondProjection in fp32."""
    cp = orc.random_cond_proj_state_dict(seed=7)
    motions, lyrics = [], []
    for i in range(batch):
        clip = orc.synthetic_clip(i, t_mel=T_MEL, time_varying_lyrics=time_varying)
        motions.append(orc.match_len_interp(clip["motion"], T_MEL))
        lyrics.append(orc.match_len_interp(clip["lyrics"], T_MEL))
    with torch.no_grad():
        return orc.cond_projection(cp, torch.from_numpy(np.stack(motions)),
                                   torch.from_numpy(np.stack(lyrics)))


def _oracle_eps_cat(sd_gpu, cfg, x, t, mf, tf, chunk=16):
    """[uncond rows | cond rows] eps of the doubled batch (sample.py:155-165), fp32 eager on the
    GPU, evaluated in chunks of clips (rows are independent)."""
    outs_u, outs_c = [], []
    with torch.no_grad():
        for i in range(0, x.shape[0], chunk):
            xs, ms, ts = x[i:i + chunk], mf[i:i + chunk], tf[i:i + chunk]
            tb = torch.full((2 * xs.shape[0],), t, device=x.device, dtype=torch.long)
            e = orc.model_forward(sd_gpu, cfg, torch.cat([xs, xs]), tb,
                                  torch.cat([ms * 0, ms]), torch.cat([ts * 0, ts]))
            outs_u.append(e[: xs.shape[0]])
            outs_c.append(e[xs.shape[0]:])
    return torch.cat(outs_u + outs_c)


def test_gpu_eager_oracle_is_pinned_to_cpu_oracle():
    """The checker of this file: fp32 eager on the GPU (TF32 off) == the CPU oracle."""
    _need_gpu()
    cfg, sd, sd_gpu, _ = _production()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 80, T_MEL, generator=g)
    mf, tf = _conditions(1, True)
    with torch.no_grad():
        cpu = orc.cfg_step_eps(sd, cfg, x, 500, mf, tf, GW)
        gpu = orc.cfg_step_eps(sd_gpu, cfg, x.cuda(), 500, mf.cuda(), tf.cuda(), GW)
    assert _rel(gpu, cpu) < 2e-5


@pytest.mark.parametrize("batch,time_varying", [(32, True), (32, False), (64, True)])
def test_guided_step_at_baseline_configs(batch, time_varying, record):
    """BASELINE config 2 exactly (B = 32 -> R = 64 rows, T = Lk = 516, t = 500, guidance 2.1,
    time-varying lyrics; and the tiled-lyrics variant that takes the one-stream launch list) and
    config 3's batch (B = 64 -> R = 128): eps of every row and the blended eps vs the oracle."""
    _need_gpu()
    from lm2a_b200 import ops
    from lm2a_b200.models import GaussianDiffusion
    cfg, sd, sd_gpu, net = _production()
    g = torch.Generator().manual_seed(100 + batch)
    x = torch.randn(batch, 80, T_MEL, generator=g).cuda()
    mf, tf = (c.cuda() for c in _conditions(batch, time_varying))
    diff = GaussianDiffusion(net, timesteps=1000, device="cuda")
    s = diff.sampler(batch, T_MEL, T_MEL, guided=True)
    s.gw = GW
    s.set_conditions(mf, tf)
    assert s.plan.const_text == (not time_varying)
    s.plan.x_in.copy_(x)
    s.plan.t_in.fill_(500)
    s.plan.run()
    eps_cat = s.plan.eps.clone()
    blended = torch.empty_like(x)
    xx = x.clone()
    ops.cfg_posterior(xx, eps_cat, torch.zeros_like(x), diff.sched, s.plan.t_in.clone(), None,
                      batch, x[0].numel(), GW, True, False, blended)
    torch.cuda.synchronize()
    ref_cat = _oracle_eps_cat(sd_gpu, cfg, x, 500, mf, tf)
    ref_blend = orc.cfg_eps(ref_cat[:batch], ref_cat[batch:], GW)
    eu, ec, eb = (_rel(eps_cat[:batch], ref_cat[:batch]), _rel(eps_cat[batch:], ref_cat[batch:]),
                  _rel(blended, ref_blend))
    worst = max(_rel(eps_cat[batch + b], ref_cat[batch + b]) for b in range(batch))
    record("guided_step", batch=batch, time_varying=time_varying, uncond=eu, cond=ec, blended=eb,
           worst_clip=worst)
    assert torch.isfinite(eps_cat).all()
    assert eu < 2e-2 and ec < 2e-2 and worst < 2e-2, (eu, ec, worst)
    assert eb < 3e-2, eb     # the blend amplifies the cond/uncond difference by the guidance


class _SeededNoise:
    """noise of loop iteration i, regenerated on demand (1000 x B x 80 x 516 floats would not be
    worth keeping): the same draw for the oracle loop and the B200 sampler."""

    def __init__(self, shape, seed):
        self.shape, self.seed = shape, seed
        self.gen = torch.Generator(device="cuda")

    def __getitem__(self, i):
        self.gen.manual_seed(self.seed + int(i))
        return torch.randn(self.shape, generator=self.gen, device="cuda")

    __call__ = __getitem__


def test_full_1000_step_trajectory_vs_fp32_oracle(record):
    """The reference's full schedule (sample.py:292: 1000 steps) under CFG at T = 516 on the
    production network, B = 4 clips, identical x_T and per-step noise on both sides."""
    _need_gpu()
    from lm2a_b200.models import GaussianDiffusion
    cfg, sd, sd_gpu, net = _production()
    batch, steps = 4, 1000
    mf, tf = (c.cuda() for c in _conditions(batch, True))
    g = torch.Generator(device="cuda").manual_seed(2024)
    x_t = torch.randn(batch, 80, T_MEL, generator=g, device="cuda")
    noises = _SeededNoise((batch, 80, T_MEL), 7000)
    diff = GaussianDiffusion(net, timesteps=steps, device="cuda")
    got = diff.sample_cfg((batch, 80, T_MEL), mf, tf, GW, x_init=x_t, noises=noises)
    with torch.no_grad():
        ref = orc.sample_loop(sd_gpu, cfg, mf, tf, (batch, 80, T_MEL), steps, GW, x_t, noises)
    assert torch.isfinite(got).all() and torch.isfinite(ref).all()
    res = []
    for b in range(batch):
        mse, cos = orc.mel_metrics(got[b].cpu().numpy(), ref[b].cpu().numpy())
        res.append((mse / float(ref[b].var()), cos))
    record("trajectory_1000", rel_mse=[r[0] for r in res], cos=[r[1] for r in res])
    for b, (rm, cos) in enumerate(res):
        assert rm < 2e-2, f"clip {b}: rel MSE {rm:.3e}"
        assert cos > 0.99, f"clip {b}: frame cosine {cos:.6f}"


def test_sharded_layout_equals_single_device_trajectory():
    """Clip sharding: the two half-batches two GPUs would sample give, bit for bit,
    the rows of the single-device batch (same per-clip x_T and noise), over a 12-step guided
    trajectory on the production network."""
    _need_gpu()
    from lm2a_b200.models import GaussianDiffusion
    cfg, sd, sd_gpu, net = _production()
    batch, steps = 6, 12
    mf, tf = (c.cuda() for c in _conditions(batch, True))
    g = torch.Generator(device="cuda").manual_seed(11)
    x_t = torch.randn(batch, 80, T_MEL, generator=g, device="cuda")
    noises = torch.randn(steps - 1, batch, 80, T_MEL, generator=g, device="cuda")
    diff = GaussianDiffusion(net, timesteps=steps, device="cuda")
    full = diff.sample_cfg((batch, 80, T_MEL), mf, tf, GW, x_init=x_t, noises=noises)
    parts = []
    for lo, hi in ((0, 2), (2, 6)):       # uneven shards: different tile shapes per shard
        parts.append(diff.sample_cfg((hi - lo, 80, T_MEL), mf[lo:hi].contiguous(),
                                     tf[lo:hi].contiguous(), GW, x_init=x_t[lo:hi].contiguous(),
                                     noises=noises[:, lo:hi].contiguous()))
    assert torch.equal(torch.cat(parts), full)


def test_nan_eps_reaches_the_non_finite_guard():
    """torch.clamp propagates NaN (sample.py:170,174): a diverged network must produce a
    non-finite x so that the periodic guard of the loop (sample.py:216-223) can stop it."""
    _need_gpu()
    from lm2a_b200 import ops
    b, n = 2, 80 * 64
    x = torch.randn(b, 80, 64, device="cuda")
    eps = torch.randn(2 * b, 80, 64, device="cuda")
    eps[b + 1, 3, 5] = float("nan")
    sched = torch.rand(10, 4, device="cuda")
    t = torch.full((2 * b,), 4, dtype=torch.int64, device="cuda")
    out = torch.empty_like(x)
    ops.cfg_posterior(x, eps, torch.zeros_like(x), sched, t, None, b, n, GW, True, False, out)
    torch.cuda.synchronize()
    assert torch.isnan(out[1, 3, 5]) and torch.isnan(x[1, 3, 5])
    assert int(torch.isnan(x).sum()) == 1
