"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (imported from
/root/reference, which exists only in the build container) on seeded inputs.

    python oracle/make_golden.py            # rewrites tests/golden/

Weights are never stored: both sides rebuild them from oracle.random_state_dict(cfg, seed).
The fixtures hold inputs + the reference's outputs only. tests/test_oracle_golden.py pins
oracle/lm2a_oracle.py against them; the GPU tests then compare the CUDA path to the oracle.
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("LM2A_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import lm2a_oracle as orc  # noqa: E402


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"{REF} not found: golden vectors can only be regenerated where the "
                         "reference is mounted")
    # matplotlib is only used for PNG side outputs (reference sample.py:11,258-276)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            sys.modules[name] = m
    plt = sys.modules["matplotlib.pyplot"]
    for fn in ("figure", "imshow", "colorbar", "title", "savefig", "close"):
        setattr(plt, fn, lambda *a, **k: None)
    sys.modules["matplotlib"].pyplot = plt
    sys.path.insert(0, REF)
    import sample as ref_sample  # noqa: E402
    from models.diffusion import GaussianDiffusion  # noqa: E402
    from models.embedding import CondProjection  # noqa: E402
    from models.unet1d_ultimate import UNet1D_ultimate  # noqa: E402
    return ref_sample, UNet1D_ultimate, CondProjection, GaussianDiffusion


def ref_unet(UNet, cfg, seed):
    net = UNet(in_dim=cfg.in_dim, base_dim=cfg.base_dim, dim_mults=cfg.dim_mults,
               cond_dim=cfg.cond_dim, time_emb_dim=cfg.time_emb_dim,
               num_res_blocks=cfg.num_res_blocks, mid_blocks=cfg.mid_blocks,
               attn_heads=cfg.attn_heads)
    sd = orc.random_state_dict(cfg, seed)
    own = net.state_dict()
    assert list(own.keys()) == list(sd.keys()), "state_dict spec drifted from the reference"
    for k in own:
        assert tuple(own[k].shape) == tuple(sd[k].shape), k
    net.load_state_dict(sd, strict=True)
    return net.eval()


def unet_case(UNet, name, cfg, seed, bsz, t_len, lk, timesteps):
    g = torch.Generator().manual_seed(1000 + seed)
    x = torch.randn(bsz, cfg.in_dim, t_len, generator=g)
    t = torch.tensor(timesteps, dtype=torch.long)
    motion_f = torch.randn(bsz, lk, cfg.cond_dim, generator=g)
    text_f = torch.randn(bsz, lk, cfg.cond_dim, generator=g)
    net = ref_unet(UNet, cfg, seed)
    with torch.no_grad():
        eps = net(x, t, motion_f, text_f)
        eps_nocond = net(x, t)
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"), x=x.numpy(), t=t.numpy(), motion_f=motion_f.numpy(),
        text_f=text_f.numpy(), eps=eps.numpy(), eps_nocond=eps_nocond.numpy(), seed=seed,
        cfg=np.array([cfg.in_dim, cfg.base_dim, cfg.cond_dim, cfg.time_emb_dim,
                      cfg.num_res_blocks, cfg.mid_blocks, cfg.attn_heads] + list(cfg.dim_mults)))
    print(name, "eps std", float(eps.std()), "shape", tuple(eps.shape))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_sample, UNet, CondProjection, GaussianDiffusion = import_reference()

    # (1) UNet forward: tiny odd-length case (pad path 9->18 vs skip 19), Lk != T
    unet_case(UNet, "unet_tiny", orc.UNetConfig(80, 16, (1, 2, 4), 32, 32, 2, 3, 4), 3, 2, 77, 23,
              [999, 0])
    # (2) class-default architecture (base 128, 4 heads)
    unet_case(UNet, "unet_default", orc.UNetConfig(), 4, 2, 100, 60, [500, 17])
    # (3) production architecture (sample.py:25-36), short clip
    unet_case(UNet, "unet_production", orc.UNetConfig.production(), 5, 1, 64, 64, [250])
    # (4) smallest architecture the sm_100a path accepts (64-channel granularity)
    unet_case(UNet, "unet_b64", orc.UNetConfig(80, 64, (1, 2, 4), 128, 256, 2, 3, 2), 6, 2, 132,
              132, [700, 3])

    # (5) CondProjection
    cp = CondProjection(234, 768, 128)
    cp_sd = orc.random_cond_proj_state_dict(seed=7)
    cp.load_state_dict(cp_sd, strict=True)
    g = torch.Generator().manual_seed(11)
    motion = torch.randn(2, 50, 234, generator=g)
    lyrics = 0.3 * torch.randn(2, 50, 768, generator=g)
    with torch.no_grad():
        mf, tf = cp(motion, lyrics)
    np.savez_compressed(os.path.join(OUT, "cond_proj.npz"), motion=motion.numpy(),
                        lyrics=lyrics.numpy(), motion_f=mf.numpy(), text_f=tf.numpy())

    # (6) GaussianDiffusion tables + p_sample (diffusion.py:14-18,61-103) on the tiny model
    cfg = orc.UNetConfig(80, 16, (1, 2, 4), 32, 32, 2, 3, 4)
    net = ref_unet(UNet, cfg, 3)
    diff = GaussianDiffusion(net, timesteps=50, device="cpu")
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 80, 40, generator=g)
    mf = torch.randn(2, 40, 32, generator=g)
    tf = torch.randn(2, 40, 32, generator=g)
    outs = {}
    for tt in (49, 7, 0):
        torch.manual_seed(300 + tt)
        outs[f"x_prev_t{tt}"] = diff.p_sample(x, tt, mf, tf).numpy()
        torch.manual_seed(300 + tt)
        outs[f"noise_t{tt}"] = torch.randn_like(x).numpy()
    np.savez_compressed(os.path.join(OUT, "p_sample.npz"), x=x.numpy(), motion_f=mf.numpy(),
                        text_f=tf.numpy(), betas=diff.betas.numpy(), alphas=diff.alphas.numpy(),
                        alpha_bars=diff.alpha_bars.numpy(), **outs)

    # (7) sample.sample_from_npz end to end (production net, 4 steps), guided and unguided
    clip = orc.synthetic_clip(0, t_mel=48, t_motion=20)
    pcfg = orc.UNetConfig.production()
    ck = {"unet": orc.random_state_dict(pcfg, 5), "cond_proj": cp_sd, "timesteps": 4,
          "dataset_mean": -4.5, "dataset_std": 2.0}
    orig_build = ref_sample.build_models

    def seeded_build(*a, **k):  # RNG state after model construction is init-dependent
        out = orig_build(*a, **k)
        torch.manual_seed(42)
        return out

    ref_sample.build_models = seeded_build
    with tempfile.TemporaryDirectory() as tmp:
        npz = os.path.join(tmp, "clip0.npz")
        np.savez(npz, **clip)
        res = {}
        for gw in (1.0, 2.1):
            ck["guidance_weight"] = gw
            ckpt = os.path.join(tmp, "ck.pt")
            torch.save(ck, ckpt)
            out_npz = ref_sample.sample_from_npz(npz, ckpt, os.path.join(tmp, "out"), device="cpu")
            d = np.load(out_npz)
            key = "gw%d" % int(gw * 10)
            res[key + "_mel"] = d["mel"]
            res[key + "_motion_proj"] = d["motion_proj"]
            res[key + "_lyrics_proj"] = d["lyrics_proj"]
            res["motion_rs"] = d["motion"]
            res["lyrics_rs"] = d["lyrics"]
    ref_sample.build_models = orig_build
    np.savez_compressed(os.path.join(OUT, "sample_from_npz.npz"), mel=clip["mel"],
                        motion=clip["motion"], lyrics=clip["lyrics"], **res)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
