"""Drop-in for the reference's sample.py (sample.py:20-311): same functions, signatures,
checkpoint handling, console lines and output npz — with the per-step loop replaced by the
CUDA-Graph-replayed sm_100a path (GaussianDiffusion.sample_cfg), plus a batched entry point
(`sample_clips`) for throughput: the reference is hard-wired to B = 1 (sample.py:135).

    python -m lm2a_b200.sample --npz clip.npz --ckpt ckpt.pt --out_dir out --guidance 2.1
"""
import argparse
import os

import numpy as np
import torch

from .models.diffusion import GaussianDiffusion
from .models.embedding import CondProjection
from .models.unet1d_ultimate import UNet1D_ultimate

# fallback statistics used when the checkpoint carries none (sample.py:47-48)
DATASET_MEAN = -4.63706636428833
DATASET_STD = 1.8648223876953125


def load_checkpoint(path, device="cpu"):
    return torch.load(path, map_location=device)


def build_models(cond_dim=128, base_dim=256, dim_mults=(1, 2, 4), time_emb_dim=256, device="cpu"):
    """Production architecture of the reference (sample.py:25-39)."""
    unet = UNet1D_ultimate(in_dim=80, base_dim=base_dim, dim_mults=tuple(dim_mults),
                           cond_dim=cond_dim, time_emb_dim=time_emb_dim, num_res_blocks=2,
                           mid_blocks=3, attn_heads=8).to(device)
    cond_proj = CondProjection(motion_dim=78 * 3, text_dim=768, out_dim=cond_dim).to(device)
    return unet, cond_proj


def match_len(arr, target_len, mode="repeat"):
    """Host-side time resampling of a condition sequence (reference datasetcode/dataset.py:
    77-106). 'interp': per-feature linear interpolation onto linspace(0, T-1, target_len);
    'repeat': truncate, or pad by repeating the last frame."""
    if arr is None:
        return None
    arr = np.asarray(arr)
    cur = arr.shape[0]
    if cur == target_len:
        return arr.astype(np.float32)
    if mode == "interp":
        # vectorised form of the reference's per-column np.interp loop (same arithmetic:
        # y0 + (x - x0) * (y1 - y0) / (x1 - x0) with unit spacing)
        x_new = np.linspace(0, cur - 1, num=target_len)
        flat = arr.reshape(cur, -1).astype(np.float64)
        out = np.empty((target_len, flat.shape[1]), dtype=np.float32)
        for d in range(flat.shape[1]):
            out[:, d] = np.interp(x_new, np.arange(cur), flat[:, d])
        return out.reshape((target_len,) + arr.shape[1:])
    if cur > target_len:
        return arr[:target_len].astype(np.float32)
    pad = np.repeat(arr[-1:].astype(np.float32), target_len - cur, axis=0)
    return np.concatenate([arr.astype(np.float32), pad], axis=0)


def _mel_length(mel):
    """(80, T) or (T, 80) -> (mel as (80, T), T); sample.py:60-71."""
    if mel.ndim != 2:
        raise RuntimeError("unexpected mel shape: " + str(mel.shape))
    if mel.shape[0] == 80:
        return mel, mel.shape[1]
    if mel.shape[1] == 80:
        return mel.T, mel.shape[0]
    return mel, mel.shape[1]


def load_weights(unet, cond_proj, ck):
    """EMA weights preferred, silent strict=False loads (sample.py:78-102)."""
    if "ema_unet" in ck or "ema_cond_proj" in ck:
        print("found EMA weights in ckpt; loading EMA for sampling")
        if "ema_unet" in ck:
            try:
                unet.load_state_dict(ck["ema_unet"], strict=False)
            except Exception:
                print("failed loading ema_unet, falling back to normal unet")
        else:
            unet.load_state_dict(ck.get("unet", {}), strict=False)
        if "ema_cond_proj" in ck:
            try:
                cond_proj.load_state_dict(ck["ema_cond_proj"], strict=False)
            except Exception:
                print("failed loading ema_cond_proj, falling back to normal cond_proj")
        else:
            cond_proj.load_state_dict(ck.get("cond_proj", {}), strict=False)
    else:
        unet.load_state_dict(ck.get("unet", {}), strict=False)
        cond_proj.load_state_dict(ck.get("cond_proj", {}), strict=False)


def dataset_stats(ck):
    mean, std = DATASET_MEAN, DATASET_STD
    if "dataset_mean" in ck and "dataset_std" in ck:
        try:
            mean, std = float(ck["dataset_mean"]), float(ck["dataset_std"])
            print(f"using dataset mean/std from ckpt: {mean} {std}")
        except Exception:
            print("found dataset_mean/std in ckpt but failed to parse; using fallback constants")
    else:
        print(f"using fallback dataset mean/std: {mean} {std}")
    return mean, std


def _reporter(verbose=True):
    """The reference's periodic stats line + non-finite early stop (sample.py:216-223)."""
    def report(t, x):
        xt = x.detach().float().cpu()
        if torch.isfinite(xt).all():
            if verbose:
                print(f"[sampling] step t={t:4d}  x min={xt.min().item():.6f} "
                      f"max={xt.max().item():.6f} mean={xt.mean().item():.6f} "
                      f"std={xt.std().item():.6f}")
            return True
        print(f"[sampling] step t={t:4d} contains non-finite values; stopping early")
        return False
    return report


@torch.no_grad()
def sample_clips(unet, cond_proj, diffusion, motions, lyrics, t_len, guidance_weight=1.0,
                 x_init=None, noises=None, use_graph=True, report=None, pinned=True):
    """Batched sampling of B clips of equal length.

    motions / lyrics: HOST float32 arrays (B, T, 234) / (B, T, 768), already resampled to
    T = t_len (match_len). Returns (mel_norm (B, 80, T) HOST float32, motion_f, text_f device
    tensors). Host->device of the conditions and device->host of the mels are part of this
    call (bench.py times it as the end-to-end number)."""
    dev = next(unet.parameters()).device
    m = torch.from_numpy(np.ascontiguousarray(motions, dtype=np.float32))
    ly = torch.from_numpy(np.ascontiguousarray(lyrics, dtype=np.float32))
    if pinned:
        m, ly = m.pin_memory(), ly.pin_memory()
    m, ly = m.to(dev, non_blocking=True), ly.to(dev, non_blocking=True)
    motion_f, text_f = cond_proj(m, ly)
    bsz = m.shape[0]
    x = diffusion.sample_cfg((bsz, 80, t_len), motion_f, text_f, guidance_weight, x_init, noises,
                             use_graph, report)
    out = torch.empty(x.shape, dtype=torch.float32, pin_memory=pinned)
    out.copy_(x, non_blocking=False)
    return out.numpy(), motion_f, text_f


_PINNED = {}   # (tag, shape) -> pinned staging buffer, reused across batches


def _staging(tag, shape, pinned):
    """Host staging buffer; pinned ones are cached (cudaHostAlloc of tens of MB per batch would
    cost more than the copy it speeds up)."""
    if not pinned:
        return torch.empty(shape, dtype=torch.float32)
    key = (tag, tuple(shape))
    if key not in _PINNED:
        _PINNED[key] = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    return _PINNED[key]


def _pad_batch(seqs, pinned, tag="cond"):
    """List of (L_i, D) host arrays -> (pinned) fp32 (B, L_max, D) tensor + int32 lengths. The
    rows past a clip's length are never read by the resampling kernel."""
    lens = [int(a.shape[0]) for a in seqs]
    d = int(seqs[0].shape[1])
    buf = _staging(tag, (len(seqs), max(lens), d), pinned)
    for i, a in enumerate(seqs):
        buf[i, : lens[i]] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return buf, torch.tensor(lens, dtype=torch.int32)


@torch.no_grad()
def sample_clips_raw(unet, cond_proj, diffusion, motions, lyrics, t_len, guidance_weight=1.0,
                     x_init=None, noises=None, use_graph=True, report=None, pinned=True,
                     want_resampled=False):
    """Batched sampling straight from the npz-shaped RAW conditions: `motions` /
    `lyrics` are lists of HOST arrays (L_i, 234) / (L_i, 768) of any per-clip length. The
    resampling to T = t_len (match_len 'interp'), CondProjection and the K/V cache build all
    run on the GPU as the per-batch prologue; nothing but the padded raw sequences crosses
    PCIe. Returns (mel_norm (B, 80, T) host fp32, extras) where extras holds the projected
    conditions and, if `want_resampled`, the resampled fp32 sequences (device tensors)."""
    dev = next(unet.parameters()).device
    m, m_lens = _pad_batch(motions, pinned, "motion")
    ly, l_lens = _pad_batch(lyrics, pinned, "lyrics")
    m, ly = m.to(dev, non_blocking=True), ly.to(dev, non_blocking=True)
    m_lens, l_lens = m_lens.to(dev, non_blocking=True), l_lens.to(dev, non_blocking=True)
    bsz = m.shape[0]
    s = diffusion.sampler(bsz, t_len, t_len, guidance_weight > 1.0)
    dst_m, dst_t = s.cond_slabs()
    mf, tf, m_rs, l_rs = cond_proj.project_raw(m, m_lens, ly, l_lens, t_len, dst_m, dst_t,
                                               want_resampled)
    x = s.run(None, None, guidance_weight, x_init, noises, use_graph, report)
    out = _staging("mel", tuple(x.shape), pinned)
    out.copy_(x, non_blocking=False)   # synchronous: also orders the staging buffers' reuse
    out = out.clone()                  # the caller owns its result; the staging buffer is reused
    extras = {"motion_f": mf.view(bsz, t_len, -1), "text_f": tf.view(bsz, t_len, -1),
              "motion_rs": m_rs, "lyrics_rs": l_rs}
    return out.numpy(), extras


def sample_from_npz(npz_path, ckpt_path, out_dir, device="cuda", timesteps=1000,
                    guidance_weight=1.0):
    """Same contract as reference sample.sample_from_npz (sample.py:42-278): returns the path of
    `<out_dir>/<base>_gen.npz` holding mel (80, T) de-normalised, motion, lyrics, motion_proj,
    lyrics_proj, sr, hop_length. `device` must be a CUDA (sm_100a) device: there is no CPU
    path."""
    os.makedirs(out_dir, exist_ok=True)
    data = np.load(npz_path, allow_pickle=True)
    realmel = data["mel"]
    motion, lyrics = data["motion"], data["lyrics"]
    sr = int(data["sr"]) if "sr" in data else 22050
    hop = int(data["hop_length"]) if "hop_length" in data else 256
    _, t_len = _mel_length(realmel)

    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("lm2a_b200.sample runs on a CUDA sm_100a device only (no CPU "
                           f"fallback); got device={device}")
    unet, cond_proj = build_models(device=device)
    ck = load_checkpoint(ckpt_path, device=device)
    load_weights(unet, cond_proj, ck)
    mean, std = dataset_stats(ck)
    unet.eval()
    cond_proj.eval()

    ck_steps = ck.get("timesteps", None)
    timesteps = int(ck_steps) if ck_steps is not None else timesteps
    diffusion = GaussianDiffusion(unet, timesteps=timesteps, device=device, dataset_mean=mean,
                                  dataset_std=std)
    guidance_weight = float(ck.get("guidance_weight", guidance_weight))

    # match_len(..., 'interp') (sample.py:124-125), CondProjection (:132) and the K/V cache
    # build run on the GPU as the clip's prologue; bit-identical to the host resampling
    mel_norm, ex = sample_clips_raw(
        unet, cond_proj, diffusion, [np.asarray(motion)], [np.asarray(lyrics)], t_len,
        guidance_weight, report=_reporter(), want_resampled=True)
    motion_rs, lyrics_rs = ex["motion_rs"][0].cpu().numpy(), ex["lyrics_rs"][0].cpu().numpy()
    motion_f, text_f = ex["motion_f"].float(), ex["text_f"].float()
    out = mel_norm[0] * std + mean  # de-normalise (sample.py:230)

    base = os.path.splitext(os.path.basename(npz_path))[0]
    out_npz = os.path.join(out_dir, base + "_gen.npz")
    np.savez_compressed(out_npz, mel=out, motion=motion_rs, lyrics=lyrics_rs,
                        motion_proj=motion_f.cpu().numpy(), lyrics_proj=text_f.cpu().numpy(),
                        sr=sr, hop_length=hop)
    print("wrote", out_npz)
    _save_pngs(out_dir, base, out, realmel)
    return out_npz


def _save_pngs(out_dir, base, gen, real):
    """Quick-look PNGs (sample.py:258-276); skipped when matplotlib is not installed."""
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        return
    for suffix, img, title in (("_gen.png", gen, "Generated mel"), ("_real.png", real, "Real mel")):
        png = os.path.join(out_dir, base + suffix)
        plt.figure(figsize=(8, 4))
        plt.imshow(img, aspect="auto", origin="lower")
        plt.colorbar()
        plt.title(title)
        plt.savefig(png)
        plt.close()
        print("wrote", png)


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--npz", default=None, help="single input npz path (overrides --index)")
    p.add_argument("--index", type=int, default=0, help="index into npz dir")
    p.add_argument("--npz_dir", default="npz_split/test")
    p.add_argument("--ckpt", default="checkpoints_adan/ckpt_step_10000.pt")
    p.add_argument("--out_dir", default="samples")
    p.add_argument("--device", default="cuda")
    p.add_argument("--guidance", type=float, default=1.0,
                   help="Classifier-free guidance weight. Default: 1.0 (no guidance)")
    p.add_argument("--steps", type=int, default=1000, help="Number of sampling steps. Default: 1000")
    return p.parse_args()


if __name__ == "__main__":
    args = parse_args()
    if args.npz:
        npz_path = args.npz
    else:
        files = sorted(f for f in os.listdir(args.npz_dir) if f.endswith(".npz"))
        if len(files) == 0:
            raise RuntimeError("no npz in " + args.npz_dir)
        npz_path = os.path.join(args.npz_dir, files[args.index % len(files)])
    print("sampling", npz_path, "->", args.out_dir)
    sample_from_npz(npz_path, args.ckpt, args.out_dir, device=args.device, timesteps=args.steps,
                    guidance_weight=args.guidance)
