"""Drop-in for the reference's val.py (val.py:25-319): the same functions and result files,
with (a) the per-clip metrics computed by one device kernel (lm2a_mel_metrics) straight on the
sampler's output and (b) `assess_batch` keeping ONE model / checkpoint / CUDA Graph alive and
sampling clips of equal length together, where the reference rebuilds the network and reloads
the checkpoint for every clip (val.py:197-204 -> sample.py:75-102). SURVEY.md §8 f2.

    python -m lm2a_b200.val --ckpt ckpt.pt --npz_dir npz_split/test --out_dir testval
"""
import argparse
import os
import random
import shutil

import numpy as np
import torch

from . import ops
from . import sample as _sample
from .models.diffusion import GaussianDiffusion

METRIC_KEYS = ("mse", "ssim", "avg_cos_sim", "mean_error", "std_error", "snr")


def _as_80_t(mel):
    """(80, T) or (T, 80) -> (80, T), the orientation fix of val.py:171-173,207-209."""
    mel = np.asarray(mel)
    if mel.ndim == 2 and mel.shape[1] == 80 and mel.shape[0] != 80:
        mel = mel.T
    return mel


def metrics_on_device(gen, real, gen_scale=1.0, gen_shift=0.0):
    """gen / real: fp32 CUDA tensors (B, 80, T) -> list of B metric dicts rounded to 6 decimals
    as val.py:106-113. `gen` is de-normalised inside the kernel (gen * scale + shift)."""
    ops.require_device(gen)
    if gen.shape != real.shape or gen.dim() != 3:
        raise RuntimeError(f"metrics: gen {tuple(gen.shape)} vs real {tuple(real.shape)}")
    b, n_mels, t = gen.shape
    out = torch.empty(b, 8, dtype=torch.float64, device=gen.device)
    ops.mel_metrics(gen.contiguous().float(), real.contiguous().float(), out, b, n_mels, t,
                    gen_scale, gen_shift)
    vals = out.cpu().numpy()
    return [{k: round(float(v), 6) for k, v in zip(METRIC_KEYS, row[:6])} for row in vals]


def compute_metrics(real_mel, gen_mel, device="cuda"):
    """Same contract as reference val.compute_metrics (val.py:25-113): two (80, T) arrays ->
    dict of mse / ssim / avg_cos_sim / mean_error / std_error / snr rounded to 6 decimals."""
    real_mel, gen_mel = np.asarray(real_mel), np.asarray(gen_mel)
    min_t = min(real_mel.shape[1], gen_mel.shape[1])
    r = torch.from_numpy(np.ascontiguousarray(real_mel[:, :min_t], dtype=np.float32))[None]
    g = torch.from_numpy(np.ascontiguousarray(gen_mel[:, :min_t], dtype=np.float32))[None]
    return metrics_on_device(g.to(device), r.to(device))[0]


def _write_metrics(out_dir, base, metrics):
    with open(os.path.join(out_dir, f"{base}_metrics.txt"), "w") as f:
        f.write(f"sample: {base}\n" + "=" * 50 + "\n")
        for k, v in metrics.items():
            f.write(f"{k}: {v}\n")


def assess_single_sample(npz_path, ckpt_path, out_dir, device="cuda"):
    """val.py:164-245: sample one clip through sample.sample_from_npz (guidance 2.1), read the
    generated mel back, compute the metrics, write `<base>_metrics.txt` and copy the generated
    npz to `<base>_gen_mel.npz`. Returns (metrics, temp_dir)."""
    os.makedirs(out_dir, exist_ok=True)
    base = os.path.splitext(os.path.basename(npz_path))[0]
    real = _as_80_t(np.load(npz_path, allow_pickle=True)["mel"])
    temp_dir = os.path.join(out_dir, f"temp_{base}")
    os.makedirs(temp_dir, exist_ok=True)
    gen_npz = _sample.sample_from_npz(npz_path=npz_path, ckpt_path=ckpt_path, out_dir=temp_dir,
                                      device=device, timesteps=1000, guidance_weight=2.1)
    gen = _as_80_t(np.load(gen_npz, allow_pickle=True)["mel"])
    metrics = compute_metrics(real, gen, device=device)
    print(f"[{base}] metrics:")
    for k, v in metrics.items():
        print(f"  {k}: {v}")
    _write_metrics(out_dir, base, metrics)
    shutil.copy(gen_npz, os.path.join(out_dir, f"{base}_gen_mel.npz"))
    return metrics, temp_dir


def select_files(npz_dir, max_samples=None, random_sample=True, random_seed=42):
    """File selection of val.py:248-270 (same RNG calls -> same subset as the reference)."""
    files = [f for f in os.listdir(npz_dir) if f.endswith(".npz")]
    if random_sample and files:
        random.seed(random_seed)
        np.random.seed(random_seed)
        random.shuffle(files)
    if max_samples and max_samples < len(files):
        files = files[:max_samples]
    return files


@torch.no_grad()
def assess_batch(npz_dir, ckpt_path, out_dir, device="cuda", max_samples=None, random_sample=True,
                 random_seed=42, batch_size=32, guidance_weight=2.1, timesteps=1000):
    """val.py:248-319 with a persistent model: the checkpoint is loaded once, clips of equal mel
    length are sampled `batch_size` at a time (raw conditions resampled / projected on the GPU)
    and their metrics come from one lm2a_mel_metrics launch on the device-resident mels. Writes
    `<base>_metrics.txt`, `<base>_gen_mel.npz` (same keys as sample.py:250-256) and
    `average_metrics.txt`; returns (averaged metrics, {clip name: metrics})."""
    os.makedirs(out_dir, exist_ok=True)
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("lm2a_b200.val runs on a CUDA sm_100a device only (no CPU fallback)")
    files = select_files(npz_dir, max_samples, random_sample, random_seed)
    if not files:
        raise RuntimeError("no npz in " + npz_dir)
    unet, cond_proj = _sample.build_models(device=device)
    ck = _sample.load_checkpoint(ckpt_path, device=device)
    _sample.load_weights(unet, cond_proj, ck)
    mean, std = _sample.dataset_stats(ck)
    unet.eval()
    cond_proj.eval()
    steps = int(ck["timesteps"]) if ck.get("timesteps") is not None else timesteps
    gw = float(ck.get("guidance_weight", guidance_weight))
    diffusion = GaussianDiffusion(unet, timesteps=steps, device=device, dataset_mean=mean,
                                  dataset_std=std)
    clips = {}
    for f in files:
        d = np.load(os.path.join(npz_dir, f), allow_pickle=True)
        real, t_len = _sample._mel_length(d["mel"])
        clips.setdefault(t_len, []).append((os.path.splitext(f)[0], np.asarray(real, np.float32),
                                            np.asarray(d["motion"]), np.asarray(d["lyrics"]),
                                            int(d["sr"]) if "sr" in d else 22050,
                                            int(d["hop_length"]) if "hop_length" in d else 256))
    all_metrics = {}
    for t_len, group in sorted(clips.items()):
        for i in range(0, len(group), batch_size):
            part = group[i:i + batch_size]
            mel_norm, ex = _sample.sample_clips_raw(
                unet, cond_proj, diffusion, [c[2] for c in part], [c[3] for c in part], t_len, gw,
                want_resampled=True)
            real = torch.from_numpy(np.stack([c[1] for c in part])).to(device)
            gen_norm = torch.from_numpy(mel_norm).to(device)
            ms = metrics_on_device(gen_norm, real, std, mean)
            mel = mel_norm * np.float32(std) + np.float32(mean)
            for j, (c, m) in enumerate(zip(part, ms)):
                base = c[0]
                all_metrics[base] = m
                _write_metrics(out_dir, base, m)
                np.savez_compressed(
                    os.path.join(out_dir, f"{base}_gen_mel.npz"), mel=mel[j],
                    motion=ex["motion_rs"][j].cpu().numpy(), lyrics=ex["lyrics_rs"][j].cpu().numpy(),
                    motion_proj=ex["motion_f"][j:j + 1].float().cpu().numpy(),
                    lyrics_proj=ex["text_f"][j:j + 1].float().cpu().numpy(), sr=c[4],
                    hop_length=c[5])
    avg = {k: round(float(np.mean([m[k] for m in all_metrics.values()])), 6) for k in METRIC_KEYS}
    with open(os.path.join(out_dir, "average_metrics.txt"), "w") as f:
        f.write(f"samples: {len(all_metrics)}\nrandom: {random_sample}\nseed: {random_seed}\n")
        f.write("=" * 50 + "\naverage metrics:\n")
        for k, v in avg.items():
            f.write(f"{k}: {v}\n")
    print("average metrics:")
    for k, v in avg.items():
        print(f"  {k}: {v}")
    return avg, all_metrics


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--ckpt", default="checkpoints/ckpt_final_adan_500epoch.pt")
    p.add_argument("--npz_dir", default="npz_split/test")
    p.add_argument("--out_dir", default="testval")
    p.add_argument("--device", default="cuda")
    p.add_argument("--max_samples", type=int, default=10)
    p.add_argument("--no-random", action="store_false", dest="random_sample", default=True)
    p.add_argument("--seed", type=int, default=100)
    p.add_argument("--batch_size", type=int, default=32)
    return p.parse_args()


if __name__ == "__main__":
    a = parse_args()
    assess_batch(a.npz_dir, a.ckpt, a.out_dir, a.device, a.max_samples, a.random_sample, a.seed,
                 a.batch_size)
