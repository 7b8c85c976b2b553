"""val.py drop-in on B200: the metrics kernel against the oracle's restatement of
reference val.compute_metrics, and the persistent-model batched assess_batch end to end.
Tolerance: the kernel accumulates in fp64 (the reference in numpy fp32), results are compared
before rounding at 2e-6 absolute / 1e-6 relative; the reference itself rounds to 6 decimals."""
import os

import numpy as np
import pytest
import torch

import lm2a_oracle as orc

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _close(got, want, key):
    assert abs(got - want) <= 2e-6 + 1e-6 * abs(want), f"{key}: {got} vs {want}"


@pytest.mark.parametrize("t_len", [516, 48, 11])
def test_mel_metrics_kernel_vs_oracle(t_len):
    _need_gpu()
    from lm2a_b200 import val
    rng = np.random.default_rng(t_len)
    real = rng.normal(-4.6, 1.9, size=(4, 80, t_len)).astype(np.float32)
    gen_norm = rng.normal(0.0, 1.0, size=(4, 80, t_len)).astype(np.float32)
    gen_norm[1] = (real[1] + 4.5) / 2.0 + 0.01 * gen_norm[1]      # a close reconstruction
    real[3] = -3.0                                                  # constant real mel
    mean, std = -4.5, 2.0
    got = val.metrics_on_device(torch.from_numpy(gen_norm).cuda(), torch.from_numpy(real).cuda(),
                                std, mean)
    for b in range(4):
        gen = gen_norm[b] * np.float32(std) + np.float32(mean)
        want = orc.compute_metrics(real[b], gen)
        for k in val.METRIC_KEYS:
            _close(got[b][k], round(want[k], 6), f"clip {b} {k}")
    assert got[3]["snr"] == 0.0
    # public single-pair signature (val.py:25)
    one = val.compute_metrics(real[0], gen_norm[0] * np.float32(std) + np.float32(mean))
    assert one == got[0]
    with pytest.raises(RuntimeError, match="SSIM window"):
        val.metrics_on_device(torch.zeros(1, 80, 8).cuda(), torch.zeros(1, 80, 8).cuda())


def test_assess_batch_persistent_model(tmp_path):
    """Three clips of two different lengths, 3-step schedule from the checkpoint: files written,
    metrics equal the oracle's on the saved mels, resampled conditions bit-equal the host's."""
    _need_gpu()
    from lm2a_b200 import val
    npz_dir = tmp_path / "npz"
    os.makedirs(npz_dir)
    lens = {"a": 64, "b": 64, "c": 40}
    for i, (name, t_len) in enumerate(lens.items()):
        clip = orc.synthetic_clip(i, t_mel=t_len, t_motion=25 + i, time_varying_lyrics=True)
        np.savez(npz_dir / f"{name}.npz", **clip)
    ck = {"unet": orc.random_state_dict(orc.UNetConfig.production(), 5),
          "cond_proj": orc.random_cond_proj_state_dict(seed=7), "timesteps": 3,
          "guidance_weight": 2.1, "dataset_mean": -4.5, "dataset_std": 2.0}
    ckpt = tmp_path / "ck.pt"
    torch.save(ck, ckpt)
    out_dir = tmp_path / "out"
    avg, per = val.assess_batch(str(npz_dir), str(ckpt), str(out_dir), device="cuda",
                                max_samples=None, random_sample=True, random_seed=42, batch_size=2)
    assert set(per) == set(lens) and set(avg) == set(val.METRIC_KEYS)
    assert os.path.exists(out_dir / "average_metrics.txt")
    for name, t_len in lens.items():
        assert os.path.exists(out_dir / f"{name}_metrics.txt")
        g = np.load(out_dir / f"{name}_gen_mel.npz")
        src = np.load(npz_dir / f"{name}.npz")
        assert g["mel"].shape == (80, t_len) and np.isfinite(g["mel"]).all()
        np.testing.assert_array_equal(g["motion"], orc.match_len_interp(src["motion"], t_len))
        want = orc.compute_metrics(src["mel"], g["mel"])
        for k in val.METRIC_KEYS:
            _close(per[name][k], round(want[k], 6), f"{name} {k}")
    for k in val.METRIC_KEYS:
        _close(avg[k], round(float(np.mean([per[n][k] for n in lens])), 6), k)
    # file selection follows the reference's RNG calls (val.py:255-262)
    import random
    files = [f for f in os.listdir(npz_dir) if f.endswith(".npz")]
    random.seed(42)
    random.shuffle(files)
    assert val.select_files(str(npz_dir), 2, True, 42) == files[:2]


def test_assess_single_sample_contract(tmp_path):
    """val.assess_single_sample (val.py:164-245): samples through sample_from_npz (schedule length
    from the checkpoint), returns (metrics, temp_dir) and writes the per-sample files."""
    _need_gpu()
    from lm2a_b200 import val
    clip = orc.synthetic_clip(3, t_mel=48, t_motion=20)
    npz = tmp_path / "clipx.npz"
    np.savez(npz, **clip)
    ck = {"unet": orc.random_state_dict(orc.UNetConfig.production(), 5),
          "cond_proj": orc.random_cond_proj_state_dict(seed=7), "timesteps": 3,
          "guidance_weight": 2.1, "dataset_mean": -4.5, "dataset_std": 2.0}
    ckpt = tmp_path / "ck.pt"
    torch.save(ck, ckpt)
    out_dir = tmp_path / "out"
    metrics, temp_dir = val.assess_single_sample(str(npz), str(ckpt), str(out_dir), device="cuda")
    assert set(metrics) == set(val.METRIC_KEYS)
    assert os.path.isdir(temp_dir) and os.path.exists(out_dir / "clipx_metrics.txt")
    g = np.load(out_dir / "clipx_gen_mel.npz")
    want = orc.compute_metrics(clip["mel"], g["mel"])
    for k in val.METRIC_KEYS:
        _close(metrics[k], round(want[k], 6), k)
