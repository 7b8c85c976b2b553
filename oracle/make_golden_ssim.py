"""Generates tests/golden/ssim.npz: scikit-image's structural_similarity exactly as reference
val.py:43-66 calls it, on seeded mel pairs — the pin for the SSIM part of oracle.compute_metrics
and of lm2a_mel_metrics.

scikit-image is NOT in this build image (no wheel in the offline wheelhouse) and the reference
pins no version, so this script could not be run here and the fixture is NOT committed: the
SSIM number is therefore excluded from the parity claim (the other five
metrics of val.compute_metrics are pinned through numpy). Run it wherever scikit-image >= 0.19 is
installed to add the pin; tests/test_oracle_golden.py::test_ssim_matches_skimage_golden picks
the file up when it exists.

    python oracle/make_golden_ssim.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    from skimage.metrics import structural_similarity as ssim   # noqa: raises where absent
    import skimage
    rng = np.random.default_rng(20240501)
    cases = {}
    for i, (t, noise) in enumerate([(516, 0.3), (129, 1.0), (40, 0.05)]):
        real = rng.normal(-4.6, 1.9, size=(80, t)).astype(np.float32)
        gen = (real + rng.normal(0, noise, size=real.shape)).astype(np.float32)
        # val.py:43-66: min-max normalise with the REAL mel's range, clip, SSIM over channel_axis 0
        lo, hi = real.min(), real.max()
        rn = np.clip((real - lo) / (hi - lo + 1e-8), 0, 1)
        gn = np.clip((gen - lo) / (hi - lo + 1e-8), 0, 1)
        val = ssim(rn, gn, data_range=1.0, channel_axis=0, gaussian_weights=True, sigma=1.5,
                   use_sample_covariance=False)
        cases[f"real_{i}"], cases[f"gen_{i}"], cases[f"ssim_{i}"] = real, gen, np.float64(val)
    cases["n"] = np.int64(3)
    cases["skimage_version"] = np.array(skimage.__version__)
    out = os.path.join(ROOT, "tests", "golden", "ssim.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out)


if __name__ == "__main__":
    main()
