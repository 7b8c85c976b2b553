"""Launches the kernels either side of the denoising loop at production sizes (B = 32 clips):
lm2a_resample_seq (motion 180 -> 516 x 234, lyrics 516 x 768), lm2a_cfg_ddim, lm2a_mel_metrics —
the command profiled under ncu for profiles/*aux*; prints CUDA-event times per launch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lm2a_b200 import ops  # noqa: E402
from lm2a_b200.models import GaussianDiffusion  # noqa: E402

B, T = 32, 516
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
motion = torch.randn(B, 180, 234, generator=g, device=dev)
lyrics = torch.randn(B, 516, 768, generator=g, device=dev)
m_rs = torch.empty(B, T, 234, device=dev)
l_rs = torch.empty(B, T, 768, device=dev)
m_slab = torch.empty(B * T, 256, dtype=torch.bfloat16, device=dev)
l_slab = torch.empty(B * T, 768, dtype=torch.bfloat16, device=dev)
x = torch.randn(B, 80, T, generator=g, device=dev)
eps = torch.randn(2 * B, 80, T, generator=g, device=dev)
noise = torch.randn(B, 80, T, generator=g, device=dev)
real = torch.randn(B, 80, T, generator=g, device=dev)
diff = GaussianDiffusion(None, timesteps=1000, device=dev)
table = diff.ddim_coefficients(999, 979, 0.5).contiguous()
step = torch.zeros(1, dtype=torch.int32, device=dev)
out = torch.empty(B, 8, dtype=torch.float64, device=dev)


def timed(name, fn, bytes_moved, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"{name}: {us:.1f} us/launch, algorithmic {bytes_moved / 1e6:.1f} MB -> "
          f"{bytes_moved / us / 1e3:.0f} GB/s")


timed("resample_seq motion 180->516 x234", lambda: ops.resample_seq(
    motion, None, m_rs, m_slab, B, 180, 234, T, T, 256),
    B * T * 234 * (8 + 4) + B * T * 256 * 2)
timed("resample_seq lyrics 516->516 x768", lambda: ops.resample_seq(
    lyrics, None, l_rs, l_slab, B, 516, 768, T, T, 768), B * T * 768 * (4 + 4 + 2))
timed("cfg_ddim (guided, eta>0)", lambda: ops.cfg_ddim(
    x, eps, noise, table, None, step, None, None, B, 80 * T, 2.1, True, False),
    B * 80 * T * 4 * 5)
timed("mel_metrics", lambda: ops.mel_metrics(x, real, out, B, 80, T, 2.0, -4.5),
      B * 80 * T * 4 * 2)

# fused Adan + EMA over every tensor of the production UNet (306 tensors, 134.3 M parameters)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lm2a_oracle as orc  # noqa: E402
from lm2a_b200.models.adan import Adan  # noqa: E402

spec = orc.state_dict_spec(orc.UNetConfig.production())
params = [torch.nn.Parameter(torch.randn(shape, device=dev) * 0.02) for _, shape in spec]
shadow = [p.detach().clone() for p in params]
for p in params:
    p.grad = torch.randn_like(p) * 0.01
opt = Adan(params, lr=2e-4, weight_decay=1e-4)
n_el = sum(p.numel() for p in params)
timed(f"adan_step + EMA ({len(params)} tensors, {n_el / 1e6:.1f} M params)",
      lambda: opt.step(ema=(shadow, 0.999)), n_el * 52, iters=10)
