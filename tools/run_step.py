"""Runs a few EAGER guided denoising steps (B clips, production net) — the command that is
profiled under ncu for profiles/ (every kernel of a step appears as its own launch)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import lm2a_oracle as orc  # noqa: E402
from lm2a_b200.models import GaussianDiffusion, UNet1D_ultimate  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = 516
dev = torch.device("cuda", 0)
cfg = orc.UNetConfig.production()
net = UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8)
net.load_state_dict(orc.random_state_dict(cfg, 5))
net = net.to(dev).eval()
diff = GaussianDiffusion(net, timesteps=1000, device=dev)
s = diff.sampler(B, T, T, guided=True)
s.gw = 2.1
g = torch.Generator().manual_seed(0)
s.set_conditions(torch.randn(B, T, 128, generator=g).to(dev), torch.randn(B, T, 128, generator=g).to(dev))
s.plan.x_in.normal_()
s.plan.t_in.fill_(999)
torch.cuda.synchronize()
from lm2a_b200 import ops  # noqa: E402
before = ops.launch_count()
for i in range(STEPS):
    s._step(True)
torch.cuda.synchronize()
per_step = (ops.launch_count() - before) // STEPS
# launches of OUR kernels before the last step / in one step (ncu -s / -c for a one-step capture)
print("ok", float(s.plan.x_in.std()), "skip", before + (STEPS - 1) * per_step, "per_step", per_step)
