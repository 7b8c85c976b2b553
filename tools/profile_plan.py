"""Per-launch device times of one CFG denoising step (B clips, production net), CUDA events.
    python tools/profile_plan.py [B] [T] [tiled] > profiles/plan_Bxx.csv
`tiled`: lyrics condition tiled over time (the reference's real data) -> one-stream launch list.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import lm2a_oracle as orc  # noqa: E402
from lm2a_b200.models import GaussianDiffusion, UNet1D_ultimate  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = int(sys.argv[2]) if len(sys.argv) > 2 else 516
dev = torch.device("cuda", 0)
cfg = orc.UNetConfig.production()
net = UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8)
net.load_state_dict(orc.random_state_dict(cfg, 5))
net = net.to(dev).eval()
diff = GaussianDiffusion(net, timesteps=1000, device=dev)
s = diff.sampler(B, T, T, guided=True)
s.gw = 2.1
g = torch.Generator().manual_seed(0)
tiled = len(sys.argv) > 3 and sys.argv[3] == "tiled"
mf = torch.randn(B, T, 128, generator=g)
tf = torch.randn(B, T, 128, generator=g)
if tiled:
    tf = tf[:, :1].expand(B, T, 128).contiguous()
s.set_conditions(mf.to(dev), tf.to(dev))
assert s.plan.const_text is tiled
s.plan.x_in.normal_()
s.plan.t_in.fill_(500)
prof = s.plan.profile(iters=10)
print("idx,kind,m,n,k,gflop,us,tflops")
tot = {}
for i, (kind, meta, sec) in enumerate(prof):
    gf = meta.get("flops", 0) / 1e9
    print(f"{i},{kind},{meta.get('m', '')},{meta.get('n', '')},{meta.get('k', '')},{gf:.3f},{sec * 1e6:.1f},"
          f"{gf / 1e3 / sec if sec > 0 else 0:.1f}")
    tot[kind] = tot.get(kind, 0) + sec
all_s = sum(tot.values())
print("# totals (serialised launches):", {k: round(v * 1e3, 3) for k, v in tot.items()}, "ms; sum",
      round(all_s * 1e3, 3), "ms", file=sys.stderr)
print("# step GFLOP", s.plan.flops() / 1e9, file=sys.stderr)
