"""Experiment: one CFG step of B clips as S independent sub-batches captured on S streams of one
CUDA Graph (kernel tails / heads of different sub-batches overlap).
    python tools/exp_two_streams.py [B] [S]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import lm2a_oracle as orc  # noqa: E402
from lm2a_b200.models import GaussianDiffusion, UNet1D_ultimate  # noqa: E402
from lm2a_b200.models.diffusion import CfgSampler  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = 516
dev = torch.device("cuda", 0)
cfg = orc.UNetConfig.production()
net = UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8)
net.load_state_dict(orc.random_state_dict(cfg, 5))
net = net.to(dev).eval()
diff = GaussianDiffusion(net, timesteps=1000, device=dev)
g = torch.Generator().manual_seed(0)
subs = []
for i in range(S):
    s = CfgSampler(diff, B // S, T, T, True, True)
    s.plan = net.engine().plan(2 * (B // S), T, T, B // S + 1, 2, True, uniform_t=True,
                               uncond_rows=B // S) if i == 0 else \
        __import__("lm2a_b200.engine", fromlist=["UNetPlan"]).UNetPlan(
            net.engine().pm, 2 * (B // S), T, T, B // S + 1, 2, True, dev, True, B // S)
    s.gw = 2.1
    s.set_conditions(torch.randn(B // S, T, 128, generator=g).to(dev),
                     torch.randn(B // S, T, 128, generator=g).to(dev))
    s.plan.x_in.normal_()
    s.plan.t_in.fill_(999)
    subs.append(s)
torch.cuda.synchronize()
for s in subs:
    s._step(True)
torch.cuda.synchronize()
streams = [torch.cuda.Stream(device=dev) for _ in range(S - 1)]
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    cur = torch.cuda.current_stream(dev)
    for st in streams:
        st.wait_stream(cur)
    subs[0]._step(True)
    for st, s in zip(streams, subs[1:]):
        with torch.cuda.stream(st):
            s._step(True)
    for st in streams:
        cur.wait_stream(st)
for s in subs:
    s.plan.t_in.fill_(999)
for _ in range(5):
    graph.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    graph.replay()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"B={B} as {S} sub-batches on {S} streams: {ms:.3f} ms/step, {B / ms:.2f} clips/s "
      f"(finite={bool(torch.isfinite(subs[0].plan.x_in).all())})")
