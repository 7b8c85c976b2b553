"""Which implicit-GEMM launches of the step gain from another tile shape IN THE GRAPH (not timed
alone)? For every conv launch of the B = 32 CFG plan: switch it to each candidate
(block_n, cta_group), recapture the step graph, time the replays, print the delta against the
all-auto (wave model) baseline. Round-1 result: no candidate beats the wave model's choice in the
graph (profiles/r1_exp_tiles_in_graph.txt; that run also had a "lite" build — 128 x 128 tiles,
4 epilogue warps, 3 stages, two CTAs per SM — as candidate cg = 3, since removed).
    python tools/exp_tiles_in_graph.py [replays]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import lm2a_oracle as orc  # noqa: E402
from lm2a_b200.models import GaussianDiffusion, UNet1D_ultimate  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
B, T = 32, 516
dev = torch.device("cuda", 0)
net = UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8)
net.load_state_dict(orc.random_state_dict(orc.UNetConfig.production(), 5))
net = net.to(dev).eval()
diff = GaussianDiffusion(net, timesteps=1000, device=dev)
s = diff.sampler(B, T, T, guided=True)
s.gw = 2.1
g = torch.Generator().manual_seed(0)
s.set_conditions(torch.randn(B, T, 128, generator=g).to(dev), torch.randn(B, T, 128, generator=g).to(dev))


def step_ms():
    s._graphs.clear()
    s._ensure_graph()
    s.plan.x_in.normal_()
    s._reset_clock()
    for _ in range(5):
        s.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(N):
        s.graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N


base = min(step_ms(), step_ms())
print(f"baseline {base:.4f} ms")
keep = []
for i, (fn, args, meta) in enumerate(s.plan.ops):
    if meta["kind"] != "conv_gemm":
        continue
    d = args[0]
    if d.gn_gamma:          # fused GroupNorm launches keep their resident-tile shape
        continue
    best = (base, 0, 0)
    for bn, cg in ((128, 1), (256, 1), (128, 2), (256, 2)):
        if d.n_pad % bn:
            continue
        d.block_n, d.cta_group = bn, cg
        try:
            t = step_ms()
        except RuntimeError:
            continue
        if t < best[0]:
            best = (t, bn, cg)
        print(f"op {i:3d} m={meta['m']:6d} n={meta['n']:5d} k={meta['k']:5d}  bn={bn} cg={cg}: "
              f"{t:.4f} ms ({(t - base) * 1e3:+.1f} us)", flush=True)
    # greedy: keep a choice that wins by > 3 us, and move the baseline
    if best[0] < base - 0.003:
        d.block_n, d.cta_group = best[1], best[2]
        base = best[0]
        keep.append((i, meta["m"], meta["n"], meta["k"], best[1], best[2]))
    else:
        d.block_n, d.cta_group = 0, 0
print("kept:", keep)
print(f"final {min(step_ms(), step_ms()):.4f} ms")
