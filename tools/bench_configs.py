"""Device-timed CFG denoising step of the other BASELINE.json configs (the bench.py headline is
configs[1]): config 3 (B = 64), config 5 (long clip T = Lk = 2064; legacy UNet1D base 256),
plus the 50-step DDIM sampler. One JSON object per line on stdout.

    python tools/bench_configs.py [steps]

Each case: graph-captured step (UNet on [uncond, cond] rows + CFG blend + posterior update),
5 warm-up replays, `steps` timed replays between CUDA events on the launching stream.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import lm2a_oracle as orc  # noqa: E402
from lm2a_b200.models import GaussianDiffusion, UNet1D, UNet1D_ultimate  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 30
dev = torch.device("cuda", 0)


def time_case(name, net, batch, t_len, lk, ddim=None):
    diff = GaussianDiffusion(net, timesteps=1000, device=dev)
    kw = {}
    if ddim:
        kw["ddim"] = (tuple(diff.ddim_timesteps(ddim)), 0.0)
    s = diff.sampler(batch, t_len, lk, True, **kw)
    s.gw = 2.1
    g = torch.Generator().manual_seed(0)
    s.set_conditions(torch.randn(batch, lk, 128, generator=g).to(dev),
                     torch.randn(batch, lk, 128, generator=g).to(dev))
    s._ensure_graph()
    s.plan.x_in.normal_()
    s._reset_clock()
    n = min(STEPS, len(s.taus) - 6) if ddim else STEPS
    for _ in range(5):
        s.graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.current_stream(dev)
    e0.record(st)
    for _ in range(n):
        s.graph.replay()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    traj = ddim or 1000
    flops = s.plan.flops()
    out = {"case": name, "batch": batch, "rows": 2 * batch, "T": t_len, "Lk": lk,
           "ms_per_step": ms, "steps_per_trajectory": traj,
           "clips_per_s": batch / (traj * ms * 1e-3), "step_gflop": flops / 1e9,
           "step_tflops": flops / (ms * 1e-3) / 1e12, "launches_per_step": len(s.plan.ops) + 1,
           "finite": bool(torch.isfinite(s.plan.x_in).all())}
    print(json.dumps(out), flush=True)
    del s, diff
    torch.cuda.empty_cache()


def main():
    cfg = orc.UNetConfig.production()
    net = UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8)
    net.load_state_dict(orc.random_state_dict(cfg, 5))
    net = net.to(dev).eval()
    time_case("config2_B32_T516", net, 32, 516, 516)
    time_case("config3_B64_T516", net, 64, 516, 516)
    time_case("ddim50_B32_T516", net, 32, 516, 516, ddim=50)
    net._engine = None
    torch.cuda.empty_cache()
    time_case("config5_longclip_B8_T2064_Lk2064", net, 8, 2064, 2064)
    del net
    torch.cuda.empty_cache()
    lcfg = orc.LegacyConfig(80, 256, (1, 2, 4), 128, 256)
    leg = UNet1D(80, 256, (1, 2, 4), 128, 256)
    leg.load_state_dict(orc.legacy_random_state_dict(lcfg, 22))
    leg = leg.to(dev).eval()
    time_case("config5_legacy_unet1d_B32_T516", leg, 32, 516, 516)


if __name__ == "__main__":
    main()
