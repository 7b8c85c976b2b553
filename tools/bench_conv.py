"""Device time of single implicit-GEMM launches at the production plan's shapes (B = 32 CFG),
with and without the GroupNorm operand transform: each launch captured `iters` times back to
back in a CUDA Graph, replay timed with CUDA events.
    python tools/bench_conv.py [tag]
Environment knobs read by the library (one process per variant): LM2A_CONV_SHARE_TAPS=0|1,
LM2A_CONV_DBG_NOSHIFT=1 (timing only), LM2A_CONV_CG=1|2.
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lm2a_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
BF16 = torch.bfloat16
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
ONLY = sys.argv[2] if len(sys.argv) > 2 else None      # substring filter on the shape name
SILU = os.environ.get("BENCH_GN_SILU", "1") != "0"

# (name, rows, T, Tp, cin, cout, taps, skip_cin, in_gn)
SHAPES = [
    ("l0 conv1 256->256 k3", 32, 516, 520, 256, 256, "k3", 0, True),
    ("l0 conv1 256->256 k3 raw", 32, 516, 520, 256, 256, "k3", 0, False),
    ("l0 q-proj 256->512 k3", 32, 516, 520, 256, 512, "k3", 0, True),
    ("l1 conv2 512->512 k3+skip256", 32, 258, 260, 512, 512, "k3", 256, True),
    ("l2 conv2 1024 k3+skip512", 64, 129, 130, 1024, 1024, "k3", 512, True),
    ("l2 conv2 1024 k3+skip512 raw", 64, 129, 130, 1024, 1024, "k3", 512, False),
    ("l3 conv1 1024 k3", 64, 64, 65, 1024, 1024, "k3", 0, True),
    ("l3 conv1 1024 k3 raw", 64, 64, 65, 1024, 1024, "k3", 0, False),
    ("l3 q-proj 1024->2048 k3", 32, 64, 65, 1024, 2048, "k3", 0, True),
    ("l2 up-conv 1024 k3 raw", 64, 129, 130, 1024, 1024, "k3", 0, False),
    ("l2 out 2048->1024 k1+skip2048 raw", 32, 129, 130, 2048, 1024, "k1", 2048, False),
    ("l0 out 512->256 k1 raw", 32, 516, 520, 512, 256, "k1", 0, False),
    ("l1 cat conv1 2048->... k3", 64, 129, 130, 2048, 1024, "k3", 0, True),
    ("l0 out_proj 256->80 k1", 64, 516, 520, 256, 128, "k1", 0, True),
]


def time_op(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


print(f"# variant {tag}: SHARE_TAPS={os.environ.get('LM2A_CONV_SHARE_TAPS')} "
      f"NOSHIFT={os.environ.get('LM2A_CONV_DBG_NOSHIFT')} CG={os.environ.get('LM2A_CONV_CG')}")
for name, r, t, tp, cin, cout, taps, skip_c, in_gn in SHAPES:
    if ONLY and ONLY not in name:
        continue
    m = r * tp
    k = 3 if taps == "k3" else 1
    tk = ops.TAPS_K3 if taps == "k3" else ops.TAPS_K1
    x = (torch.randn(m, cin, device=dev) * 0.5).to(BF16)
    x.view(r, tp, cin)[:, t:, :] = 0
    ktot = k * cin + skip_c
    n_pad = (cout + 127) // 128 * 128
    w = (torch.randn(n_pad, ktot, device=dev) / math.sqrt(ktot)).to(BF16)
    bias = torch.zeros(n_pad, device=dev)
    segs = [ops.Seg(x, cin, cin, tk, m)]
    if skip_c:
        xs = (torch.randn(m, skip_c, device=dev) * 0.5).to(BF16)
        segs.append(ops.Seg(xs, skip_c, skip_c, ops.TAPS_K1, m))
    out = torch.zeros(m, cout, dtype=BF16, device=dev)
    st_out = None if os.environ.get("BENCH_NO_STATS") == "1" else ops.Stats(r, cout, 8, dev)
    gn = None
    if in_gn:
        st = ops.Stats(r, cin, 8, dev)
        ops.bias_add(x, cin, 0, torch.empty_like(x), cin, 0, torch.zeros(cin, device=dev), m, tp, t,
                     cin, st)
        gn = (st, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 1e-5, SILU)
    d = ops.make_conv_desc(segs, w, bias, cout, m, tp, t, out, cout, stats=st_out, in_gn=gn)
    sec = time_op(lambda: ops.conv1d(d))
    fl = 2.0 * r * t * cout * ktot
    print(f"{name:38s} M={m:6d} N={cout:5d} K={ktot:5d} {sec * 1e6:7.1f} us {fl / sec / 1e12:7.1f} TF/s")
