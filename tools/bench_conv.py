"""Implicit-GEMM conv alone: time vs number of tile waves (fixed per-launch cost vs per-wave cost).
    python tools/bench_conv.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lm2a_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
BF16 = torch.bfloat16


def time_conv(m, n, cin, taps, block_n, cg, iters=20, stats=False, mode="graph"):
    tp = 128
    rows = m // tp
    m = rows * tp
    ntap = {ops.TAPS_K1: 1, ops.TAPS_K3: 3}[taps]
    x = torch.randn(m, cin, device=dev).to(BF16)
    w = (torch.randn(n, ntap * cin, device=dev) / (ntap * cin) ** 0.5).to(BF16)
    b = torch.zeros(n, device=dev)
    out = torch.zeros(m, n, dtype=BF16, device=dev)
    st = ops.Stats(rows, tp, n, 32, dev) if stats else None
    d = ops.make_conv_desc([ops.Seg(x, cin, cin, taps, m)], w, b, n, m, tp, tp - 1, out, n,
                           block_n=block_n, cta_group=cg, stats=st)
    ops.conv1d(d)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            ops.conv1d(d)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    fl = 2.0 * m * n * ntap * cin
    return m, us, fl / us / 1e6


for (n, cin, taps, name) in [(256, 256, ops.TAPS_K3, "n256 k768"), (1024, 1024, ops.TAPS_K3, "n1024 k3072"),
                             (512, 256, ops.TAPS_K1, "n512 k256")]:
    for cg in (1, 2):
        for bn in (256, 128):
            line = []
            for tiles in (1, 37, 74, 148, 296, 444, 592):
                m, us, tf = time_conv(128 * tiles * (n // 256 if False else 1), n, cin, taps, bn, cg)
                line.append(f"{tiles}:{us:.1f}us/{tf:.0f}TF")
            print(f"{name} cg{cg} bn{bn}  " + "  ".join(line), flush=True)
