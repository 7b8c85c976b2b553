"""Where do the implicit-GEMM mainloop cycles go? Needs the instrumented build:

    LM2A_NVCC_DEFS=-DLM2A_CONV_TIMING python -m lm2a_b200.build --force
    python tools/conv_stall_probe.py
    python -m lm2a_b200.build --force            # back to the product build

Per shape: fraction of the MMA-issuing thread's lifetime spent waiting for operands (full
barriers: the TMA data has not landed) and for a free accumulator (epilogue behind), and the
producer thread's wait for free stages (MMA behind), summed over CTAs by the kernel.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lm2a_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
BF16 = torch.bfloat16
lib = ctypes.CDLL(_lib.LIB_PATH)
if not hasattr(lib, "lm2a_conv_timing_read"):
    raise SystemExit("liblm2a_b200.so is not the instrumented build (see the docstring)")
buf = (ctypes.c_ulonglong * 8)()


def probe(rows, tp, n, cin, taps, block_n, cg, iters=20):
    m = rows * tp
    ntap = 3 if taps == ops.TAPS_K3 else 1
    x = torch.randn(m, cin, device=dev).to(BF16)
    w = (torch.randn(n, ntap * cin, device=dev) / (ntap * cin) ** 0.5).to(BF16)
    out = torch.zeros(m, n, dtype=BF16, device=dev)
    d = ops.make_conv_desc([ops.Seg(x, cin, cin, taps, m)], w, torch.zeros(n, device=dev), n, m, tp,
                           tp - 1, out, n, block_n=block_n, cta_group=cg)
    for _ in range(3):
        ops.conv1d(d)
    lib.lm2a_conv_timing_read(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.conv1d(d)
    e1.record()
    torch.cuda.synchronize()
    lib.lm2a_conv_timing_read(buf)
    full, acc, prod, life, nthr = (float(buf[i]) for i in range(5))
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"M={m} N={n} K={ntap * cin} bn={block_n} cg={cg}: {us:.1f} us/launch "
          f"({2.0 * m * n * ntap * cin / us / 1e6:.0f} TF/s); MMA thread: {100 * full / life:.1f}% waiting for "
          f"operands, {100 * acc / life:.1f}% for an accumulator; producer waits for a free stage "
          f"{100 * prod / life:.1f}% of the MMA thread's lifetime ({life / nthr / 1.9e3:.1f} us per CTA)")


probe(64, 130, 1024, 1024, ops.TAPS_K3, 256, 2)     # M 8320, the large up-path conv
probe(64, 130, 1024, 1024, ops.TAPS_K3, 128, 2)
probe(64, 130, 1024, 1024, ops.TAPS_K3, 256, 1)
probe(32, 65, 2048, 1024, ops.TAPS_K3, 128, 2)      # mid level q-projection
probe(32, 65, 1024, 2048, ops.TAPS_K1, 128, 2)      # mid level output GEMM
probe(32, 520, 256, 256, ops.TAPS_K3, 256, 2)       # level 0 cond rows
probe(256, 520, 1024, 1024, ops.TAPS_K3, 256, 2)    # many waves: steady state
