"""Cross-attention core alone at a production level's shape (cond rows of a B-clip CFG batch).
    python tools/bench_attn.py [level 0-3] [B] [iters]   -> avg us per launch (CUDA events)
Also the command profiled under ncu for profiles/*attn*.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from lm2a_b200 import ops  # noqa: E402

LEVELS = [(256, 516), (512, 258), (1024, 129), (1024, 64)]
lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
COND = len(sys.argv) > 4 and sys.argv[4] == "cond"   # heads attend to the raw condition slabs
e, t = LEVELS[lvl]
heads, lk = 8, 516
tp = [520, 260, 130, 65][lvl]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
slots = B + 1
q = (torch.randn(B * tp, 2 * e, generator=g, device=dev) * 0.5).to(torch.bfloat16)
o = torch.zeros_like(q)
kv = [torch.randn(slots * lk, 2 * e, generator=g, device=dev).to(torch.bfloat16) for _ in range(2)]
lk_pad = (lk + 7) // 8 * 8
vt = [torch.zeros(slots * e, lk_pad, dtype=torch.bfloat16, device=dev) for _ in range(2)]
for a, b in zip(kv, vt):
    ops.transpose_kv(a, 2 * e, e, b, lk_pad, slots, lk, e)
kv_slot = torch.arange(1, B + 1, dtype=torch.int32, device=dev)


cond = [(torch.randn(slots * lk, 128, generator=g, device=dev)).to(torch.bfloat16) for _ in range(2)]


def run():
    if COND:
        assert e // heads == 128
        ops.cross_attn_cond(q, 2 * e, o, 2 * e, ops._ptr(cond[0]), ops._ptr(cond[1]), 128, kv_slot,
                            slots, B, tp, t, lk, heads)
        return
    ops.cross_attn(q, 2 * e, o, 2 * e, ops._ptr(kv[0]), ops._ptr(vt[0]), ops._ptr(kv[1]),
                   ops._ptr(vt[1]), 2 * e, lk_pad, kv_slot, slots, B, tp, t, lk, e, heads)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / iters
flops = 2 * 4 * B * t * lk * e
print(f"level {lvl} E={e} T={t} dh={e // heads} B={B}: {us:.1f} us/launch, {flops / us / 1e6:.1f} TFLOP/s, "
      f"KV bytes {2 * 2 * B * lk * e * 2 / 1e6:.0f} MB -> {2 * 2 * B * lk * e * 2 / us / 1e3:.0f} GB/s")
