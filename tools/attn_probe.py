"""Where the resident attention kernel's time goes (instrumented build: -DLM2A_ATTN_TIMING into
tools/probe/liblm2a_b200_probe.so; run with LM2A_LIB_PATH pointing at it).
    python tools/attn_probe.py build          (here, no GPU needed)
    LM2A_LIB_PATH=tools/probe/liblm2a_b200_probe.so python tools/attn_probe.py run [level] [cond]
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PROBE = os.path.join(ROOT, "tools", "probe", "liblm2a_b200_probe.so")

if sys.argv[1] == "build":
    from lm2a_b200 import build
    print(build.build(out=PROBE, defs=["-DLM2A_ATTN_TIMING"]))
    sys.exit(0)

os.environ.setdefault("LM2A_LIB_PATH", PROBE)
from lm2a_b200 import _lib  # noqa: E402
lib = _lib.load()
args = sys.argv[2:]
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_attn.py")] + args,
                     capture_output=True, text=True, env=os.environ)
print(out.stdout.strip(), out.stderr.strip()[-300:])
# the counters live in the child process: run the kernel here instead
import torch  # noqa: E402
sys.argv = ["bench_attn.py"] + args
buf = (ctypes.c_ulonglong * 16)()
lib.lm2a_attn_timing_read.restype = ctypes.c_int
lib.lm2a_attn_timing_read(buf)   # reset
exec(open(os.path.join(ROOT, "tools", "bench_attn.py")).read())
torch.cuda.synchronize()
lib.lm2a_attn_timing_read(buf)
v = list(buf)
n, ctas = max(v[6], 1), max(v[7], 1)
print(f"softmax warp 0 / slot 0, cycles per chunk: loop {v[1] / n:.0f} = wait S {v[0] / n:.0f} + "
      f"S load {v[8] / n:.0f} + wait P buffer {v[2] / n:.0f} + P store {v[9] / n:.0f} + rest")
print(f"issuer slot 0, cycles per chunk: lifetime {v[5] / n:.0f} = wait P {v[3] / n:.0f} + "
      f"wait S buffer {v[4] / n:.0f} + rest;  CTAs {ctas}, chunks {n}")
