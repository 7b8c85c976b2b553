"""Builds liblm2a_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblm2a_b200.so")
SOURCES = ["api.cu", "conv_gemm.cu", "gn_silu.cu", "attention_tc.cu", "attention_res.cu", "attention_tail.cu", "elementwise.cu",
           "step_update.cu", "ref_f32.cu"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "lm2a_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, out=None, defs=()):
    """out / defs: a probe build (extra -D flags) into another file, objects in a scratch dir."""
    if out is not None:
        return _build(out, list(defs), verbose, os.path.join(os.path.dirname(out), "obj"))
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.environ.get("LM2A_NVCC_DEFS", "").split(), verbose, CSRC)


def _build(lib, extra, verbose, objdir):
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
               "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v" if verbose else "-O3", *extra,
               "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + [
        "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
