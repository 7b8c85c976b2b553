"""Runs tools/probe/umma_probe.cu (built by the caller) and prints max errors per mode."""
import ctypes, os, torch
here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "umma_probe.so"))
torch.manual_seed(0)
P = torch.randn(128, 64, device="cuda").to(torch.bfloat16)
C = torch.randn(64, 128, device="cuda").to(torch.bfloat16)
ref = P.float() @ C.float()
for mode in (0, 1):
    O = torch.zeros(128, 128, device="cuda")
    rc = lib.umma_probe(ctypes.c_void_p(P.data_ptr()), ctypes.c_void_p(C.data_ptr()),
                        ctypes.c_void_p(O.data_ptr()), mode)
    err = (O - ref).abs().max().item()
    print(f"mode {mode}: rc={rc} max|err|={err:.4g} ref max={ref.abs().max().item():.3g}")
    if rc != 0:
        break
