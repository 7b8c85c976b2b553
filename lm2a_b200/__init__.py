"""lm2a_b200 — B200-native (sm_100a) implementation of LM2A's classifier-free-guided
reverse-diffusion sampling path behind the reference's Python API.

    from lm2a_b200.models import UNet1D_ultimate, CondProjection, GaussianDiffusion
    from lm2a_b200.sample import sample_from_npz, build_models

The compute path is liblm2a_b200.so (hand-written CUDA: tcgen05/TMEM/TMA implicit-GEMM
conv, fused GroupNorm+SiLU, attention, CFG+posterior) reached through the C ABI in
include/lm2a_b200.h. There is no PyTorch, Triton or CPU fallback.
"""
__version__ = "0.1.0"
