"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the
drop-in modules keep the reference's state_dict layout, slab geometry / launch plan host logic,
and the clip-sharding driver under gloo with world_size 2. No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import lm2a_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from lm2a_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "lm2a_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lm2a_[a-z0-9_]+)\s*\(", code))
    assert len(declared) >= 17
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported by liblm2a_b200.so"
    assert declared == set(_lib.SIGNATURES), "ctypes table out of sync with the header"
    assert _lib.load().lm2a_abi_version() == _lib.ABI_VERSION


def test_conv_desc_struct_matches_header_layout():
    from lm2a_b200 import _lib
    assert ctypes.sizeof(_lib.ConvSeg) == 32
    assert ctypes.sizeof(_lib.ConvDesc) == 232 and _lib.ConvDesc.k_order.offset == 224
    assert _lib.ConvDesc.out.offset == 136 and _lib.ConvDesc.m.offset == 80
    assert _lib.ConvDesc.stats.offset == 152 and _lib.ConvDesc.in_gn_stats.offset == 176


def test_state_dict_layout_matches_reference_inventory():
    from lm2a_b200.models import CondProjection, UNet1D_ultimate
    net = UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8)
    spec = orc.state_dict_spec(orc.UNetConfig.production())  # pinned to the reference in make_golden
    sd = net.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(tuple(sd[k].shape) == s for k, s in spec)
    assert sum(v.numel() for v in sd.values()) == 134292816
    # class defaults are the reference's (base 128, 4 heads: 35.4 M params)
    small = UNet1D_ultimate()
    assert abs(sum(p.numel() for p in small.parameters()) - 35.4e6) < 0.1e6
    assert set(CondProjection().state_dict()) == {"motion_proj.weight", "motion_proj.bias",
                                                  "text_proj.weight", "text_proj.bias"}
    res = net.load_state_dict({"in_proj.bias": torch.zeros(256)}, strict=False)  # silent partial load
    assert len(res.missing_keys) == 305


def test_torch_custom_ops_are_cuda_only():
    """torch.ops.lm2a.* (lm2a_b200/torch_ops.py) exist with CUDA kernels only: a CPU tensor finds no
    implementation (no fallback was registered)."""
    import lm2a_b200.torch_ops  # noqa: F401
    for name in ("cfg_posterior", "cfg_ddim", "resample_seq", "mel_metrics", "gn_silu", "upsample2x",
                 "conv1d", "cross_attn", "cross_attn_cond", "cross_attn_tail", "transpose_kv",
                 "time_mlp", "film", "philox_normal"):
        assert hasattr(torch.ops.lm2a, name)
    with pytest.raises(NotImplementedError):
        torch.ops.lm2a.resample_seq(torch.zeros(1, 4, 3), None, 8)


def test_cpu_tensors_are_refused():
    from lm2a_b200.models import UNet1D_ultimate
    net = UNet1D_ultimate(80, 64, (1, 2, 4), 128, 256, 2, 3, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 80, 64), torch.zeros(1, dtype=torch.long))


def test_geometry_and_plan(monkeypatch):
    from lm2a_b200.engine import Geometry, PackedModel, UNetPlan
    from lm2a_b200.models import UNet1D_ultimate
    g = Geometry(64, 516, 3)
    assert g.T == [516, 258, 129, 64] and g.Tp == [520, 260, 130, 65]
    g = Geometry(2, 517, 3)
    assert g.T == [517, 258, 129, 64] and all(tp >= t + 1 for tp, t in zip(g.Tp, g.T))
    with pytest.raises(RuntimeError):
        Geometry(1, 7, 3)
    net = UNet1D_ultimate(80, 64, (1, 2, 4), 128, 256, 2, 3, 2)
    pm = PackedModel(net, torch.device("cpu"))
    plan = UNetPlan(pm, 4, 132, 132, 3, 2, True, torch.device("cpu"))
    kinds = [m["kind"] for _, _, m in plan.ops]
    # at base 64 every GroupNorm + SiLU up to 256 channels runs inside the conv that consumes it
    # (operand transform): 30 of the 47 GEMM launches carry one; the single wider operand (the
    # 512-channel concat slab of the deepest up block) takes a streaming gn_apply pass
    assert kinds.count("conv_gemm") == 47 and kinds.count("gn_apply") == 1
    assert sum(1 for _, _, m in plan.ops if m.get("in_gn")) == 30
    assert kinds.count("cross_attn") == 9 and len(plan.kv_ops) == 36
    assert kinds[:3] == ["time_mlp", "film", "ingest_x"] and len(plan.ops) == 60
    # opt-in (LM2A_ATTN_TAIL=1): T = 132 leaves 4 query rows over at level 0; the two level-0
    # attention blocks put them on the CUDA cores, on the parallel branch next to the tensor-core
    # launch (whole tiles only), joined in front of the output GEMM
    monkeypatch.setenv("LM2A_ATTN_TAIL", "1")
    tplan = UNetPlan(pm, 4, 132, 132, 3, 2, True, torch.device("cpu"))
    monkeypatch.delenv("LM2A_ATTN_TAIL")
    tk = [m["kind"] for _, _, m in tplan.ops]
    tails = [i for i, k in enumerate(tk) if k == "cross_attn_tail"]
    assert len(tails) == 2 and len(tplan.ops) == 62
    for i in tails:
        (_, targs, tmeta), (_, margs, mmeta), (_, _, nmeta) = tplan.ops[i], tplan.ops[i + 1], tplan.ops[i + 2]
        assert tmeta.get("side") and mmeta["kind"] == "cross_attn" and not mmeta.get("join")
        assert nmeta["kind"] == "conv_gemm" and nmeta.get("join")
        assert targs[-7:-5] == (128, 4) and margs[-5] == 128    # t0, n_tail | t_valid of the main launch
    # the x2 interpolation of the three UpSampleConvs runs inside their convs
    assert kinds.count("upsample2x") == 0 and sum(1 for _, _, m in plan.ops if m.get("up2x")) == 3
    # (clips too short for the operand transform would fall back to the stand-alone gn_apply
    # pass; with <= 8 groups every clip length the geometry admits is long enough)
    from lm2a_b200 import ops as _ops
    assert _ops.in_gn_supported(3, 8) and not _ops.in_gn_supported(3, 32)
    # K/V hoisted, out_proj.fuse folded; the composed conv2 o q-proj GEMM executes 29.58 GFLOP
    # per row-step in total, of which 28.23 are the reference's algorithmic work (credited)
    big = PackedModel(UNet1D_ultimate(80, 256, (1, 2, 4), 128, 256, 2, 3, 8), torch.device("cpu"))
    bp = UNetPlan(big, 2, 516, 516, 2, 2, True, torch.device("cpu"))
    assert abs(bp.flops() / 2 / 1e9 - 28.23) < 0.02
    assert abs(sum(m.get("flops_executed", m["flops"]) for _, _, m in bp.ops) / 2 / 1e9 - 29.58) < 0.02
    # production CFG step (B = 32, uncond shortcut, shared leading rows): operands of 256 channels
    # (level 0) are normalised / upsampled inside their conv (9 + 1 launches), the wider ones
    # (MMA-bound GEMMs) by 22 streaming gn_apply and 2 upsample2x passes in front of plain
    # launches; the five d_h = 128 attention blocks attend to the raw condition slabs
    cfgp = UNetPlan(big, 64, 516, 516, 33, 2, True, torch.device("cpu"), uniform_t=True,
                    uncond_rows=32)
    ck = [m["kind"] for _, _, m in cfgp.ops]
    assert len(cfgp.ops) == 92 and abs(cfgp.flops() / 1e9 - 1185.0) < 0.5
    assert ck.count("cross_attn_tail") == 0 and ck.count("cross_attn") == 9
    assert ck.count("gn_apply") == 22 and ck.count("upsample2x") == 2
    assert sum(1 for _, _, m in cfgp.ops if m.get("in_gn")) == 9
    assert sum(1 for _, _, m in cfgp.ops if m.get("cond")) == 5
    with pytest.raises(RuntimeError, match="multiples of 64"):
        PackedModel(UNet1D_ultimate(80, 16, (1, 2, 4), 32, 32, 2, 3, 4), torch.device("cpu"))


def test_legacy_model_layout_and_plan():
    """Legacy UNet1D mirror keeps the reference's state_dict (models/unet1d.py) and its launch
    plan has the expected shape; the transposed-conv even/odd split is exact in fp64."""
    from lm2a_b200.engine import PackedLegacy, UNetPlan, _convT_w
    from lm2a_b200.models import UNet1D
    cfg = orc.LegacyConfig(80, 128, (1, 2, 4), 128, 256)
    net = UNet1D(80, 128, (1, 2, 4), 128, 256)
    spec = orc.legacy_state_dict_spec(cfg)   # pinned to the reference by make_golden_legacy.py
    sd = net.state_dict()
    assert list(sd.keys()) == [k for k, _ in spec]
    assert all(tuple(sd[k].shape) == shp for k, shp in spec)
    with pytest.raises(RuntimeError, match="motion_f"):
        net(torch.zeros(1, 80, 64), torch.zeros(1, dtype=torch.long), None, None)
    pm = PackedLegacy(net, torch.device("cpu"))
    assert [b.e // b.heads for b in pm.attn_blocks] == [32, 32, 64, 128, 192, 96, 64]
    plan = UNetPlan(pm, 4, 100, 60, 3, 2, True, torch.device("cpu"), uniform_t=True, uncond_rows=2)
    kinds = [m["kind"] for _, _, m in plan.ops]
    assert kinds.count("cross_attn") == 7 and kinds.count("bias_add") == 7
    assert len(plan.kv_ops) == 28
    # ConvTranspose1d k4 s2 p1 == two k3 GEMMs over (x[m-1], x[m], x[m+1]) -> slots 2m / 2m+1
    g = torch.Generator().manual_seed(5)
    w = torch.randn(6, 4, 4, generator=g, dtype=torch.float64)
    x = torch.randn(2, 6, 9, generator=g, dtype=torch.float64)
    ref = torch.nn.functional.conv_transpose1d(x, w, stride=2, padding=1)
    even, odd = _convT_w(w)
    xp = torch.nn.functional.pad(x, (1, 1))
    cols = torch.cat([xp[:, :, 0:9], xp[:, :, 1:10], xp[:, :, 2:11]], dim=1)  # [B, 3*Cin, T]
    got = torch.stack([torch.einsum("ok,bkt->bot", even, cols),
                       torch.einsum("ok,bkt->bot", odd, cols)], dim=-1).reshape(2, 4, 18)
    assert float((got - ref).abs().max()) < 1e-12


def test_weight_folds_are_exact_in_fp64():
    """q-scale, kv_proj o in_proj and out_proj o fuse_proj folds (engine.pack_block) reproduce
    CrossAttentionFusion (cross_attention.py:38-67) when evaluated in fp64."""
    from lm2a_b200 import engine
    cfg = orc.UNetConfig(80, 64, (1,), 128, 64, 1, 1, 2)
    sd = orc.cast_state_dict(orc.random_state_dict(cfg, 2), torch.float64)
    pre = "mid.blocks.0.cross_attn"
    e, heads, lk, tq = 64, 2, 11, 7
    g = torch.Generator().manual_seed(0)
    h = torch.randn(1, tq, e, generator=g, dtype=torch.float64)
    mf = torch.randn(1, lk, 128, generator=g, dtype=torch.float64)
    tf = torch.randn(1, lk, 128, generator=g, dtype=torch.float64)
    ref = orc.cross_attention_fusion(sd, pre, h, mf, tf, heads)

    class Holder:
        pass
    outs = []
    for s, cond in (("attn_motion", mf), ("attn_text", tf)):
        ipw, ipb = sd[f"{pre}.{s}.in_proj_weight"], sd[f"{pre}.{s}.in_proj_bias"]
        kvn = "motion_kv_proj" if s == "attn_motion" else "text_kv_proj"
        wp, bp = sd[f"{pre}.{kvn}.weight"], sd[f"{pre}.{kvn}.bias"]
        qs = engine.LOG2E / np.sqrt(e // heads)
        q = (h @ (ipw[:e] * qs).T + ipb[:e] * qs).view(1, tq, heads, -1).transpose(1, 2)
        k = (cond @ (ipw[e:2 * e] @ wp).T + ipw[e:2 * e] @ bp + ipb[e:2 * e]).view(1, lk, heads, -1).transpose(1, 2)
        v = (cond @ (ipw[2 * e:] @ wp).T + ipw[2 * e:] @ bp + ipb[2 * e:]).view(1, lk, heads, -1).transpose(1, 2)
        sc = q @ k.transpose(-1, -2)
        p = torch.exp2(sc - sc.max(-1, keepdim=True).values)
        outs.append(((p / p.sum(-1, keepdim=True)) @ v).transpose(1, 2).reshape(1, tq, e))
    wf, bf = sd[f"{pre}.fuse_proj.weight"], sd[f"{pre}.fuse_proj.bias"]
    wo = [sd[f"{pre}.{s}.out_proj.weight"] for s in ("attn_motion", "attn_text")]
    bo = [sd[f"{pre}.{s}.out_proj.bias"] for s in ("attn_motion", "attn_text")]
    wof = torch.cat([wf[:, :e] @ wo[0], wf[:, e:] @ wo[1]], dim=1)
    bof = wf[:, :e] @ bo[0] + wf[:, e:] @ bo[1] + bf
    got = torch.cat(outs, dim=-1) @ wof.T + bof
    assert float((got - ref).abs().max()) < 1e-12


def test_match_len_modes():
    from lm2a_b200.sample import match_len
    a = np.arange(20, dtype=np.float32).reshape(10, 2)
    np.testing.assert_array_equal(match_len(a, 25, "interp"), orc.match_len_interp(a, 25))
    assert match_len(a, 10, "interp") is not None and match_len(a, 4).shape == (4, 2)
    np.testing.assert_array_equal(match_len(a, 12)[-1], a[-1])


def _shard_worker(rank, world, port, n_clips, batch, ret):
    import torch.distributed as dist
    from lm2a_b200 import distributed as ldist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = ldist.init_from_env("gloo")
    calls = []

    def fake_sampler(idx):  # "mel" of clip c is filled with c
        calls.append(list(idx))
        return torch.stack([torch.full((3, 5), float(c)) for c in idx])

    full = ldist.sample_sharded(n_clips, batch, fake_sampler, (3, 5), "cpu", r, w)
    ok = full.shape == (n_clips, 3, 5) and all(bool((full[c] == c).all()) for c in range(n_clips))
    ok = ok and sorted(sum(calls, [])) == ldist.shard_indices(n_clips, r, w)
    ok = ok and all(len(c) <= batch for c in calls)
    ret[rank] = ok
    dist.destroy_process_group()


def test_clip_sharding_world2_gloo():
    from lm2a_b200 import distributed as ldist
    assert ldist.shard_indices(7, 1, 2) == [1, 3, 5] and ldist.padded_shard_len(1868, 8) == 234
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, 7, 2, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
    assert ret.get(0) is True and ret.get(1) is True
