"""Per-kernel parity tests on a real B200: every C-ABI entry point against the matching
torch primitive (the op the reference dispatches to), on the same bf16-rounded operands so
only accumulation order and the final bf16 store differ. Tolerances are written per test."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lm2a_b200 import ops as _ops
    _ops.require_device(torch.zeros(1, device="cuda"))
    return _ops


def to_slab(x_nct, tp, ld=None, chan_off=0):
    """fp32 [R, C, T] -> bf16 slab [R*tp, ld] with zero pad slots / channels."""
    r, c, t = x_nct.shape
    ld = ld or c
    s = torch.zeros(r, tp, ld, dtype=BF16, device=x_nct.device)
    s[:, :t, chan_off:chan_off + c] = x_nct.permute(0, 2, 1).to(BF16)
    return s.view(r * tp, ld)


def from_slab(slab, r, tp, t, c, chan_off=0):
    return slab.view(r, tp, -1)[:, :t, chan_off:chan_off + c].permute(0, 2, 1).float()


def pads_are_zero(slab, r, tp, t):
    return bool((slab.view(r, tp, -1)[:, t:, :] == 0).all())


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randn(*shape, generator=g, device="cuda") * scale


def pack_w(w, n_pad=None, cin_pad=None):
    cout, cin, k = w.shape
    cin_pad = cin_pad or cin
    n_pad = n_pad or (cout + 127) // 128 * 128
    out = torch.zeros(n_pad, k, cin_pad, device=w.device)
    out[:cout, :, :cin] = w.permute(0, 2, 1)
    return out.reshape(n_pad, k * cin_pad).to(BF16).contiguous()


def pad_bias(b, n_pad):
    out = torch.zeros(n_pad, device=b.device)
    out[: b.numel()] = b
    return out


def bf(x):
    return x.to(BF16).float()


def assert_close(got, ref, rel=1e-2, what=""):
    err = (got - ref).norm() / ref.norm().clamp_min(1e-20)
    mx = (got - ref).abs().max()
    assert torch.isfinite(got).all(), what
    assert err < rel, f"{what}: rel-L2 {err:.3e} max-abs {mx:.3e}"


@pytest.mark.parametrize("cta_group", [1, 2, 0])  # one CTA per tile, CTA pair (cta_group::2), auto
@pytest.mark.parametrize("r,t,tp,cin,cout,block_n", [
    (2, 60, 64, 64, 128, 128),      # single M tile (pair: the second CTA's rows are all padding)
    (3, 129, 130, 128, 256, 256),   # tiles span clip boundaries
    (5, 258, 260, 256, 512, 0),     # multi-tile persistent, auto block_n
    (64, 64, 65, 1024, 1024, 256),  # production mid level: K = 3072, all SMs busy
    (40, 516, 520, 256, 256, 128),  # more tiles than SMs: several tiles per CTA / pair
])
def test_conv_k3(ops, r, t, tp, cin, cout, block_n, cta_group):
    x = rnd(r, cin, t, seed=1)
    w = rnd(cout, cin, 3, scale=1 / math.sqrt(3 * cin), seed=2)
    b = rnd(cout, scale=0.1, seed=3)
    xs = to_slab(x, tp)
    out = torch.full((r * tp, cout), 7.0, dtype=BF16, device="cuda")
    d = ops.make_conv_desc([ops.Seg(xs, cin, cin, ops.TAPS_K3, r * tp)], pack_w(w), pad_bias(b, (cout + 127) // 128 * 128),
                           cout, r * tp, tp, t, out, cout, block_n=block_n, cta_group=cta_group)
    ops.conv1d(d)
    torch.cuda.synchronize()
    ref = F.conv1d(bf(x), bf(w), b, padding=1)
    assert_close(from_slab(out, r, tp, t, cout), ref, 6e-3, "conv k3")
    assert pads_are_zero(out, r, tp, t)


def test_conv_k1_film_residual_and_strided_io(ops):
    """k=1 conv reading a channel-offset view, FiLM epilogue, residual add, strided output."""
    r, t, tp, cin, cout = 4, 100, 104, 128, 128
    x = rnd(r, cin, t, seed=4)
    res = rnd(r, cout, t, seed=5)
    w = rnd(cout, cin, 1, scale=1 / math.sqrt(cin), seed=6)
    b = rnd(cout, scale=0.1, seed=7)
    film = rnd(r, 2 * cout + 64, scale=0.5, seed=8)  # table with a column offset of 64
    xs = to_slab(x, tp, ld=2 * cin, chan_off=cin)     # x lives in the second half of a wider slab
    rs = to_slab(res, tp)
    out = torch.zeros(r * tp, 2 * cout, dtype=BF16, device="cuda")
    d = ops.make_conv_desc([ops.Seg(xs, 2 * cin, cin, ops.TAPS_K1, r * tp, chan_off=cin)], pack_w(w),
                           pad_bias(b, 128), cout, r * tp, tp, t, out, 2 * cout, out_chan_off=cout,
                           film=film, film_col=64, film_shift_off=cout, residual=rs, res_ld=cout)
    ops.conv1d(d)
    torch.cuda.synchronize()
    h = F.conv1d(bf(x), bf(w), b)
    sc, sh = film[:, 64:64 + cout], film[:, 64 + cout:64 + 2 * cout]
    ref = h * (1 + sc[:, :, None]) + sh[:, :, None] + bf(res)
    assert_close(from_slab(out, r, tp, t, cout, chan_off=cout), ref, 6e-3, "k1+film+res")
    assert bool((out[:, :cout] == 0).all()), "first half of the output slab must be untouched"


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("r,t_in,c", [(2, 64, 128), (3, 129, 256), (9, 516, 256)])
def test_conv_k4s2(ops, r, t_in, c, cta_group):
    t_out = t_in // 2
    tp_out = t_out + 1
    tp_in = 2 * tp_out
    x = rnd(r, c, t_in, seed=9)
    w = rnd(c, c, 4, scale=1 / math.sqrt(4 * c), seed=10)
    b = rnd(c, scale=0.1, seed=11)
    # input lives in the second half of a concat slab, exactly like the skip connection
    xs = to_slab(x, tp_in, ld=2 * c, chan_off=c)
    out = torch.zeros(r * tp_out, c, dtype=BF16, device="cuda")
    d = ops.make_conv_desc([ops.Seg(xs, 2 * c, c, ops.TAPS_K4S2, r * tp_in, chan_off=c)], pack_w(w),
                           pad_bias(b, (c + 127) // 128 * 128), c, r * tp_out, tp_out, t_out, out, c,
                           cta_group=cta_group)
    ops.conv1d(d)
    torch.cuda.synchronize()
    ref = F.conv1d(bf(x), bf(w), b, stride=2, padding=1)
    assert ref.shape[2] == t_out
    assert_close(from_slab(out, r, tp_out, t_out, c), ref, 6e-3, "conv k4s2")
    assert pads_are_zero(out, r, tp_out, t_out)


@pytest.mark.parametrize("cta_group", [1, 2])
def test_conv_two_segments_and_f32_nct(ops, cta_group):
    """conv2 + fused 1x1 skip conv (second K segment), and the fp32 [R, C, T] eps epilogue."""
    r, t, tp, c1, c2, cout = 3, 77, 80, 128, 256, 128
    a = rnd(r, c1, t, seed=12)
    x = rnd(r, c2, t, seed=13)
    w2 = rnd(cout, c1, 3, scale=1 / math.sqrt(3 * c1), seed=14)
    ws = rnd(cout, c2, 1, scale=1 / math.sqrt(c2), seed=15)
    b = rnd(cout, scale=0.1, seed=16)
    wcat = torch.cat([pack_w(w2), pack_w(ws)], dim=1).contiguous()
    out = torch.zeros(r * tp, cout, dtype=BF16, device="cuda")
    d = ops.make_conv_desc([ops.Seg(to_slab(a, tp), c1, c1, ops.TAPS_K3, r * tp),
                            ops.Seg(to_slab(x, tp), c2, c2, ops.TAPS_K1, r * tp)], wcat,
                           pad_bias(b, 128), cout, r * tp, tp, t, out, cout, cta_group=cta_group)
    ops.conv1d(d)
    ref = F.conv1d(bf(a), bf(w2), b, padding=1) + F.conv1d(bf(x), bf(ws))
    torch.cuda.synchronize()
    assert_close(from_slab(out, r, tp, t, cout), ref, 6e-3, "two segments")

    n_valid = 80
    wo = rnd(n_valid, c1, 1, scale=1 / math.sqrt(c1), seed=17)
    bo = rnd(n_valid, scale=0.1, seed=18)
    eps = torch.zeros(r, n_valid, t, device="cuda")
    d = ops.make_conv_desc([ops.Seg(to_slab(a, tp), c1, c1, ops.TAPS_K1, r * tp)], pack_w(wo, 128),
                           pad_bias(bo, 128), n_valid, r * tp, tp, t, eps, 0,
                           out_mode=ops.OUT_F32_NCT, block_n=128, cta_group=cta_group)
    ops.conv1d(d)
    torch.cuda.synchronize()
    assert_close(eps, F.conv1d(bf(a), bf(wo), bo), 2e-3, "f32 nct epilogue")


def test_conv_rejects_bad_arguments(ops):
    x = torch.zeros(64, 64, dtype=BF16, device="cuda")
    w = torch.zeros(128, 64, dtype=BF16, device="cuda")
    b = torch.zeros(128, device="cuda")
    out = torch.zeros(64, 128, dtype=BF16, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 64"):
        ops.conv1d(ops.make_conv_desc([ops.Seg(x, 64, 40, ops.TAPS_K1, 64)], w, b, 128, 64, 64, 64, out, 128))
    with pytest.raises(RuntimeError, match="slots"):
        ops.conv1d(ops.make_conv_desc([ops.Seg(x, 64, 64, ops.TAPS_K1, 32)], w, b, 128, 64, 64, 64, out, 128))


@pytest.mark.parametrize("r,t,tp,c,groups", [(2, 37, 40, 64, 8), (3, 129, 130, 2048, 8),
                                             (4, 516, 520, 256, 8), (2, 516, 520, 512, 8),
                                             (1, 2064, 2080, 256, 8), (2, 2064, 2080, 512, 8)])
def test_gn_silu(ops, r, t, tp, c, groups):
    x = rnd(r, c, t, seed=20) * 1.7 + 0.3
    gamma = 1 + 0.1 * rnd(c, seed=21)
    beta = 0.1 * rnd(c, seed=22)
    xs = to_slab(x, tp)
    y = torch.full((r * tp, c), 3.0, dtype=BF16, device="cuda")
    ops.gn_silu(xs, c, y, c, gamma, beta, r, tp, t, c, groups)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(bf(x), groups, gamma, beta, 1e-5))
    assert_close(from_slab(y, r, tp, t, c), ref, 5e-3, "gn+silu")
    assert pads_are_zero(y, r, tp, t)


@pytest.mark.parametrize("r,t,tp,cin,cout,groups,r0", [
    (3, 37, 40, 64, 64, 8, 0),        # tiny: 8 channels per group, several clips per warp segment
    (5, 129, 130, 128, 256, 8, 2),    # row sub-range (cond rows of a CFG batch), clip straddles
    (4, 516, 520, 256, 256, 8, 0),    # production level 0
    (6, 64, 65, 1024, 2048, 8, 3),    # 256 channels per group, 65-slot clips
])
@pytest.mark.parametrize("cta_group", [1, 2])
def test_conv_stats_feed_gn_apply(ops, r, t, tp, cin, cout, groups, r0, cta_group):
    """conv epilogue accumulates the exact GroupNorm sums; gn_apply consumes them: together they
    must equal F.group_norm + SiLU of the conv output (unet1d_ultimate.py:136-147)."""
    nr = r - r0
    x = rnd(r, cin, t, seed=50)
    w = rnd(cout, cin, 3, scale=1 / math.sqrt(3 * cin), seed=51)
    b = rnd(cout, scale=0.3, seed=52)
    gamma = 1 + 0.1 * rnd(cout, seed=53)
    beta = 0.1 * rnd(cout, seed=54)
    xs = to_slab(x, tp)
    h = torch.zeros(r * tp, cout, dtype=BF16, device="cuda")
    st = ops.Stats(r, cout, groups, "cuda")
    # launch over rows [r0, r): slabs and stats addressed through row-offset views
    d = ops.make_conv_desc([ops.Seg(xs, cin, cin, ops.TAPS_K3, nr * tp, chan_off=r0 * tp * cin)],
                           pack_w(w), pad_bias(b, (cout + 127) // 128 * 128), cout, nr * tp, tp, t,
                           h, cout, out_chan_off=r0 * tp * cout, stats=st.view(r0, 0),
                           cta_group=cta_group)
    ops.conv1d(d)
    y = torch.full((r * tp, cout), 3.0, dtype=BF16, device="cuda")
    ops.gn_apply(h, cout, y, cout, st.view(r0, 0), gamma, beta, nr, tp, t, cout, groups,
                 x_chan_off=r0 * tp * cout, y_chan_off=r0 * tp * cout)
    torch.cuda.synchronize()
    hf = from_slab(h, r, tp, t, cout)[r0:]
    ref = F.silu(F.group_norm(hf, groups, gamma, beta, 1e-5))
    assert_close(from_slab(y, r, tp, t, cout)[r0:], ref, 6e-3, "conv stats -> gn_apply")
    assert pads_are_zero(y[r0 * tp:], nr, tp, t)
    # the statistics themselves: per-(row, group) sums of the conv output (fp32 values before
    # the bf16 store; compared with the stored bf16 values)
    sums = st.sums()[r0:]
    assert bool((st.sums()[:r0] == 0).all()), "rows outside the launch must stay untouched"
    ref1 = hf.double().reshape(nr, groups, -1).sum(-1)
    ref2 = (hf.double().reshape(nr, groups, -1) ** 2).sum(-1)
    assert torch.allclose(sums[..., 0], ref1, rtol=2e-2, atol=0.5)
    assert torch.allclose(sums[..., 1], ref2, rtol=2e-2)
    # exactness: integer accumulation -> a second run (different atomic order) and any other
    # tile shape give bit-identical sums
    before = st.buf.clone()
    for bn, cg in ((128, 1), (128, 2)) + (((256, 1), (256, 2)) if cout % 256 == 0 else ()):
        st.zero_()
        d.block_n, d.cta_group = bn, cg
        ops.conv1d(d)
        torch.cuda.synchronize()
        assert torch.equal(before, st.buf), f"stats differ for tile {bn} x cta_group {cg}"


def _gn_input_case(ops, r, t, tp, cin, cout, groups, taps, r0, cta_group, block_n, skip_c=0,
                   silu=True):
    """y = conv(SiLU(GroupNorm(x))) [+ 1x1 skip conv of a raw second segment] with the
    normalisation applied inside the conv (operand transform) vs torch, and vs the same conv
    run on gn_apply's output (must be bit-identical: same arithmetic on the same operands)."""
    nr = r - r0
    k = 3 if taps == ops.TAPS_K3 else 1
    x = rnd(r, cin, t, seed=80) * 1.3 + 0.2
    w = rnd(cout, cin, k, scale=1 / math.sqrt(k * cin), seed=81)
    b = rnd(cout, scale=0.3, seed=82)
    gamma = 1 + 0.1 * rnd(cin, seed=83)
    beta = 0.1 * rnd(cin, seed=84)
    n_pad = (cout + 127) // 128 * 128
    # x and its exact sums come from a producer kernel (bias_add with a zero bias)
    xs = torch.zeros(r * tp, cin, dtype=BF16, device="cuda")
    st = ops.Stats(r, cin, groups, "cuda")
    ops.bias_add(to_slab(x, tp), cin, 0, xs, cin, 0, torch.zeros(cin, device="cuda"), r * tp, tp, t,
                 cin, st)
    segs = [ops.Seg(xs, cin, cin, taps, nr * tp, chan_off=r0 * tp * cin)]
    wk = pack_w(w)
    ref_extra = 0
    if skip_c:
        xsk = rnd(r, skip_c, t, seed=85)
        wsk = rnd(cout, skip_c, 1, scale=1 / math.sqrt(skip_c), seed=86)
        sks = to_slab(xsk, tp)
        segs.append(ops.Seg(sks, skip_c, skip_c, ops.TAPS_K1, nr * tp, chan_off=r0 * tp * skip_c))
        wk = torch.cat([wk, pack_w(wsk)], dim=1).contiguous()
        ref_extra = F.conv1d(bf(xsk), bf(wsk))[r0:]
    out = torch.full((r * tp, cout), 7.0, dtype=BF16, device="cuda")
    d = ops.make_conv_desc(segs, wk, pad_bias(b, n_pad), cout, nr * tp, tp, t, out, cout,
                           out_chan_off=r0 * tp * cout, block_n=block_n, cta_group=cta_group,
                           in_gn=(st.view(r0, 0), gamma, beta, 1e-5, silu))
    ops.conv1d(d)
    torch.cuda.synchronize()
    xf = from_slab(xs, r, tp, t, cin)
    n = F.group_norm(xf, groups, gamma, beta, 1e-5)
    n = F.silu(n) if silu else n
    ref = F.conv1d(bf(n), bf(w), b, padding=k // 2)[r0:] + ref_extra
    got = from_slab(out, r, tp, t, cout)[r0:]
    assert_close(got, ref, 8e-3, "conv with GroupNorm'd operand")
    assert pads_are_zero(out[r0 * tp:], nr, tp, t)
    assert bool((out[: r0 * tp] == 7.0).all()), "rows before the launch range were touched"
    # reference path: stand-alone gn_apply -> plain conv
    norm = torch.zeros(r * tp, cin, dtype=BF16, device="cuda")
    ops.gn_apply(xs, cin, norm, cin, st, gamma, beta, r, tp, t, cin, groups, silu=silu)
    segs2 = [ops.Seg(norm, cin, cin, taps, nr * tp, chan_off=r0 * tp * cin)] + segs[1:]
    out2 = torch.full((r * tp, cout), 7.0, dtype=BF16, device="cuda")
    d2 = ops.make_conv_desc(segs2, wk, pad_bias(b, n_pad), cout, nr * tp, tp, t, out2, cout,
                            out_chan_off=r0 * tp * cout, block_n=block_n, cta_group=cta_group,
                            k_order=1)
    ops.conv1d(d2)
    torch.cuda.synchronize()
    assert torch.equal(out, out2), "operand transform differs from gn_apply + conv"


@pytest.mark.parametrize("block_n,cta_group", [(0, 0), (256, 1), (256, 2), (128, 1), (128, 2)])
@pytest.mark.parametrize("r,t,tp,cin,cout,groups,r0", [
    (4, 516, 520, 256, 256, 8, 0),       # production level 0: tiles inside one clip
    (64, 516, 520, 256, 256, 8, 32),     # production conv1, cond rows of a CFG batch
    (32, 64, 65, 1024, 1024, 8, 0),      # mid level: every tile touches 3 clips, K = 3072
    (5, 129, 130, 128, 512, 8, 2),       # 16-channel groups: four groups per 64-channel block
])
def test_conv_k3_groupnorm_operand(ops, r, t, tp, cin, cout, groups, r0, block_n, cta_group):
    _gn_input_case(ops, r, t, tp, cin, cout, groups, ops.TAPS_K3, r0, cta_group, block_n)


@pytest.mark.parametrize("cta_group", [1, 2])
def test_conv_groupnorm_operand_variants(ops, cta_group):
    # k1 (out_proj: GroupNorm + SiLU + 1x1 conv), 80 real output channels would need the fp32
    # epilogue: here a 128-channel slab
    _gn_input_case(ops, 3, 100, 104, 256, 128, 8, ops.TAPS_K1, 0, cta_group, 128)
    # conv2 + raw 1x1 skip segment (only the first segment is normalised)
    _gn_input_case(ops, 3, 129, 130, 256, 128, 8, ops.TAPS_K3, 1, cta_group, 128, skip_c=128)
    # very short clips: a tile touches 11 of them (12-slot pitch)
    _gn_input_case(ops, 40, 11, 12, 64, 128, 8, ops.TAPS_K3, 0, cta_group, 128)
    # legacy concat width 1536 (192-channel groups straddle 64-channel blocks)
    _gn_input_case(ops, 2, 64, 65, 1536, 128, 8, ops.TAPS_K3, 0, cta_group, 128)
    # the largest supported channel count
    _gn_input_case(ops, 2, 130, 132, 2048, 256, 8, ops.TAPS_K3, 0, cta_group, 256)


@pytest.mark.parametrize("block_n,cta_group", [(0, 0), (256, 1), (128, 2)])
@pytest.mark.parametrize("r,t_in,cin,cout", [(3, 129, 256, 128), (64, 64, 1024, 1024),
                                              (5, 258, 512, 256), (2, 20, 64, 128)])
def test_conv_fused_upsampling(ops, r, t_in, cin, cout, block_n, cta_group):
    """UpSampleConv (unet1d_ultimate.py:210-239): x2 linear interpolation (align_corners) fused
    into the k3 conv's operand path == upsample2x kernel followed by the plain conv, bit for bit,
    and == F.interpolate + F.conv1d within bf16 tolerance. The output slab is one slot longer
    than 2 * t_in per clip (the skip's length): those slots must come out as conv-of-zero-pad."""
    if block_n and ((cout + 127) // 128 * 128) % block_n:
        pytest.skip("block_n does not divide n_pad")
    tp_in = t_in + 1
    tp_out = 2 * tp_in
    t_up = 2 * t_in
    x = rnd(r, cin, t_in, seed=90)
    w = rnd(cout, cin, 3, scale=1 / math.sqrt(3 * cin), seed=91)
    b = rnd(cout, scale=0.2, seed=92)
    xs = to_slab(x, tp_in)
    n_pad = (cout + 127) // 128 * 128
    st = ops.Stats(r, cout, 8, "cuda")
    out = torch.full((r * tp_out, cout), 7.0, dtype=BF16, device="cuda")
    d = ops.make_conv_desc([ops.Seg(xs, cin, cin, ops.TAPS_K3, r * tp_in)], pack_w(w),
                           pad_bias(b, n_pad), cout, r * tp_out, tp_out, t_up, out, cout, stats=st,
                           block_n=block_n, cta_group=cta_group, up2x=(tp_in, t_in))
    ops.conv1d(d)
    # unfused: upsample kernel -> plain conv
    up = torch.zeros(r * tp_out, cin, dtype=BF16, device="cuda")
    ops.upsample2x(xs, cin, up, cin, r, tp_in, t_in, tp_out, cin)
    st2 = ops.Stats(r, cout, 8, "cuda")
    out2 = torch.full((r * tp_out, cout), 7.0, dtype=BF16, device="cuda")
    d2 = ops.make_conv_desc([ops.Seg(up, cin, cin, ops.TAPS_K3, r * tp_out)], pack_w(w),
                            pad_bias(b, n_pad), cout, r * tp_out, tp_out, t_up, out2, cout,
                            stats=st2, block_n=block_n, cta_group=cta_group, k_order=1)
    ops.conv1d(d2)
    torch.cuda.synchronize()
    assert torch.equal(out, out2), "fused upsampling differs from upsample2x + conv"
    assert torch.equal(st.buf, st2.buf)
    ref = F.conv1d(bf(F.interpolate(bf(x), scale_factor=2, mode="linear", align_corners=True)),
                   bf(w), b, padding=1)
    assert_close(from_slab(out, r, tp_out, t_up, cout), ref, 8e-3, "fused upsample + conv")
    assert pads_are_zero(out, r, tp_out, t_up)


def test_conv_groupnorm_operand_rejects_unsupported(ops):
    x = torch.zeros(64 * 2, 64, dtype=BF16, device="cuda")
    w = torch.zeros(128, 192, dtype=BF16, device="cuda")
    b = torch.zeros(128, device="cuda")
    g = torch.ones(64, device="cuda")
    out = torch.zeros(128, 128, dtype=BF16, device="cuda")
    st = ops.Stats(64, 64, 8, "cuda")
    assert not ops.in_gn_supported(2, 8) and ops.in_gn_supported(65, 8)
    with pytest.raises(RuntimeError, match="clip-rows"):   # 2-slot clips: 66 clips per tile
        ops.conv1d(ops.make_conv_desc([ops.Seg(x, 64, 64, ops.TAPS_K3, 128)], w, b, 128, 128, 2, 1,
                                      out, 128, in_gn=(st, g, g, 1e-5, True)))
    st2 = ops.Stats(2, 64, 8, "cuda")
    with pytest.raises(RuntimeError, match="SiLU"):        # GroupNorm alone: gn_apply's job
        ops.conv1d(ops.make_conv_desc([ops.Seg(x, 64, 64, ops.TAPS_K3, 128)], w, b, 128, 128, 64, 60,
                                      out, 128, in_gn=(st2, g, g, 1e-5, False)))


@pytest.mark.parametrize("r,t,tp,c,groups", [(3, 129, 130, 512, 8), (33, 64, 65, 1024, 8),
                                             (5, 9, 10, 64, 8), (4, 516, 520, 256, 8)])
def test_bias_add_stats(ops, r, t, tp, c, groups):
    x = rnd(r, c, t, seed=60)
    bias = rnd(c, scale=0.5, seed=61)
    gamma = 1 + 0.1 * rnd(c, seed=62)
    beta = 0.1 * rnd(c, seed=63)
    xs = to_slab(x, tp)
    y = torch.full((r * tp, c), 2.0, dtype=BF16, device="cuda")
    st = ops.Stats(r, c, groups, "cuda")
    ops.bias_add(xs, c, 0, y, c, 0, bias, r * tp, tp, t, c, st)
    z = torch.zeros_like(y)
    ops.gn_apply(y, c, z, c, st, gamma, beta, r, tp, t, c, groups)
    torch.cuda.synchronize()
    yf = from_slab(y, r, tp, t, c)
    assert_close(yf, bf(x) + bias[None, :, None], 4e-3, "bias_add")
    assert pads_are_zero(y, r, tp, t)
    ref = F.silu(F.group_norm(yf, groups, gamma, beta, 1e-5))
    assert_close(from_slab(z, r, tp, t, c), ref, 6e-3, "bias_add stats -> gn_apply")
    # exact sums do not depend on where the rows sit: the same clips in another order
    perm = torch.randperm(r, device="cuda")
    st2 = ops.Stats(r, c, groups, "cuda")
    xs2 = to_slab(x[perm], tp)
    ops.bias_add(xs2, c, 0, torch.empty_like(y), c, 0, bias, r * tp, tp, t, c, st2)
    torch.cuda.synchronize()
    assert torch.equal(st2.buf.view(r, groups, 2), st.buf.view(r, groups, 2)[perm])


@pytest.mark.parametrize("e,heads,t,lk,qgain", [
    (256, 8, 100, 77, 1.0), (512, 8, 258, 516, 1.0), (1024, 8, 64, 516, 1.0), (128, 4, 33, 64, 1.0),
    (256, 8, 516, 516, 1.0),     # production level 0: dh = 32, five query tiles, ragged key tail
    (1024, 8, 129, 516, 1.0),    # production level 2: dh = 128, two query tiles
    (512, 8, 130, 2064, 1.0),    # long-clip K/V
    (256, 8, 200, 300, 6.0), (1024, 8, 70, 200, 6.0),  # peaked softmax: running max grows, O is rescaled
    # legacy UNet1D head dims (4 heads over 384 / 768 / 1024 / 1536 channels): 96 (32-channel
    # panels), 192, 256, 384 (two V^T boxes, one 512-column TMEM allocation, single-stage rings)
    (384, 4, 50, 60, 1.0), (768, 4, 129, 150, 1.0), (1024, 4, 64, 516, 1.0),
    (1536, 4, 129, 516, 1.0), (1536, 4, 40, 100, 6.0),
    # tail rows taken over by the producer warps of the full tiles: two per CTA (dh = 32, 64B
    # swizzle), dh = 96 panels, a peaked softmax on the tail path, rem > 2 * nfull (own tile)
    (128, 4, 260, 64, 1.0), (384, 4, 129, 60, 1.0), (512, 8, 258, 300, 6.0), (256, 8, 261, 100, 1.0),
])
def test_cross_attention_core(ops, e, heads, t, lk, qgain):
    _attention_core_case(ops, e, heads, t, lk, qgain)


@pytest.mark.parametrize("e,heads,t,lk,qgain", [
    (256, 8, 516, 516, 1.0), (512, 8, 258, 516, 1.0), (256, 8, 100, 77, 1.0), (128, 4, 33, 64, 1.0),
    (256, 8, 200, 300, 6.0), (512, 8, 258, 300, 6.0), (128, 4, 260, 64, 1.0), (256, 8, 261, 100, 1.0),
    (512, 8, 130, 1100, 1.0),
])
def test_cross_attention_core_resident_kernel(ops, monkeypatch, e, heads, t, lk, qgain):
    """The opt-in resident-K/V kernel (attention_res.cu, LM2A_ATTN_RESIDENT=1) for d_h = 32 / 64:
    full + tail tiles split over CTAs, two tile slots, P through tensor memory."""
    monkeypatch.setenv("LM2A_ATTN_RESIDENT", "1")
    from lm2a_b200 import ops as _o
    before = _o.launch_count()
    _attention_core_case(ops, e, heads, t, lk, qgain)
    assert _o.launch_count() > before


def _attention_core_case(ops, e, heads, t, lk, qgain):
    r, slots, tp = 3, 2, t + 2
    dh = e // heads
    q = rnd(r, 2 * e, t, seed=30) * qgain
    kv_m = rnd(slots * lk, 2 * e, seed=31).to(BF16)
    kv_t = rnd(slots * lk, 2 * e, seed=32).to(BF16)
    kv_slot = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
    scale = 1.0 / math.sqrt(dh)
    qs = to_slab(q * (scale * 1.4426950408889634), tp)
    o = torch.zeros(r * tp, 2 * e, dtype=BF16, device="cuda")
    lk_pad = (lk + 7) // 8 * 8
    vts = []
    for kv in (kv_m, kv_t):
        vt = torch.zeros(slots * e, lk_pad, dtype=BF16, device="cuda")
        ops.transpose_kv(kv, 2 * e, e, vt, lk_pad, slots, lk, e)
        torch.cuda.synchronize()
        want = kv.view(slots, lk, 2 * e)[:, :, e:].permute(0, 2, 1).reshape(slots * e, lk)
        assert torch.equal(vt[:, :lk], want) and bool((vt[:, lk:] == 0).all())
        vts.append(vt)
    ops.cross_attn(qs, 2 * e, o, 2 * e, ops._ptr(kv_m), ops._ptr(vts[0]), ops._ptr(kv_t),
                   ops._ptr(vts[1]), 2 * e, lk_pad, kv_slot, slots, r, tp, t, lk, e, heads)
    torch.cuda.synchronize()
    got = from_slab(o, r, tp, t, 2 * e)
    for s, kv in enumerate((kv_m, kv_t)):
        kvf = kv.float().view(slots, lk, 2 * e)[kv_slot.long()]
        k = kvf[:, :, :e].view(r, lk, heads, dh).transpose(1, 2)
        v = kvf[:, :, e:].view(r, lk, heads, dh).transpose(1, 2)
        qq = qs.float().view(r, tp, 2 * e)[:, :t, s * e:(s + 1) * e] / 1.4426950408889634
        qq = qq.reshape(r, t, heads, dh).transpose(1, 2)
        p = torch.softmax(qq @ k.transpose(-1, -2), dim=-1)
        ref = (p @ v).transpose(1, 2).reshape(r, t, e).permute(0, 2, 1)
        assert_close(got[:, s * e:(s + 1) * e], ref, 1e-2, f"attention stream {s}")


@pytest.mark.parametrize("heads,t,lk,qgain", [
    (8, 64, 516, 1.0),      # production level 3: two heads per 128-row tile
    (8, 129, 516, 1.0),     # production level 2: 8 full tiles + one tile of the 8 leftover rows
    (8, 258, 300, 1.0),     # two full tiles per head, 2-row tails packed 8 heads to a tile
    (4, 33, 64, 1.0),       # 64-row sub-blocks, two heads per tile; one key chunk
    (3, 100, 77, 1.0),      # odd head count: a tail tile with an unused sub-block
    (8, 70, 200, 6.0),      # peaked softmax: the running max grows, O is rescaled
    (2, 516, 640, 1.0),     # four full tiles per head, the largest resident key count
    (8, 17, 16, 1.0),       # a single 16-key chunk; 32-row sub-blocks
])
@pytest.mark.parametrize("split", ["0", "2"])
def test_cross_attention_cond(ops, monkeypatch, heads, t, lk, qgain, split):
    """lm2a_cross_attn_cond_bf16: every head attends to the raw condition sequence itself,
    softmax(q'_h C^T) C (head dim = condition width = 128), vs fp32 torch. split: the tail tiles
    (T mod 128 rows) in a CTA of their own (LM2A_ATTN_SPLIT_TAIL: never / whenever possible; the
    default is a cost model)."""
    monkeypatch.setenv("LM2A_ATTN_SPLIT_TAIL", split)
    r, slots, tp, dh = 3, 2, t + 2, 128
    e = heads * dh
    assert ops.cond_attn_supported(lk) and not ops.cond_attn_supported(700)
    q = rnd(r, 2 * e, t, seed=33) * qgain
    cm = rnd(slots * lk, dh, seed=34).to(BF16)
    ct = rnd(slots * lk, dh, seed=35).to(BF16)
    kv_slot = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
    scale = 1.0 / math.sqrt(dh)
    qs = to_slab(q * (scale * 1.4426950408889634), tp)
    for n_streams in (2, 1):
        o = torch.full((r * tp, 2 * e), 3.0, dtype=BF16, device="cuda")
        ops.cross_attn_cond(qs, 2 * e, o, 2 * e, ops._ptr(cm), ops._ptr(ct), dh, kv_slot, slots, r,
                            tp, t, lk, heads, n_streams)
        torch.cuda.synchronize()
        got = from_slab(o, r, tp, t, 2 * e)
        for s, c in enumerate((cm, ct)[:n_streams]):
            cf = c.float().view(slots, lk, dh)[kv_slot.long()]            # [r, lk, dh]
            qq = qs.float().view(r, tp, 2 * e)[:, :t, s * e:(s + 1) * e] / 1.4426950408889634
            qq = qq.reshape(r, t, heads, dh).transpose(1, 2)              # [r, h, t, dh]
            p = torch.softmax(qq @ cf[:, None].transpose(-1, -2), dim=-1)
            ref = (p @ cf[:, None]).transpose(1, 2).reshape(r, t, e).permute(0, 2, 1)
            assert_close(got[:, s * e:(s + 1) * e], ref, 1e-2, f"cond attention stream {s}")
        if n_streams == 1:
            assert bool((o.view(r, tp, 2 * e)[:, :, e:] == 3.0).all()), "stream 1 was touched"


@pytest.mark.parametrize("e,heads,t,lk,qgain,cond", [
    (256, 8, 516, 516, 1.0, False),    # production level 0: 4 leftover rows, d_h = 32
    (512, 8, 258, 516, 1.0, False),    # level 1: 2 rows, d_h = 64
    (1024, 8, 129, 516, 1.0, True),    # level 2 on the condition slab: 1 row
    (1024, 8, 129, 516, 1.0, False),   # the same block with per-head K / V (d_h = 128)
    (256, 8, 136, 77, 6.0, False),     # 8 rows, odd Lk, peaked softmax
    (512, 4, 261, 300, 1.0, True),     # condition mode, 5 rows: one head per CTA
    (384, 3, 131, 100, 1.0, True),     # condition mode, 3 rows x 2 heads per CTA, odd head count
])
def test_cross_attention_tail_rows(ops, e, heads, t, lk, qgain, cond):
    """lm2a_cross_attn_tail_bf16 (the T mod 128 query rows on the CUDA cores) next to the
    tensor-core launch over the whole tiles: together they cover every row; vs fp32 torch."""
    r, slots, tp = 3, 2, t + 2
    dh = e // heads
    t0, n_tail = t // 128 * 128, t % 128
    q = rnd(r, 2 * e, t, seed=36) * qgain
    kv_slot = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
    qs = to_slab(q * (1.4426950408889634 / math.sqrt(dh)), tp)
    lk_pad = (lk + 7) // 8 * 8
    if cond:
        assert dh == 128
        src = [rnd(slots * lk, dh, seed=37 + i).to(BF16) for i in range(2)]
    else:
        src = [rnd(slots * lk, 2 * e, seed=37 + i).to(BF16) for i in range(2)]
        vts = []
        for kv in src:
            vt = torch.zeros(slots * e, lk_pad, dtype=BF16, device="cuda")
            ops.transpose_kv(kv, 2 * e, e, vt, lk_pad, slots, lk, e)
            vts.append(vt)
    for n_streams in (2, 1):
        o = torch.full((r * tp, 2 * e), 3.0, dtype=BF16, device="cuda")
        if cond:
            ops.cross_attn_cond(qs, 2 * e, o, 2 * e, ops._ptr(src[0]), ops._ptr(src[1]), dh, kv_slot,
                                slots, r, tp, t0, lk, heads, n_streams)
            ops.cross_attn_tail(qs, 2 * e, o, 2 * e, ops._ptr(src[0]), ops._ptr(src[0]),
                                ops._ptr(src[1]), ops._ptr(src[1]), dh, dh, kv_slot, slots, r, tp,
                                t0, n_tail, lk, e, heads, n_streams, True)
        else:
            ops.cross_attn(qs, 2 * e, o, 2 * e, ops._ptr(src[0]), ops._ptr(vts[0]), ops._ptr(src[1]),
                           ops._ptr(vts[1]), 2 * e, lk_pad, kv_slot, slots, r, tp, t0, lk, e, heads,
                           n_streams)
            # keys and values row-major: the K and V halves of the projection output
            ops.cross_attn_tail(qs, 2 * e, o, 2 * e, ops._ptr(src[0]), ops._ptr(src[0], e),
                                ops._ptr(src[1]), ops._ptr(src[1], e), 2 * e, 2 * e, kv_slot, slots,
                                r, tp, t0, n_tail, lk, e, heads, n_streams, False)
        torch.cuda.synchronize()
        got = from_slab(o, r, tp, t, 2 * e)
        for s in range(n_streams):
            if cond:
                cf = src[s].float().view(slots, lk, dh)[kv_slot.long()]
                k = v = cf[:, None]
            else:
                kvf = src[s].float().view(slots, lk, 2 * e)[kv_slot.long()]
                k = kvf[:, :, :e].view(r, lk, heads, dh).transpose(1, 2)
                v = kvf[:, :, e:].view(r, lk, heads, dh).transpose(1, 2)
            qq = qs.float().view(r, tp, 2 * e)[:, :t, s * e:(s + 1) * e] / 1.4426950408889634
            qq = qq.reshape(r, t, heads, dh).transpose(1, 2)
            p = torch.softmax(qq @ k.transpose(-1, -2), dim=-1)
            ref = (p @ v).transpose(1, 2).reshape(r, t, e).permute(0, 2, 1)
            assert_close(got[:, s * e:(s + 1) * e, t0:], ref[:, :, t0:], 6e-3, f"tail rows, stream {s}")
            assert_close(got[:, s * e:(s + 1) * e], ref, 1e-2, f"tiles + tail rows, stream {s}")
        ov = o.view(r, tp, 2 * e)
        assert bool((ov[:, t:, :] == 3.0).all()), "rows past T were touched"
        if n_streams == 1:
            assert bool((ov[:, :, e:] == 3.0).all()), "stream 1 was touched"
    with pytest.raises(RuntimeError, match="bad geometry"):
        ops.cross_attn_tail(qs, 2 * e, o, 2 * e, ops._ptr(src[0]), ops._ptr(src[0]), ops._ptr(src[1]),
                            ops._ptr(src[1]), dh if cond else 2 * e, dh if cond else 2 * e, kv_slot,
                            slots, r, tp, t0, 9, lk, e, heads, 2, cond)


@pytest.mark.parametrize("r,t,tp,c,groups", [(2, 129, 130, 1536, 8), (3, 50, 52, 768, 8),
                                               (2, 33, 36, 384, 8), (2, 70, 72, 192, 8)])
def test_gn_apply_channel_counts_off_the_cta_grid(ops, r, t, tp, c, groups):
    """c / 8 that neither divides nor is a multiple of the 256-thread CTA (legacy UNet1D concat
    widths 1536 / 768 / 384 / 192): statistics from bias_add, then the streaming apply."""
    x = rnd(r, c, t, seed=60)
    xs = to_slab(x, tp)
    bias = rnd(c, seed=61, scale=0.1)
    gamma, beta = 1.0 + 0.1 * rnd(c, seed=62), 0.1 * rnd(c, seed=63)
    st = ops.Stats(r, c, groups, "cuda")
    y = torch.zeros_like(xs)
    z = torch.zeros_like(xs)
    ops.bias_add(xs, c, 0, y, c, 0, bias, r * tp, tp, t, c, st)
    ops.gn_apply(y, c, z, c, st, gamma, beta, r, tp, t, c, groups)
    torch.cuda.synchronize()
    yf = from_slab(y, r, tp, t, c)
    ref = F.silu(F.group_norm(yf, groups, gamma, beta, 1e-5))
    assert_close(from_slab(z, r, tp, t, c), ref, 6e-3, "gn_apply")
    assert pads_are_zero(z, r, tp, t)


@pytest.mark.parametrize("r,t_lo,cin,cout,skip", [(2, 64, 128, 128, 64), (3, 129, 256, 128, 128),
                                                    (2, 16, 192, 64, 64)])
def test_conv_transpose_k4s2_as_two_k3_gemms(ops, r, t_lo, cin, cout, skip):
    """ConvTranspose1d k4 s2 p1 (legacy models/unet1d.py:105) = an even-slot and an odd-slot k3
    GEMM writing one row pair of the [M_lo, 2 * ld] view of the level-above concat slab; both
    launches add into the same exact sums, which feed gn_apply."""
    from lm2a_b200.engine import _convT_w, _finish
    tp_lo, tp_hi = t_lo + 1, 2 * (t_lo + 1)
    t_hi = 2 * t_lo + 1          # the skip is one frame longer (129 vs 2 * 64): F.pad path
    ld = cout + skip
    x = rnd(r, cin, t_lo, seed=70)
    w = rnd(cin, cout, 4, seed=71, scale=1.0 / math.sqrt(2 * cin))
    b = rnd(cout, seed=72, scale=0.1)
    xs = to_slab(x, tp_lo)
    even, odd = _convT_w(w)
    we, bias = _finish(even, b.double(), "cuda")
    wo, _ = _finish(odd, b.double(), "cuda")
    cat = torch.zeros(r * tp_hi, ld, dtype=BF16, device="cuda")
    cat[:, cout:] = 1.0   # the skip half must survive untouched
    st = ops.Stats(r, cout, 8, "cuda")
    for half, wt in enumerate((we, wo)):
        d = ops.make_conv_desc([ops.Seg(xs, cin, cin, ops.TAPS_K3, r * tp_lo)], wt, bias, cout,
                               r * tp_lo, tp_lo, t_lo, cat, 2 * ld, out_chan_off=half * ld,
                               stats=st)
        ops.conv1d(d)
    torch.cuda.synchronize()
    ref = F.conv_transpose1d(bf(x), bf(w), b, stride=2, padding=1)       # [r, cout, 2 * t_lo]
    got = from_slab(cat, r, tp_hi, 2 * t_lo, cout)
    assert_close(got, ref, 1e-2, "conv transpose")
    v = cat.view(r, tp_hi, ld)
    assert bool((v[:, 2 * t_lo:, :cout] == 0).all()), "slots past 2*T_lo must be zero (F.pad)"
    assert bool((v[:, :, cout:] == 1.0).all()), "skip half overwritten"
    # statistics of the h half: GroupNorm over [cout] channels of the written slab
    gamma, beta = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    z = torch.zeros(r * tp_hi, cout, dtype=BF16, device="cuda")
    ops.gn_apply(cat, ld, z, cout, st, gamma, beta, r, tp_hi, t_hi, cout, 8, silu=False)
    torch.cuda.synchronize()
    hpad = F.pad(got, (0, 1))
    refn = F.group_norm(hpad, 8, gamma, beta, 1e-5)
    assert_close(from_slab(z, r, tp_hi, t_hi, cout), refn, 6e-3, "convT stats -> gn_apply")


@pytest.mark.parametrize("t_in,t_out,c", [(180, 516, 234), (516, 516, 768), (720, 2064, 234),
                                          (1, 40, 8), (600, 77, 130), (2, 3, 5)])
def test_resample_seq_is_bit_exact_vs_numpy_interp(ops, t_in, t_out, c):
    """lm2a_resample_seq == the reference's interpolate_seq (np.interp per feature on
    linspace(0, L-1, T), float32 result): bit for bit, including ragged batches."""
    import numpy as np
    import lm2a_oracle as orc
    rng = np.random.default_rng(t_in * 1000 + t_out)
    rows = 3
    lens = [t_in, max(1, t_in // 2), max(1, t_in - 1)]
    x = np.zeros((rows, t_in, c), np.float32)
    for r_, ln in enumerate(lens):
        x[r_, :ln] = rng.normal(0, 3.0, size=(ln, c)).astype(np.float32)
    want = np.stack([orc.match_len_interp(x[r_, :ln], t_out) for r_, ln in enumerate(lens)])
    xd = torch.from_numpy(x).cuda()
    ld = (c + 63) // 64 * 64
    tp = t_out + 2
    out = torch.full((rows, t_out, c), float("nan"), device="cuda")
    slab = torch.full((rows * tp, ld), float("nan"), dtype=BF16, device="cuda")
    ops.resample_seq(xd, torch.tensor(lens, dtype=torch.int32, device="cuda"), out, slab, rows,
                     t_in, c, t_out, tp, ld)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out.cpu().numpy(), want.reshape(rows, t_out, c))
    sv = slab.view(rows, tp, ld)
    assert torch.equal(sv[:, :t_out, :c], out.to(BF16))
    assert bool((sv[:, t_out:, :] == 0).all()) and bool((sv[:, :, c:] == 0).all())


def test_loop_side_entry_points_reject_bad_arguments(ops):
    """Argument violations return an error BEFORE any launch (C ABI convention)."""
    x = torch.zeros(2, 10, 8, device="cuda")
    with pytest.raises(RuntimeError, match="resample_seq"):
        ops.resample_seq(x, None, None, None, 2, 10, 8, 20, 20, 8)          # no output at all
    with pytest.raises(RuntimeError, match="resample_seq"):
        ops.resample_seq(x, None, torch.zeros(2, 20, 8, device="cuda"), None, 2, 10, 8, 20, 19, 8)
    with pytest.raises(RuntimeError, match="mel_metrics"):
        ops.mel_metrics(torch.zeros(1, 80, 8, device="cuda"), torch.zeros(1, 80, 8, device="cuda"),
                        torch.zeros(1, 8, dtype=torch.float64, device="cuda"), 1, 80, 8)
    with pytest.raises(RuntimeError, match="cfg_ddim"):
        ops.cfg_ddim(torch.zeros(1, 80, 7, device="cuda"), torch.zeros(1, 80, 7, device="cuda"), None,
                     torch.zeros(8, device="cuda"), None, torch.zeros(1, dtype=torch.int32, device="cuda"),
                     None, None, 1, 80 * 7 + 1, 1.0, False, False)
    q = torch.zeros(4 * 66, 256, dtype=BF16, device="cuda")
    with pytest.raises(RuntimeError, match="n_streams"):
        ops.cross_attn(q, 256, q, 256, ops._ptr(q), ops._ptr(q), ops._ptr(q), ops._ptr(q), 256, 64,
                       torch.zeros(4, dtype=torch.int32, device="cuda"), 1, 4, 66, 64, 64, 128, 4, 3)


def test_torch_custom_op_layer(ops):
    """torch.ops.lm2a.* dispatch to the same C-ABI kernels as lm2a_b200.ops."""
    import numpy as np
    import lm2a_oracle as orc
    import lm2a_b200.torch_ops  # noqa: F401
    x = rnd(2, 37, 10, seed=90)
    got = torch.ops.lm2a.resample_seq(x, None, 64)
    want = np.stack([orc.match_len_interp(x[i].cpu().numpy(), 64) for i in range(2)])
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # in-place posterior update through the op == through ops.cfg_posterior
    xa, eps, nz = rnd(2, 80, 64, seed=91), rnd(4, 80, 64, seed=92), rnd(2, 80, 64, seed=93)
    sched = torch.rand(10, 4, device="cuda")
    xb = xa.clone()
    t1 = torch.full((2,), 7, dtype=torch.int64, device="cuda")
    t2 = t1.clone()
    torch.ops.lm2a.cfg_posterior(xa, eps, nz, sched, t1, None, 2.1, True, False)
    ops.cfg_posterior(xb, eps, nz, sched, t2, None, 2, 80 * 64, 2.1, True, False)
    assert torch.equal(xa, xb)
    m = torch.ops.lm2a.mel_metrics(rnd(1, 80, 40, seed=94), rnd(1, 80, 40, seed=95), 1.0, 0.0)
    assert m.shape == (1, 8) and torch.isfinite(m).all()


def test_torch_custom_ops_conv_attention_embedding(ops):
    """The tensor-in / tensor-out custom ops over the tensor-core kernels: torch.ops.lm2a.conv1d,
    cross_attn (+ transpose_kv), cross_attn_cond, cross_attn_tail, time_mlp, film, philox_normal
    against plain torch on the same bf16-rounded operands."""
    import lm2a_b200.torch_ops  # noqa: F401
    L = torch.ops.lm2a
    # conv k3 p1 + bias + residual (nn.Conv1d, unet1d_ultimate.py:87-88 + the :159 residual)
    r, t, tp, cin, cout = 3, 70, 72, 64, 128
    x = rnd(r, cin, t, seed=70)
    w = rnd(cout, cin, 3, seed=71) / math.sqrt(3 * cin)
    b = rnd(cout, seed=72, scale=0.1)
    res = rnd(r, cout, t, seed=73)
    wp = w.permute(0, 2, 1).reshape(cout, 3 * cin).to(BF16).contiguous()   # K = (tap, channel)
    got = L.conv1d(to_slab(x, tp), wp, b, cout, tp, t, ops.TAPS_K3, to_slab(res, tp))
    ref = F.conv1d(bf(x), bf(w), b, padding=1) + bf(res)
    assert_close(from_slab(got, r, tp, t, cout), ref, 6e-3, "torch.ops.lm2a.conv1d")
    assert pads_are_zero(got, r, tp, t)
    # attention: per-head K / V from the projection output, on the condition slab, and tail rows
    e, heads, lk, slots, t = 256, 8, 77, 2, 132
    tp, dh = t + 2, 32
    kv_slot = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
    q = to_slab(rnd(r, 2 * e, t, seed=74) * (1.4426950408889634 / math.sqrt(dh)), tp)
    kvs = [rnd(slots * lk, 2 * e, seed=75 + i).to(BF16) for i in range(2)]
    o = L.cross_attn(q, kvs[0], kvs[1], kv_slot, tp, t, lk, heads, 2)
    vt = L.transpose_kv(kvs[0], slots, lk, e)
    assert torch.equal(vt[:, :lk], kvs[0].view(slots, lk, 2 * e)[:, :, e:].permute(0, 2, 1).reshape(slots * e, lk))

    def attn_ref(qs, k, v, s, width, hd):
        qq = qs.float().view(r, tp, -1)[:, :t, s * width:(s + 1) * width] / 1.4426950408889634
        qq = qq.reshape(r, t, -1, hd).transpose(1, 2)
        p = torch.softmax(qq @ k.transpose(-1, -2), dim=-1)
        return (p @ v).transpose(1, 2).reshape(r, t, width).permute(0, 2, 1)

    got = from_slab(o, r, tp, t, 2 * e)
    for s_, kv in enumerate(kvs):
        kvf = kv.float().view(slots, lk, 2 * e)[kv_slot.long()]
        k = kvf[:, :, :e].view(r, lk, heads, dh).transpose(1, 2)
        v = kvf[:, :, e:].view(r, lk, heads, dh).transpose(1, 2)
        assert_close(got[:, s_ * e:(s_ + 1) * e], attn_ref(q, k, v, s_, e, dh), 1e-2, "ops cross_attn")
    # tail rows (4 of 132) recomputed by the CUDA-core op agree with the tensor-core rows
    o2 = o.clone()
    o2.view(r, tp, 2 * e)[:, 128:132] = 0
    L.cross_attn_tail(o2, q, kvs[0], kvs[1], kv_slot, tp, 128, 4, lk, heads, 2, False)
    assert_close(from_slab(o2, r, tp, t, 2 * e)[:, :, 128:], got[:, :, 128:], 8e-3, "ops cross_attn_tail")
    hc = 2
    qc = to_slab(rnd(r, 2 * hc * 128, t, seed=77) * (1.4426950408889634 / math.sqrt(128)), tp)
    conds = [rnd(slots * lk, 128, seed=78 + i).to(BF16) for i in range(2)]
    oc = from_slab(L.cross_attn_cond(qc, conds[0], conds[1], kv_slot, tp, t, lk, hc, 2), r, tp, t,
                   2 * hc * 128)
    for s_, c_ in enumerate(conds):
        cf = c_.float().view(slots, lk, 128)[kv_slot.long()][:, None]
        assert_close(oc[:, s_ * hc * 128:(s_ + 1) * hc * 128], attn_ref(qc, cf, cf, s_, hc * 128, 128),
                     1e-2, "ops cross_attn_cond")
    # timestep MLP + FiLM tables (embedding.py:19-43, unet1d_ultimate.py:43-65)
    dim, cols = 256, 384
    tt = torch.tensor([0, 17, 999], dtype=torch.int64, device="cuda")
    w1, b1 = rnd(dim, dim, seed=80) / 16, rnd(dim, seed=81, scale=0.1)
    w2, b2 = rnd(cols, dim, seed=82) / 16, rnd(cols, seed=83, scale=0.1)
    import lm2a_oracle as orc
    emb = orc.sinusoidal_pos_emb(tt.cpu(), dim).double()
    s_ref = F.silu(F.silu(F.linear(emb, w1.cpu().double(), b1.cpu().double())))
    s_got = L.time_mlp(tt, w1, b1)
    assert_close(s_got.cpu().double(), s_ref, 2e-4, "ops time_mlp")
    assert_close(L.film(s_got, w2, b2).cpu().double(),
                 F.linear(s_got.cpu().double(), w2.cpu().double(), b2.cpu().double()), 1e-5, "ops film")
    # Philox normals: same (seed, step) -> same draw; standard normal moments
    seeds = torch.tensor([5, 6], dtype=torch.int64, device="cuda")
    z1, z2 = torch.empty(2, 80, 64, device="cuda"), torch.empty(2, 80, 64, device="cuda")
    L.philox_normal(z1, seeds, 3)
    ops.philox_normal(z2, seeds, 3)
    assert torch.equal(z1, z2) and abs(float(z1.mean())) < 0.05 and abs(float(z1.std()) - 1) < 0.05


def test_upsample2x(ops):
    r, t_in, c = 3, 129, 128
    tp_in, tp_out = 130, 260
    x = rnd(r, c, t_in, seed=40)
    y = torch.full((r * tp_out, c), 5.0, dtype=BF16, device="cuda")
    ops.upsample2x(to_slab(x, tp_in), c, y, c, r, tp_in, t_in, tp_out, c)
    torch.cuda.synchronize()
    ref = F.interpolate(bf(x), scale_factor=2, mode="linear", align_corners=True)
    assert_close(from_slab(y, r, tp_out, 2 * t_in, c), ref, 4e-3, "upsample")
    assert pads_are_zero(y, r, tp_out, 2 * t_in)


def test_ingest(ops):
    b, c, t, tp, ld = 3, 80, 77, 80, 128
    x = rnd(b, c, t, seed=41)
    slab = torch.full((2 * b * tp, ld), 9.0, dtype=BF16, device="cuda")
    arena = torch.full((1000,), 5, dtype=torch.int64, device="cuda")
    ops.ingest_x(x, slab, b, 2, c, t, tp, ld, zero=arena[:998])
    torch.cuda.synchronize()
    ref = to_slab(torch.cat([x, x], 0), tp, ld)
    assert torch.equal(slab, ref)
    assert bool((arena[:998] == 0).all()) and bool((arena[998:] == 5).all())
    seq = rnd(2, 50, 234, seed=42)
    s2 = torch.full((2 * 50, 256), 9.0, dtype=BF16, device="cuda")
    ops.ingest_seq(seq, s2, 2, 50, 234, 50, 256)
    torch.cuda.synchronize()
    ref2 = torch.zeros(2, 50, 256, dtype=BF16, device="cuda")
    ref2[:, :, :234] = seq.to(BF16)
    assert torch.equal(s2.view(2, 50, 256), ref2)


def test_time_mlp_and_film(ops):
    import lm2a_oracle as orc
    rows, dim, cols = 5, 256, 1536
    t = torch.tensor([999, 500, 1, 0, 37], dtype=torch.int64, device="cuda")
    w = rnd(dim, dim, scale=1 / 16, seed=50)
    b = rnd(dim, scale=0.02, seed=51)
    fw = rnd(cols, dim, scale=1 / 16, seed=52)
    fb = rnd(cols, scale=0.02, seed=53)
    s = torch.zeros(rows, dim, device="cuda")
    film = torch.zeros(rows, cols, device="cuda")
    ops.time_mlp(t, w, b, s, rows, dim)
    ops.film(s, fw, fb, film, rows, dim, cols)
    torch.cuda.synchronize()
    emb = orc.sinusoidal_pos_emb(t.cpu(), dim).double()
    temb = F.silu(F.linear(emb, w.cpu().double(), b.cpu().double()))
    ref = F.linear(F.silu(temb), fw.cpu().double(), fb.cpu().double())
    # fp32 sin/cos of arguments up to ~1e3 rad: absolute error ~1e-4 on the embedding
    assert_close(film.cpu().double(), ref, 2e-4, "time mlp + film")


@pytest.mark.parametrize("guided", [True, False])
def test_cfg_posterior_bit_exact(ops, guided):
    """Same operation order as the reference's elementwise sequence -> bit-identical."""
    import lm2a_oracle as orc
    b, c, t, steps = 3, 80, 516, 1000
    x = rnd(b, c, t, seed=60)
    eps = rnd(2 * b if guided else b, c, t, scale=3.0, seed=61)
    noise = rnd(b, c, t, seed=62)
    betas, alphas, abars = orc.diffusion_tables(steps, "cuda")
    sched = torch.stack([1.0 / alphas.sqrt(), betas / (1.0 - abars).sqrt(), betas.sqrt(),
                         torch.zeros_like(betas)], dim=1).contiguous()
    for tt in (999, 1, 0):
        xx = x.clone()
        t_dev = torch.full((2 * b,), tt, dtype=torch.int64, device="cuda")
        ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
        eps_out = torch.zeros(b, c, t, device="cuda")
        ops.cfg_posterior(xx, eps, noise, sched, t_dev, ticket, b, c * t, 2.1, guided, True, eps_out)
        torch.cuda.synchronize()
        e = orc.cfg_eps(eps[:b], eps[b:], 2.1) if guided else eps
        ref = orc.posterior_step(x, e, tt, (betas, alphas, abars), noise)
        assert torch.equal(eps_out, e)
        assert torch.equal(xx, ref), f"t={tt}: max diff {(xx - ref).abs().max()}"
        assert bool((t_dev == tt - 1).all()) and int(ticket) == 0
