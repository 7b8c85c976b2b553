"""Fused Adan step (lm2a_adan_step, SURVEY §8 f4 first piece) on B200 against the reference
optimizer's own outputs (tests/golden/adan.npz <- oracle/make_golden_adan.py) and against the
oracle restatement at larger tensors. Tolerance: the moments (m, v, n, prev_grad) are compared
bit for bit; the parameters within 4 fp32 ulp of the CPU golden (torch's CPU sqrt is not
correctly rounded — it differs from IEEE sqrt in ~0.5 % of elements — and torch's CUDA div_ by a
scalar multiplies by the reciprocal; torch CPU and torch CUDA do not agree bit for bit on this
update either), also against the same torch op sequence executed on the GPU."""
import os

import numpy as np
import pytest
import torch

import lm2a_oracle as orc
from make_golden_adan import EMA_DECAY, HP, SHAPES, STEPS

pytestmark = pytest.mark.gpu


def test_adan_matches_reference_golden(golden_dir):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lm2a_b200.models.adan import Adan
    d = np.load(os.path.join(golden_dir, "adan.npz"))
    params = [torch.nn.Parameter(torch.from_numpy(d[f"p0_{i}"].copy()).cuda())
              for i in range(len(SHAPES))]
    shadow = [p.detach().clone() for p in params]
    opt = Adan(params, **HP)
    for step in range(STEPS):
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(d[f"g{step}_{i}"].copy()).cuda()
        opt.step(ema=(shadow, EMA_DECAY))
        torch.cuda.synchronize()
        for i, p in enumerate(params):
            np.testing.assert_allclose(p.detach().cpu().numpy(), d[f"p{step + 1}_{i}"],
                                       rtol=5e-7, atol=1e-9)
            np.testing.assert_allclose(shadow[i].cpu().numpy(), d[f"ema{step + 1}_{i}"],
                                       rtol=5e-7, atol=1e-9)
    for i, p in enumerate(params):
        st = opt.state[p]
        assert st["step"] == STEPS
        for k in ("m", "v", "n", "prev_grad"):
            np.testing.assert_array_equal(st[k].cpu().numpy(), d[f"{k}_{i}"])
    # same state layout as the reference optimizer (state_dict round trip)
    sd = opt.state_dict()
    assert set(sd["state"][0]) == {"step", "prev_grad", "m", "v", "n"}
    with pytest.raises(RuntimeError, match="restart_cond"):
        Adan(params, restart_cond=lambda s: False)


def test_adan_large_ragged_tensors_vs_oracle():
    """Chunking: a tensor larger than one 64 K chunk with a ragged tail, an unaligned view, a
    parameter without gradient (skipped) — against the oracle's torch op sequence executed on
    the same GPU (what the reference optimizer does in training): moments bit for bit,
    parameters within 4 ulp, no EMA."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from lm2a_b200.models.adan import Adan
    g = torch.Generator().manual_seed(3)
    shapes = [(200003,), (3, 65537), (7,)]
    cpu = [torch.randn(s, generator=g).cuda() for s in shapes]
    params = [torch.nn.Parameter(c.clone()) for c in cpu] + [
        torch.nn.Parameter(torch.zeros(5).cuda())]
    opt = Adan(params, lr=1e-3, weight_decay=0.0)
    states = [{"step": 0, "prev_grad": torch.zeros_like(c), "m": torch.zeros_like(c),
               "v": torch.zeros_like(c), "n": torch.zeros_like(c)} for c in cpu]
    for step in range(3):
        grads = [torch.randn(s, generator=g) for s in shapes]
        for p, gr in zip(params, grads):
            p.grad = gr.cuda()
        opt.step()
        for c, gr, st in zip(cpu, grads, states):
            orc.adan_step(c, gr.cuda(), st, 1e-3, (0.02, 0.08, 0.01), 1e-8, 0.0)
        for p, c, st in zip(params, cpu, states):
            assert torch.equal(opt.state[p]["m"], st["m"]) and torch.equal(opt.state[p]["n"], st["n"])
            assert torch.equal(opt.state[p]["v"], st["v"])
            np.testing.assert_allclose(p.detach().cpu().numpy(), c.cpu().numpy(), rtol=5e-7,
                                       atol=1e-9, err_msg=f"step {step}")
    assert float(params[3].detach().abs().sum()) == 0.0 and len(opt.state[params[3]]) == 0
