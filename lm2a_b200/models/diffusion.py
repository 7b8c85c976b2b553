"""GaussianDiffusion — drop-in for reference models/diffusion.py:5-165 (sampling side).

Same constructor, attributes (model, T, device, dataset_mean, dataset_std, betas, alphas,
alpha_bars) and methods. `p_sample` / `sample` are the reference's plain DDPM
(diffusion.py:61-119); `sample_cfg` is the batched generalisation of the guided loop the
reference only has inline for B = 1 (sample.py:144-210): CFG batch doubling, guidance
blend with both clamps and the posterior update, one CUDA Graph replay per step.

Every tensor op on the path is a C-ABI kernel: the UNet plan plus ONE update kernel per step
(lm2a_cfg_step: CFG blend, clamps, posterior, noise drawn in the kernel from a per-clip
counter-based generator, the next step's input slab). torch is used for memory, the per-clip
seeds (torch.randint on the CPU generator: torch.manual_seed makes a run reproducible) and
stream / graph plumbing. Injected-noise runs (parity tests) take lm2a_cfg_posterior with the
caller's tensors instead.
"""
import os

import torch

from .. import ops


class GaussianDiffusion:
    def __init__(self, model, timesteps=1000, device="cuda", dataset_mean=0.0, dataset_std=1.0):
        self.model = model
        self.T = timesteps
        self.device = device
        self.dataset_mean = dataset_mean
        self.dataset_std = dataset_std
        # linear beta schedule, computed exactly as the reference does (diffusion.py:14-18)
        betas = torch.linspace(1e-4, 0.02, timesteps).to(device)
        self.betas = betas
        self.alphas = 1.0 - betas
        self.alpha_bars = torch.cumprod(self.alphas, dim=0)
        # per-step posterior coefficients with the reference's expressions (diffusion.py:99-102)
        self.sched = torch.stack([1.0 / self.alphas.sqrt(),
                                  self.betas / (1.0 - self.alpha_bars).sqrt(),
                                  self.betas.sqrt(), torch.zeros_like(betas)], dim=1).contiguous()
        self._samplers = {}
        # classifier-free guidance: skip the Q-independent attention branch of the uncond rows
        self.uncond_shortcut = True

    # ---- forward process (training side; trivial torch, kept for API parity) ----------
    def q_sample(self, x0, t, noise=None):
        """diffusion.py:21-37."""
        if noise is None:
            noise = torch.randn_like(x0)
        sqrt_ab = self.alpha_bars[t].sqrt()
        sqrt_mab = (1 - self.alpha_bars[t]).sqrt()
        while sqrt_ab.dim() < x0.dim():
            sqrt_ab = sqrt_ab[..., None]
            sqrt_mab = sqrt_mab[..., None]
        return sqrt_ab * x0 + sqrt_mab * noise

    @torch.no_grad()
    def loss(self, x0, motion_f, text_f):
        """diffusion.py:40-58 evaluated forward-only (the B200 path is inference; training
        and its backward kernels are out of scope — SURVEY.md §2 row 9)."""
        bsz = x0.shape[0]
        t = torch.randint(0, self.T, (bsz,), device=x0.device, dtype=torch.long)
        noise = torch.randn_like(x0)
        x_t = self.q_sample((x0 - self.dataset_mean) / self.dataset_std, t, noise)
        pred = self.model(x_t, t, motion_f, text_f)
        return torch.mean((noise - pred) ** 2)

    # ---- reverse process ---------------------------------------------------------------
    def _t_vector(self, t, bsz, device):
        if not isinstance(t, torch.Tensor):
            return torch.full((bsz,), int(t), device=device, dtype=torch.long)
        if t.dim() == 0:
            return t.view(1).expand(bsz).to(device).contiguous()
        return t.to(device).contiguous()

    @torch.no_grad()
    def p_sample(self, x_t, t, motion_f, text_f):
        """One reverse step x_t -> x_{t-1} (diffusion.py:62-103). Draws randn_like(x_t) for
        every call — also at t == 0, where it is masked — like the reference."""
        t = self._t_vector(t, x_t.size(0), x_t.device)
        eps = self.model(x_t, t, motion_f, text_f)
        noise = torch.randn_like(x_t)
        x_prev = x_t.contiguous().clone()
        ops.cfg_posterior(x_prev, eps, noise, self.sched, t.clone(), None, x_t.size(0),
                          x_t[0].numel(), 1.0, False, False)
        return x_prev

    @torch.no_grad()
    def sample(self, shape, motion_f, text_f):
        """diffusion.py:106-119: x_T ~ N(0, I), then T calls of p_sample."""
        x = torch.randn(shape, device=self.device)
        for t in reversed(range(self.T)):
            t_batch = torch.full((shape[0],), t, device=self.device, dtype=torch.long)
            x = self.p_sample(x, t_batch, motion_f, text_f)
        return x

    def ddim_coefficients(self, t, t_prev, eta=0.0):
        """The five scalars of one DDIM update, evaluated with the reference's own expressions
        (diffusion.py:125-157) in fp32 on `device`: one row of lm2a_cfg_ddim's table
        {sqrt(1-abar_t), sqrt(abar_t), sqrt(abar_prev), sqrt(1-abar_prev-sigma^2), sigma,
        noise gate, 0, 0}."""
        one = torch.tensor(1.0, device=self.alpha_bars.device)
        abar_prev = one if t_prev < 0 else self.alpha_bars[t_prev]
        abar = one if t < 0 else self.alpha_bars[t]
        sigma = eta * torch.sqrt((1 - abar_prev) / (1 - abar) * (1 - abar / abar_prev))
        sigma = torch.nan_to_num(sigma, nan=0.0, posinf=0.0, neginf=0.0)
        zero = torch.zeros_like(one)
        return torch.stack([torch.sqrt(1 - abar), torch.sqrt(abar), torch.sqrt(abar_prev),
                            torch.sqrt(1 - abar_prev - sigma ** 2), sigma,
                            one if t_prev > 0 else zero, zero, zero]).float()

    @torch.no_grad()
    def ddim_sample(self, x, t, t_prev, eps, eta=0.0):
        """One DDIM update (x_t, eps) -> (x_{t_prev}, clamped x0 prediction); same signature
        and arithmetic as reference diffusion.py:124-165, executed by lm2a_cfg_ddim. Draws
        randn_like(x) only when t_prev > 0, like the reference."""
        ops.require_device(x)
        table = self.ddim_coefficients(int(t), int(t_prev), eta).contiguous()
        noise = torch.randn_like(x) if t_prev > 0 else None
        x_prev = x.contiguous().float().clone()
        x0_pred = torch.empty_like(x_prev)
        step = torch.zeros(1, dtype=torch.int32, device=x.device)
        ops.cfg_ddim(x_prev, eps.contiguous().float(), noise, table, None, step, None, None,
                     x.size(0), x[0].numel(), 1.0, False, False, x0_pred)
        return x_prev, x0_pred

    def ddim_timesteps(self, num_steps):
        """Strided sub-sequence of the T training timesteps, descending from T-1 to 0."""
        num_steps = max(1, min(int(num_steps), self.T))
        taus = torch.linspace(self.T - 1, 0, num_steps).round().long().tolist()
        out = []
        for t in taus:  # strictly decreasing (rounding can repeat a value when S ~ T)
            if not out or t < out[-1]:
                out.append(int(t))
        return out

    @torch.no_grad()
    def sample_ddim(self, shape, motion_f, text_f, num_steps=50, eta=0.0, guidance_weight=1.0,
                    x_init=None, noises=None, use_graph=True, report=None):
        """Few-step sampling: the CFG loop of sample.py:144-210 with the posterior
        update replaced by `ddim_sample` over `ddim_timesteps(num_steps)`; step i goes from
        tau_i to tau_{i+1} (t_prev = -1 after the last). Same graph-replayed launch sequence as
        sample_cfg, num_steps model evaluations instead of T."""
        bsz, _, t_len = shape
        s = self.sampler(bsz, t_len, motion_f.shape[1], guidance_weight > 1.0,
                         ddim=(tuple(self.ddim_timesteps(num_steps)), float(eta)))
        return s.run(motion_f, text_f, guidance_weight, x_init, noises, use_graph, report)

    # ---- batched classifier-free-guided sampling (sample.py:131-225, any B) -----------
    def sampler(self, batch, t_len, lk, guided=True, uncond_shortcut=None, ddim=None):
        if uncond_shortcut is None:
            uncond_shortcut = self.uncond_shortcut
        key = (batch, t_len, lk, bool(guided), bool(uncond_shortcut), ddim)
        if key not in self._samplers:
            self._samplers[key] = CfgSampler(self, batch, t_len, lk, guided, uncond_shortcut, ddim)
        return self._samplers[key]

    @torch.no_grad()
    def sample_cfg(self, shape, motion_f, text_f, guidance_weight=1.0, x_init=None, noises=None,
                   use_graph=True, report=None, clip_seeds=None):
        """Returns the final normalised mel x_0 of shape (B, 80, T).

        guidance_weight <= 1 -> plain conditional sampling (sample.py:148-150); otherwise the
        [uncond, cond] doubled batch with the uncond rows attending to zeroed conditions.
        x_init / noises ([steps-1, B, 80, T] or a callable i -> tensor) inject the randomness
        for parity runs. By default x_T and the per-step noise come from a counter-based
        generator keyed per clip (lm2a_cfg_step): `clip_seeds` (B int64 values, e.g. derived from
        the clips' dataset indices) or, if None, seeds drawn with torch.randint - a clip's result
        then depends on its seed only, not on the batch or GPU it was sampled on. (The reference
        never seeds torch, sample.py:136,204: there is no generator stream to reproduce.)"""
        bsz, _, t_len = shape
        s = self.sampler(bsz, t_len, motion_f.shape[1], guidance_weight > 1.0)
        return s.run(motion_f, text_f, guidance_weight, x_init, noises, use_graph, report,
                     clip_seeds)


class CfgSampler:
    """State + CUDA Graph of the per-step launch sequence for a fixed (B, T, Lk)."""

    def __init__(self, diffusion, batch, t_len, lk, guided, uncond_shortcut=True, ddim=None):
        self.d = diffusion
        # ddim = (descending timestep tuple, eta): DDIM updates over that sub-sequence instead
        # of the T-step DDPM posterior
        self.ddim = ddim
        self.batch, self.t_len, self.lk, self.guided = batch, t_len, lk, guided
        eng = diffusion.model.engine()
        self.dev = eng.dev
        rows = 2 * batch if guided else batch
        nslots = batch + 1 if guided else batch
        # uncond rows [0, B) attend to all-zero conditions: their attention blocks reduce to
        # skip(x) + const (exact in real arithmetic; executed FLOPs drop accordingly)
        self.plan = eng.plan(rows, t_len, lk, nslots, 2 if guided else 1, True, uniform_t=True,
                             uncond_rows=batch if (guided and uncond_shortcut) else 0)
        self.noise = torch.zeros(batch, eng.pm.in_dim, t_len, dtype=torch.float32, device=self.dev)
        self.clip_seed = torch.zeros(batch, dtype=torch.int64, device=self.dev)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.gw = 1.0
        self.graph = None          # the captured step in effect
        self._graphs = {}          # (guidance weight, plan launch-list variant) -> graph
        if ddim is not None:
            taus, eta = ddim
            prevs = list(taus[1:]) + [-1]
            self.taus, self.eta = list(taus), eta
            self.table = torch.stack([diffusion.ddim_coefficients(t, tp, eta)
                                      for t, tp in zip(taus, prevs)]).to(self.dev).contiguous()
            self.t_seq = torch.tensor(list(taus) + [0], dtype=torch.int64, device=self.dev)
            self.step_idx = torch.zeros(1, dtype=torch.int32, device=self.dev)
            self.noise_gate = [eta > 0 and tp > 0 for tp in prevs]
        if guided:  # rows [0, B) uncond -> slot 0 (zero conditions); rows [B, 2B) -> 1 + b
            self.kv_slot = torch.cat([torch.zeros(batch, dtype=torch.int32),
                                      torch.arange(1, batch + 1, dtype=torch.int32)]).to(self.dev)
        else:
            self.kv_slot = torch.arange(batch, dtype=torch.int32, device=self.dev)

    def set_conditions(self, motion_f, text_f):
        if self.guided:
            z = torch.zeros_like(motion_f[:1])
            motion_f = torch.cat([z, motion_f], dim=0)
            text_f = torch.cat([z, text_f], dim=0)
        self.plan.set_conditions(motion_f, text_f, self.kv_slot)

    def cond_slabs(self):
        """Where CondProjection.project_raw should write the B clips' projected conditions (the
        zero slot of the uncond rows, if any, sits in front and is left untouched)."""
        return self.plan.cond_slabs(1 if self.guided else 0)

    def _step(self, draw_noise):
        """One step with the noise taken from self.noise (injected draws; draw_noise refills it
        with torch's generator): UNet plan incl. x ingest, then lm2a_cfg_posterior / cfg_ddim."""
        p = self.plan
        if draw_noise:
            self.noise.normal_()
        p.run()
        if self.ddim is None:
            ops.cfg_posterior(p.x_in, p.eps, self.noise, self.d.sched, p.t_in, self.ticket,
                              self.batch, p.x_in[0].numel(), self.gw, self.guided, True)
        else:
            ops.cfg_ddim(p.x_in, p.eps, self.noise if self.eta > 0 else None, self.table,
                         self.t_seq, self.step_idx, p.t_in, self.ticket, self.batch,
                         p.x_in[0].numel(), self.gw, self.guided, True)

    def _fused_step(self):
        """One DDPM step of the default sampling path: UNet launches (input slab and cleared
        statistics already in place), then ONE update kernel that blends, updates, injects the
        per-clip Philox noise, writes the next step's input slab and clears the statistics."""
        p = self.plan
        pm = p.pm
        p.run(skip_ingest=True)
        ops.cfg_step(p.x_in, p.eps, None, self.clip_seed, self.d.sched, p.t_in, self.ticket,
                     self.batch, pm.in_dim, self.t_len, self.gw, self.guided, True,
                     slab=p.x_slab, copies=p.copies, tp=p.geo.Tp[0], ld=pm.in_pad, zero=p.arena)

    def _start_fused(self):
        """Input slab of the first step + cleared statistics (every later step gets both from
        the previous step's update kernel)."""
        p = self.plan
        ops.ingest_x(p.x_in, p.x_slab, self.batch, p.copies, p.pm.in_dim, self.t_len,
                     p.geo.Tp[0], p.pm.in_pad, p.arena)

    def _reset_clock(self):
        """Device-side step state at the start of a trajectory."""
        if self.ddim is None:
            self.plan.t_in.fill_(self.d.T - 1)
        else:
            self.plan.t_in.fill_(self.taus[0])
            self.step_idx.zero_()

    def _ensure_graph(self, fused=None):
        """One captured step per (guidance weight, launch-list variant of the plan: full, or
        constant lyrics stream - chosen per batch when the conditions are set, update kernel).
        fused (default for DDPM): the lm2a_cfg_step path, no library kernel in the graph."""
        p = self.plan
        if fused is None:
            fused = self.ddim is None and not p.fp32
        key = (self.gw, p.const_text, fused)
        self.fused = fused
        if key in self._graphs:
            self.graph = self._graphs[key]
            return
        keep_x, keep_t = p.x_in.clone(), p.t_in.clone()
        # the unfused capture draws from the CUDA generator: put its state back afterwards
        rng_state = torch.cuda.get_rng_state(self.dev)
        if os.environ.get("LM2A_AUTOTUNE", "0") == "1":
            # measured tile shape per GEMM launch instead of the wave model: opt-in, see DESIGN.md
            p.autotune()
        draw = self.ddim is None or self.eta > 0

        def one_step():
            if fused:
                self._fused_step()
            else:
                self._step(draw)

        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self._reset_clock()
            if fused:
                self._start_fused()
            one_step()  # warm-up outside capture: first-launch attribute setup
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        self._reset_clock()
        with torch.cuda.graph(g):
            one_step()
        p.x_in.copy_(keep_x)
        p.t_in.copy_(keep_t)
        self.ticket.zero_()
        torch.cuda.synchronize(self.dev)
        torch.cuda.set_rng_state(rng_state, self.dev)
        self._graphs[key] = self.graph = g

    @torch.no_grad()
    def run(self, motion_f, text_f, guidance_weight=1.0, x_init=None, noises=None,
            use_graph=True, report=None, clip_seeds=None):
        """motion_f / text_f None: the condition slabs were already filled in place
        (cond_slabs + CondProjection.project_raw); only the K/V caches are (re)built.
        noises given -> injected draws through lm2a_cfg_posterior (parity runs, eager launches).
        Otherwise DDPM steps run the fused update kernel with per-clip Philox noise
        (`clip_seeds`, default: torch.randint), x_T from the same generator unless x_init is
        given; DDIM keeps torch's generator for its (eta > 0) draws."""
        with torch.cuda.device(self.dev):
            return self._run(motion_f, text_f, guidance_weight, x_init, noises, use_graph, report,
                             clip_seeds)

    def set_clip_seeds(self, clip_seeds=None):
        if clip_seeds is None:
            clip_seeds = torch.randint(0, 2 ** 62, (self.batch,), dtype=torch.int64)
        seeds = torch.as_tensor(clip_seeds, dtype=torch.int64).reshape(-1)
        if seeds.numel() != self.batch:
            raise RuntimeError(f"clip_seeds must hold {self.batch} values; got {seeds.numel()}")
        self.clip_seed.copy_(seeds.to(self.dev))

    def _run(self, motion_f, text_f, guidance_weight, x_init, noises, use_graph, report,
             clip_seeds):
        d, p = self.d, self.plan
        steps = d.T if self.ddim is None else len(self.taus)
        self.gw = float(guidance_weight)
        if motion_f is None and text_f is None:
            if self.guided:
                p.cond_m[: self.lk].zero_()
                p.cond_t[: self.lk].zero_()
            p.build_kv(self.kv_slot)
        else:
            self.set_conditions(motion_f, text_f)
        injected = noises is not None
        fused = self.ddim is None and not injected and not p.fp32   # (fp32 validation path: unfused)
        if fused or x_init is None:
            self.set_clip_seeds(clip_seeds)
        if x_init is None:
            x_init = torch.empty((self.batch, p.x_in.shape[1], self.t_len), device=self.dev)
            ops.philox_normal(x_init, self.clip_seed, steps)   # counter word one past t = T-1
        if use_graph and not injected:
            self._ensure_graph(fused)
        p.x_in.copy_(x_init)
        self._reset_clock()
        if fused:
            self._start_fused()
        interval = max(1, steps // 10)
        for i in range(steps):
            if self.ddim is None:
                t = steps - 1 - i
                needs_noise = t > 0
            else:
                t = self.taus[i]
                needs_noise = self.noise_gate[i]
            if injected:
                if needs_noise:
                    nz = noises(i) if callable(noises) else noises[i]
                    self.noise.copy_(nz)
                self._step(False)
            elif use_graph:
                self.graph.replay()
            elif fused:
                self._fused_step()
            else:
                self._step(needs_noise)
            if report is not None and (i % interval == 0 or i == steps - 1):
                # the reference's periodic non-finite guard (sample.py:216-223); outside the graph
                if not report(t, p.x_in):
                    break
        return p.x_in.clone()
