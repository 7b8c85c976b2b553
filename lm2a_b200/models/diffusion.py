"""GaussianDiffusion — drop-in for reference models/diffusion.py:5-165 (sampling side).

Same constructor, attributes (model, T, device, dataset_mean, dataset_std, betas, alphas,
alpha_bars) and methods. `p_sample` / `sample` are the reference's plain DDPM
(diffusion.py:61-119); `sample_cfg` is the batched generalisation of the guided loop the
reference only has inline for B = 1 (sample.py:144-210): CFG batch doubling, guidance
blend with both clamps and the posterior update, one CUDA Graph replay per step.

Every tensor op on the path is a C-ABI kernel (lm2a_cfg_posterior + the UNet plan);
torch is used for memory, RNG draws (torch.randn — the same generator stream the
reference consumes) and stream / graph plumbing.
"""
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

    def ddim_sample(self, x, t, t_prev, eps, eta=0.0):
        """diffusion.py:124-165 (the reference marks it undebugged and never calls it; kept
        for API parity, elementwise torch)."""
        one = torch.tensor(1.0, device=x.device)
        abar_prev = one if t_prev < 0 else self.alpha_bars[t_prev]
        abar = one if t < 0 else self.alpha_bars[t]
        abar, abar_prev = abar[..., None, None], abar_prev[..., None, None]
        x0_pred = torch.clamp((x - eps * torch.sqrt(1 - abar)) / torch.sqrt(abar), -2.0, 2.0)
        sigma = eta * torch.sqrt((1 - abar_prev) / (1 - abar) * (1 - abar / abar_prev))
        sigma = torch.nan_to_num(sigma, nan=0.0, posinf=0.0, neginf=0.0)
        noise = torch.randn_like(x) if t_prev > 0 else torch.zeros_like(x)
        x_prev = (torch.sqrt(abar_prev) * x0_pred
                  + torch.sqrt(1 - abar_prev - sigma ** 2) * eps + sigma * noise)
        return x_prev, x0_pred

    # ---- batched classifier-free-guided sampling (sample.py:131-225, any B) -----------
    def sampler(self, batch, t_len, lk, guided=True, uncond_shortcut=None):
        if uncond_shortcut is None:
            uncond_shortcut = self.uncond_shortcut
        key = (batch, t_len, lk, bool(guided), bool(uncond_shortcut))
        if key not in self._samplers:
            self._samplers[key] = CfgSampler(self, batch, t_len, lk, guided, uncond_shortcut)
        return self._samplers[key]

    @torch.no_grad()
    def sample_cfg(self, shape, motion_f, text_f, guidance_weight=1.0, x_init=None, noises=None,
                   use_graph=True, report=None):
        """Returns the final normalised mel x_0 of shape (B, 80, T).

        guidance_weight <= 1 -> plain conditional sampling (sample.py:148-150); otherwise the
        [uncond, cond] doubled batch with the uncond rows attending to zeroed conditions.
        x_init / noises ([steps-1, B, 80, T] or a callable i -> tensor) inject the randomness
        for parity runs; by default x_T and the per-step noise come from torch.randn on
        `device` in the reference's order (one draw per step; none consumed at t = 0)."""
        bsz, _, t_len = shape
        s = self.sampler(bsz, t_len, motion_f.shape[1], guidance_weight > 1.0)
        return s.run(motion_f, text_f, guidance_weight, x_init, noises, use_graph, report)


class CfgSampler:
    """State + CUDA Graph of the per-step launch sequence for a fixed (B, T, Lk)."""

    def __init__(self, diffusion, batch, t_len, lk, guided, uncond_shortcut=True):
        self.d = diffusion
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
        self.ticket = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.gw = 1.0
        self.graph = None
        self.graph_gw = None
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

    def _step(self, draw_noise):
        p = self.plan
        if draw_noise:
            self.noise.normal_()
        p.run()
        ops.cfg_posterior(p.x_in, p.eps, self.noise, self.d.sched, p.t_in, self.ticket, self.batch,
                          p.x_in[0].numel(), self.gw, self.guided, True)

    def _ensure_graph(self):
        if self.graph is not None and self.graph_gw == self.gw:
            return
        p = self.plan
        keep_x, keep_t = p.x_in.clone(), p.t_in.clone()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            p.t_in.fill_(self.d.T - 1)
            self._step(True)  # warm-up outside capture: first-launch attribute setup
        torch.cuda.current_stream(self.dev).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        p.t_in.fill_(self.d.T - 1)
        with torch.cuda.graph(g):
            self._step(True)
        p.x_in.copy_(keep_x)
        p.t_in.copy_(keep_t)
        self.ticket.zero_()
        self.graph, self.graph_gw = g, self.gw

    @torch.no_grad()
    def run(self, motion_f, text_f, guidance_weight=1.0, x_init=None, noises=None,
            use_graph=True, report=None):
        d, p = self.d, self.plan
        steps = d.T
        self.gw = float(guidance_weight)
        self.set_conditions(motion_f, text_f)
        if x_init is None:
            x_init = torch.randn((self.batch, p.x_in.shape[1], self.t_len), device=self.dev)
        injected = noises is not None
        if use_graph and not injected:
            self._ensure_graph()
        p.x_in.copy_(x_init)
        p.t_in.fill_(steps - 1)
        interval = max(1, steps // 10)
        for i in range(steps):
            t = steps - 1 - i
            if injected:
                if t > 0:
                    nz = noises(i) if callable(noises) else noises[i]
                    self.noise.copy_(nz)
                self._step(False)
            elif use_graph:
                self.graph.replay()
            else:
                self._step(t > 0)
            if report is not None and (i % interval == 0 or t == 0):
                # the reference's periodic non-finite guard (sample.py:216-223); outside the graph
                if not report(t, p.x_in):
                    break
        return p.x_in.clone()
