"""``KLHR`` -- KL Hit-and-Run with a Gaussian line family, for a batch of chains on B200.

Drop-in for reference ``klhr.py:15-223``: same class name, constructor keywords and
defaults (klhr.py:16-34), same attributes (``D``, ``theta``, ``acceptance_probability``,
``grad_evals``, ``_mean``, ``_cov``, ``_eigvecs``, ``_eigvals``) and methods (``draw``,
``sample``, ``fit``, ``KL``-equivalent fit trace).  Extra keywords: ``chains``, ``dtype``,
``device``, ``process_group`` and the fixed iteration budget ``fit_budget``.

What runs where
  * every draw of every chain: ``klhr_run`` (one launch per stretch of draws between
    adaptation events) -- direction, line fit, proposal, MH, accumulation, all in the kernel;
  * window closures (klhr.py:202-214, at most ~a dozen per run): pooled raw sums are
    all-reduced over the process group and turned into ``_mean``, ``_cov``, eigenpairs on the
    host, identically on every rank.

Documented deviations from the reference (DESIGN.md has the full list)
  * ``theta=`` is honoured (the reference's ``_initialize`` overwrites it, klhr.py:94);
  * adaptation pools over chains instead of over one chain's history; the PCA is the
    eigen-decomposition of the pooled second-moment matrix, sampled every ``pca_stride`` draws;
  * over-relaxed proposals (klhr.py:160-173) draw their binomial / beta variates from the chain's
    Philox stream instead of SciPy's global RNG; the Smoother signal that adapts K is pooled.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from .adaptation import OnlineMoments, OnlinePCA, Smoother, WindowedAdaptation, allreduce_adaptation
from .mcmc import MCMCBase


class KLHR(MCMCBase):
    _family = "gauss"
    _eigen_weights_normalised = False          # klhr.py:150 sums evals * eigvecs (un-normalised)
    _adapt_K = True                            # klhr.py:212-214 adapts K at window closures

    def __init__(self, bsmodel, theta=None, seed=None, N=8, K=10, J=2, l=4, initscale=0.1, warmup=1_000,
                 windowsize=50, windowscale=2, tol=1e-12, grad_clip=1e15, scale_clip=600,
                 scale_dir_cov=False, overrelaxed=False, eigen_method_one=True, max_init_tries=100, *,
                 chains=1, dtype=torch.float64, device=None, process_group=None, chain_offset=None,
                 pca_stride=None, moments_every_draw=False, fit_budget=None):
        super().__init__(bsmodel, -1, theta=theta, seed=seed, chains=chains, dtype=dtype, device=device)
        if not 1 <= int(K) <= 50:
            raise ValueError("K must be in 1..50 (the reference clips it there, klhr.py:213)")
        self.N, self.K, self.l = N, int(K), l
        self.J = self._clip_J(J)
        self._tol, self._grad_clip, self._scale_clip = tol, grad_clip, scale_clip
        self._max_init_tries = max_init_tries
        self._initscale = initscale
        self.x, self.w = engine.gauss_hermite(N)
        budget = dict(fit_budget or {})
        self._fit = engine.FitConfig(family=self._family, N=N, initscale=initscale, tol=tol,
                                     scale_clip=float(scale_clip), x=self.x, w=self.w,
                                     **{"n2": 24 if self._family == "gauss" else 48,
                                        "kmax": 0 if self._family == "gauss" else 32, **budget}).for_dtype(dtype)
        self._fit.overrelax_K = self.K if overrelaxed else 0       # klhr.py:160-173 / klhr_sinh.py:215-228
        self._fit.grad_clip = self._kl_grad_clip()
        self._smoothK = Smoother(self.K)
        self._windowedadaptation = WindowedAdaptation(warmup, windowsize=windowsize, windowscale=windowscale)
        self._scale_dir_cov = scale_dir_cov
        self._overrelaxed = overrelaxed
        self._eigen_method_one = eigen_method_one
        # snapshots of the ensemble that feed the window moments / PCA sums: every `pca_stride` draws, or (None) as
        # sparse as keeps >= ~2048 pooled samples and >= 4 snapshots per window -- one chain: every draw, like the
        # reference's per-draw update (klhr.py:216-218); 65 536 chains: 4 per window
        self._pca_stride = None if pca_stride is None else max(1, int(pca_stride))
        # False (default): the window moments come from the same ensemble snapshots as the PCA sums
        # (every pca_stride draws, all chains) and the warm-up runs on the fast kernels; True: every
        # non-closure draw of every chain is accumulated in-kernel like klhr.py:216-218 does per draw
        self._moments_every_draw = bool(moments_every_draw)
        self._group = process_group
        dev = self.device
        self._onlinemoments = OnlineMoments(self.D, device=dev)
        self._onlinemoments_density = OnlineMoments(self.D, device=dev)
        self._onlinepca = OnlinePCA(self.D, K=self.J, l=l, device=dev)
        self._mean = np.zeros(self.D)
        self._cov = np.ones(self.D)
        ncol = self.J + 1 if eigen_method_one else self.J
        self._eigvecs = np.zeros((self.D, ncol))
        self._eigvals = np.ones(ncol)
        self._draw = 0
        self._outer_scratch = None
        self._acc_seen = 0.0
        self._smooth_log = []                      # (accepted draws of this rank, steps, smoother updates) per launch
        self._accept_count = torch.zeros(self.chains, dtype=torch.int64, device=dev)
        self._evals_total = torch.zeros(1, dtype=torch.int64, device=dev)
        if chain_offset is None:
            chain_offset = 0
            if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
                chain_offset = torch.distributed.get_rank(process_group) * self.chains
        self._chain_offset = int(chain_offset)
        self._shift_dev = torch.zeros(self.D, dtype=dtype, device=dev)
        self._direction = None
        self._refresh_direction()
        if theta is None:
            self._initialize()

    # ------------------------------------------------------------------ reference quirks kept per class
    def _kl_grad_clip(self):
        """Elementwise clip of the model gradient inside ``KL``: none in KLHR (klhr.py:106-120 calls the model
        directly; ``_logp_grad`` :101-104 is unused)."""
        return 0.0

    def _clip_J(self, J):
        return J if J < self.D else self.D - 1                 # klhr.py:39

    # ------------------------------------------------------------------ start points (klhr.py:87-99)
    def _initialize(self):
        """Random finite start per chain (klhr.py:87-99: N(0, initscale^2), retried until lp and the gradient are
        finite).  Chain c's start is a function of (seed, chain_offset + c) only -- blocks of 4096 global chain ids
        own one generator each -- so a sharded run starts exactly where the unsharded one does."""
        blk = 4096
        lo, hi = self._chain_offset, self._chain_offset + self.chains
        for b in range(lo // blk, (hi - 1) // blk + 1):
            g = torch.Generator(device="cpu").manual_seed((self.seed * 0x9E3779B97F4A7C15 + b * 0xD1B54A32D192ED03 + 1) % (1 << 63))
            a, z = max(lo, b * blk), min(hi, (b + 1) * blk)
            sl = slice(a - lo, z - lo)
            todo = torch.ones(z - a, dtype=torch.bool, device=self.device)
            for _ in range(self._max_init_tries):
                if not bool(todo.any()):
                    break
                cand = (torch.randn(blk, self.D, generator=g, dtype=torch.float64) * self._initscale)[a - b * blk:z - b * blk]
                cand = cand.to(self.device, self.dtype)
                lp, grad = self.model.log_density_gradient(cand)
                ok = todo & torch.isfinite(lp) & torch.isfinite(grad.norm(dim=1))
                self._theta[sl][ok] = cand[ok]
                todo &= ~ok
            if bool(todo.any()):
                raise RuntimeError("failed to initialize")

    # ------------------------------------------------------------------ direction law (klhr.py:143-153)
    def _refresh_direction(self):
        dev, dt = self.device, self.dtype
        lam = np.asarray(self._eigvals, dtype=np.float64)
        p = lam / np.sum(lam)
        sd = torch.as_tensor(np.sqrt(np.maximum(self._cov, 0.0)), dtype=dt, device=dev).contiguous()
        if self._eigen_method_one:
            # the extra column J is identically zero (klhr.py:64-66): it is not shipped to the device
            cols = torch.as_tensor(np.ascontiguousarray(self._eigvecs[:, :self.J].T), dtype=dt, device=dev)
            cols = cols.reshape(self.J, self.D).contiguous()
            cdf = np.cumsum(p)
            cdf /= cdf[-1]
            cdf_t = torch.as_tensor(cdf, dtype=dt, device=dev).contiguous()
            self._direction = engine.Direction(mean_cols=cols if self.J > 0 else None, sd=sd, cdf=cdf_t,
                                               n_zero_cols=1 if self.J > 0 else 0)
        else:
            wgt = p if self._eigen_weights_normalised else lam
            m = np.sum(wgt * self._eigvecs, axis=1) if self._eigvecs.shape[1] else np.zeros(self.D)
            cols = torch.as_tensor(m[None, :], dtype=dt, device=dev).contiguous()
            self._direction = engine.Direction(mean_cols=cols, sd=sd, cdf=None)

    # ------------------------------------------------------------------ the loop (klhr.py:196-223)
    def _advance(self, n, draws=None, thin=1, chain_s1=None, chain_s2=None, stat_shift=None):
        """Advance every chain by ``n`` draws.  ``draws`` (n // thin, B, D) receives thinned
        states; chain_s1/chain_s2 accumulate per-chain sums (post-warm-up diagnostics)."""
        wa = self._windowedadaptation
        done = 0
        while done < n:
            nxt = wa.next_closure(self._draw)
            adapting = nxt is not None
            steps = n - done
            closes = False
            snap = False
            if adapting:
                # ensemble snapshots sit at ABSOLUTE draw indices (window start + k * stride), so that a run split
                # by sample() calls or a checkpoint takes them where the uninterrupted run does
                start, stride = self._window_start(nxt), self._stride_now(nxt)
                nxt_snap = start + ((self._draw - start) // stride + 1) * stride
                steps = min(steps, nxt - self._draw, nxt_snap - self._draw)
                closes = self._draw + steps == nxt
                snap = self._draw + steps == nxt_snap
            kw = {}
            if adapting and self._moments_every_draw:
                mom = self._onlinemoments
                kw.update(shift=self._shift_dev, pooled_s1=mom.s1, pooled_s2=mom.s2, skip_accum_last=closes)
            elif chain_s1 is not None:
                kw.update(shift=stat_shift, chain_s1=chain_s1, chain_s2=chain_s2)
            self._launch(steps, draws=draws, thin=thin, thin_offset=done, **kw)
            if adapting and self._overrelaxed and self._adapt_K:
                # Smoother signal (klhr.py:219-221: +1 for a chain that moved, -1 otherwise), pooled over ALL ranks:
                # the accept counts stay on the device until the closure, where they are summed with the moments
                self._smooth_log.append((self._accept_count.sum(dtype=torch.float64), steps, steps - (1 if closes else 0)))
            self._draw += steps
            done += steps
            if adapting:
                if self._moments_every_draw:
                    self._onlinemoments.add_sums(self.chains * (steps - (1 if closes else 0)))
                if closes:
                    self._close_window()
                elif snap:
                    self._snapshot_update()

    def _window_start(self, closure):
        prev = 0
        for c in self._windowedadaptation.closures:
            if c >= closure:
                break
            prev = c
        return prev

    def _stride_now(self, closure):
        if self._pca_stride is not None:
            return self._pca_stride
        length = max(1, closure - self._window_start(closure))
        return max(1, min(length // 4, (self.chains * length) // 2048))

    def _launch(self, steps, **kw):
        """``steps`` draws for every chain in one launch (subclasses swap the transition kernel)."""
        engine.run(self.model, self._fit, self._theta, steps, self.seed, self._direction,
                   chain_offset=self._chain_offset, draw_offset=self._draw,
                   accept_count=self._accept_count, evals_total=self._evals_total, **kw)

    def _snapshot_update(self):
        """Pooled analogue of klhr.py:216-219 on the current ensemble: PCA second moments of
        (theta - _mean) and, for ``scale_dir_cov``, gradient moments.  The sums of every slice of 1024 chains go to
        its own scratch plane; the planes are folded at the closure (bit-reproducible, sharding-invariant)."""
        pca, mom = self._onlinepca, self._onlinemoments
        if self._outer_scratch is None:
            self._outer_scratch = engine.outer_scratch(self._theta)
        engine.outer_accumulate(self._theta, self._shift_dev, pca.outer, mom.s1, scratch=self._outer_scratch)
        if not self._moments_every_draw:        # first moments ride along; second moments are the diagonal of the outer-product sum
            mom.N += self.chains
        pca.add_sums(self.chains)
        if self._scale_dir_cov:
            _, g = self.model.log_density_gradient(self._theta)
            self._onlinemoments_density.update(torch.clamp(g, -self._grad_clip, self._grad_clip))

    def _close_window(self):
        """klhr.py:202-214 with pooled sums; identical result on every rank."""
        mom, gmom, pca = self._onlinemoments, self._onlinemoments_density, self._onlinepca
        if self._outer_scratch is not None:
            s1_planes = torch.zeros_like(mom.s1) if self._moments_every_draw else mom.s1
            engine.outer_reduce(self._outer_scratch, pca.outer, s1_planes, self.chains, self.D)
        if not self._moments_every_draw:
            mom.s2.copy_(torch.diagonal(pca.outer))
        extra = []
        if self._smooth_log:
            extra = [torch.stack([t for t, _, _ in self._smooth_log])]
        allreduce_adaptation([mom, gmom], pca, group=self._group, extra=extra)
        self._mean = mom.mean().cpu().numpy()
        self._cov = mom.var().cpu().numpy()
        if self._scale_dir_cov:
            self._cov = self._cov / (self._tol + gmom.var().cpu().numpy())
        self._eigvecs[:, :self.J] = pca.vectors()
        self._eigvals[:self.J] = pca.values()
        self._shift_dev = torch.as_tensor(self._mean, dtype=self.dtype, device=self.device).contiguous()
        mean64 = torch.as_tensor(self._mean, dtype=torch.float64, device=self.device)
        mom.reset(shift=mean64)
        gmom.reset()
        pca.reset()
        if self._overrelaxed and self._adapt_K:                    # klhr.py:212-214
            world = 1
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                world = torch.distributed.get_world_size(self._group)
            if self._smooth_log:
                tot = extra[0].cpu().numpy() / (self.chains * world)      # mean accepted draws per chain, all ranks
                for (_, steps, n_upd), acc_now in zip(self._smooth_log, tot):
                    frac = (acc_now - self._acc_seen) / steps
                    self._acc_seen = acc_now
                    for _ in range(n_upd):
                        self._smoothK.update(2.0 * frac - 1.0)
            self.K = int(np.clip(self._smoothK.optimum(), 1, 50))
            self._fit.overrelax_K = self.K
        self._smooth_log = []
        self._smoothK.reset()
        self._refresh_direction()

    # ------------------------------------------------------------------ public API
    def draw(self):
        self._advance(1)
        return self.theta

    def sample(self, M, thin=1, out=None, chunk_rows=None):
        """``M`` rows, row 0 the current state (mcmc.py:31-37); ``thin`` keeps every thin-th draw.

        ``out``: a pinned host tensor (M, chains, D) of the sampler's dtype.  The rows are then streamed to it while
        the chains keep running: the kernel writes chunk k of ``chunk_rows`` rows into one of two device buffers
        while chunk k-1 travels device -> host on a copy stream, so the device never holds more than two chunks
        (at 65 536 chains x D = 100 x fp64 one row is 52 MB and PCIe, not the kernel, sets the pace)."""
        if out is None:
            dev_out = torch.empty(M, self.chains, self.D, dtype=self.dtype, device=self.device)
            dev_out[0] = self._theta
            if M > 1:
                self._advance((M - 1) * thin, draws=dev_out[1:], thin=thin)
            if self.chains == 1:
                return dev_out[:, 0].double().cpu().numpy()
            return dev_out
        if (tuple(out.shape) != (M, self.chains, self.D) or out.dtype != self.dtype or out.is_cuda
                or not out.is_pinned() or not out.is_contiguous()):
            raise ValueError("out must be a contiguous pinned host tensor of shape (M, chains, D) and the sampler's dtype")
        row_bytes = self.chains * self.D * out.element_size()
        chunk = int(chunk_rows) if chunk_rows else max(1, min(M, (128 << 20) // row_bytes))
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        cs = self._copy_stream
        bufs = [torch.empty(chunk, self.chains, self.D, dtype=self.dtype, device=self.device) for _ in range(2)]
        freed = [None, None]
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(cs):
            cs.wait_event(ready)
            out[0].copy_(self._theta, non_blocking=True)
            first = torch.cuda.Event()
            first.record(cs)
        cur.wait_event(first)                                        # row 0 is read before the state moves on
        for k, r0 in enumerate(range(1, M, chunk)):
            n = min(chunk, M - r0)
            buf = bufs[k % 2]
            if freed[k % 2] is not None:
                cur.wait_event(freed[k % 2])                         # its previous content has reached the host
            self._advance(n * thin, draws=buf[:n], thin=thin)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(cs):
                cs.wait_event(done)
                out[r0:r0 + n].copy_(buf[:n], non_blocking=True)
                freed[k % 2] = torch.cuda.Event()
                freed[k % 2].record(cs)
        cs.synchronize()
        return out

    def run(self, n, chain_stats=False):
        """Advance ``n`` draws without storing them.  With ``chain_stats`` returns per-chain
        sums ``(s1, s2)`` of shape (B, D) over the draws (for ESS / MCSE, diagnostics.py)."""
        if not chain_stats:
            self._advance(n)
            return None
        if self._moments_every_draw and self._windowedadaptation.next_closure(self._draw) is not None:
            # the kernel has ONE accumulator slot per chain: during warm-up it belongs to the per-draw adaptation moments
            raise RuntimeError("chain_stats during warm-up needs moments_every_draw=False (the per-draw warm-up "
                               "accumulation uses the kernel's accumulator slot); finish the warm-up first")
        s1 = torch.zeros(self.chains, self.D, dtype=torch.float64, device=self.device)
        s2 = torch.zeros_like(s1)
        self._advance(n, chain_s1=s1, chain_s2=s2, stat_shift=None)
        return s1, s2

    def fit(self, rho, z_init=None):
        """Line fit through the current state along ``rho`` ((D,) or (B, D)); returns eta
        ((2,) / (B, 2); (4,) for the sinh family) like reference ``fit`` (klhr.py:126-141).
        Does not advance the chains."""
        rho_t = torch.as_tensor(np.asarray(rho.detach().cpu() if torch.is_tensor(rho) else rho,
                                           dtype=np.float64)).reshape(-1, self.D)
        rho_t = rho_t.expand(self.chains, self.D).to(self.device, self.dtype).contiguous()
        B = self.chains
        z = (torch.as_tensor(self.rng.normal(size=B)) if z_init is None
             else torch.as_tensor(np.broadcast_to(np.asarray(z_init, dtype=np.float64), (B,)).copy()))
        z = z.to(self.device, self.dtype).contiguous()
        zeros = torch.zeros(B, dtype=self.dtype, device=self.device)
        half = torch.full((B,), 0.5, dtype=self.dtype, device=self.device)
        init4 = None
        if self._family == "sinh":
            init4 = torch.as_tensor(self.rng.normal(size=(B, 4))).to(self.device, self.dtype).contiguous()
        scratch = self._theta.clone()
        import dataclasses
        fit_only = dataclasses.replace(self._fit, overrelax_K=0)      # the proposal is irrelevant to the fit
        tr = engine.step_replay(self.model, fit_only, scratch, rho_t, z, zeros, half, init4=init4)
        eta = tr.eta[0]
        return eta[0].double().cpu().numpy() if self.chains == 1 else eta

    def KL(self, eta, rho):
        """The KL objective and its gradient at ``eta`` along ``rho`` through the current state(s), like
        reference ``KL`` (klhr.py:106-120 / klhr_sinh.py:163-176): ``(float, ndarray)`` for one chain,
        ``(B,)`` / ``(B, n)`` tensors for a batch.  ``eta``: (n,) or (B, n); ``rho``: (D,) or (B, D)."""
        f, g = engine.kl_eval(self.model, self._fit, self._theta, self._bcast(rho, self.D), self._bcast(eta, self._fit.n_eta))
        if self.chains == 1:
            return float(f[0]), g[0].double().cpu().numpy()
        return f, g

    def _bcast(self, v, n):
        t = torch.as_tensor(np.asarray(v.detach().cpu() if torch.is_tensor(v) else v, dtype=np.float64)).reshape(-1, n)
        return t.expand(self.chains, n).to(self.device, self.dtype).contiguous()

    def swap_state(self, theta):
        """Use ``theta`` (B, D) -- a contiguous device tensor of the sampler's dtype -- as the live chain
        state WITHOUT copying, e.g. to double-buffer host transfers against ``run``."""
        if tuple(theta.shape) != (self.chains, self.D) or theta.dtype != self.dtype or not theta.is_cuda \
                or not theta.is_contiguous():
            raise ValueError("swap_state needs a contiguous (chains, D) CUDA tensor of the sampler's dtype")
        self._theta = theta

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self):
        """Everything needed to continue the run bit for bit: the RNG is counter based (a function of
        seed, chain id and draw index), so no generator state has to be saved.  (The reference has no
        checkpointing; its parquet dumps are commented out.)"""
        mom, gmom, pca = self._onlinemoments, self._onlinemoments_density, self._onlinepca
        return {
            "theta": self._theta.detach().cpu().clone(), "seed": self.seed, "draw": self._draw,
            "chain_offset": self._chain_offset, "K": self.K, "acc_seen": self._acc_seen,
            "accept_count": self._accept_count.cpu().clone(), "evals_total": self._evals_total.cpu().clone(),
            "mean": np.array(self._mean), "cov": np.array(self._cov), "eigvecs": np.array(self._eigvecs),
            "eigvals": np.array(self._eigvals),
            "moments": [(m.N, m.s1.cpu().clone(), m.s2.cpu().clone(), m.shift.cpu().clone()) for m in (mom, gmom)],
            "pca": (pca.n, pca.outer.cpu().clone()),
            "planes": None if self._outer_scratch is None else self._outer_scratch.cpu().clone(),
            "smoothK": (self._smoothK._x, self._smoothK._count),
            "smooth_log": [(float(t), a, b) for t, a, b in self._smooth_log],
        }

    def load_state_dict(self, sd):
        if tuple(sd["theta"].shape) != (self.chains, self.D):
            raise ValueError("checkpoint shape does not match this sampler")
        dev = self.device
        self._theta.copy_(sd["theta"].to(dev, self.dtype))
        self.seed, self._draw, self._chain_offset = int(sd["seed"]), int(sd["draw"]), int(sd["chain_offset"])
        self.K, self._acc_seen = int(sd["K"]), float(sd["acc_seen"])
        if self._fit.overrelax_K:
            self._fit.overrelax_K = self.K
        self._accept_count.copy_(sd["accept_count"].to(dev))
        self._evals_total.copy_(sd["evals_total"].to(dev))
        self._mean, self._cov = np.array(sd["mean"]), np.array(sd["cov"])
        self._eigvecs, self._eigvals = np.array(sd["eigvecs"]), np.array(sd["eigvals"])
        for m, (n, s1, s2, shift) in zip((self._onlinemoments, self._onlinemoments_density), sd["moments"]):
            m.N = int(n)
            m.s1.copy_(s1.to(dev))
            m.s2.copy_(s2.to(dev))
            m.shift = shift.to(dev, torch.float64).clone()
        self._onlinepca.n = int(sd["pca"][0])
        self._onlinepca.outer.copy_(sd["pca"][1].to(dev))
        self._onlinepca._eig = None
        self._smoothK._x, self._smoothK._count = sd["smoothK"]
        self._smooth_log = [(torch.tensor(t, dtype=torch.float64, device=dev), a, b) for t, a, b in sd.get("smooth_log", [])]
        self._outer_scratch = None if sd.get("planes") is None else sd["planes"].to(dev)
        self._shift_dev = torch.as_tensor(self._mean, dtype=self.dtype, device=dev).contiguous()
        self._refresh_direction()

    @property
    def acceptance_probability(self):
        """Mean accept rate over all chains and draws (running mean of klhr.py:192-193)."""
        if self._draw == 0:
            return 0.0
        return float(self._accept_count.double().mean()) / self._draw

    @property
    def grad_evals(self):
        """Line evaluations executed, summed over chains (klhr.py:132,140 counts model calls)."""
        return int(self._evals_total.item())
