"""Thin host layer over the C ABI: descriptors, launches, trace buffers.

Nothing here computes on the CPU: tensors are allocated by PyTorch on the CUDA device and
handed to ``libklhr_sm100.so`` by pointer.  The samplers (``klhr.py``, ``klhr_sinh.py``)
are built on these three calls:

    step_replay(...)   one draw for every chain with injected rho / variates
    run(...)           n draws for every chain with in-kernel Philox streams
    outer_accumulate() pooled second moments for the adaptation PCA
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch
from numpy.polynomial.hermite import hermgauss

from . import _lib
from .bsmodel import BSModel, _dtype_code


def gauss_hermite(N: int):
    """Probabilists' Gauss-Hermite table with sum(w) = 1 (reference ``klhr.py:46-49``)."""
    x, w = hermgauss(N)
    return x * np.sqrt(2), w / np.sqrt(np.pi)


@dataclass
class FitConfig:
    """Reference constructor arguments that reach the fit plus the fixed iteration budget
    replacing ``scipy.optimize.minimize`` (DESIGN.md "Optimiser")."""
    family: str = "gauss"
    N: int = 8
    initscale: float = 0.1
    tol: float = 1e-12
    scale_clip: float = 600.0
    n1: int = 12
    n2: int = 24
    nb: int = 8
    kmax: int = 0                      # cap on stage-2 KL evaluations (0: 1 + n2 * nb)
    overrelax_K: int = 0               # K > 0: over-relaxed proposals with K trials (klhr.py:160-173)
    fix_d: bool = False                # sinh family with d = 1 frozen (reference sub_klhr_sinh.py)
    gtol1: float = 1e-4                # stage 1 only supplies the start of stage 2 (1e-8: up to +36 % evaluations on arK)
    gtol2: float = 1e-10
    step_cap: float = 2.0
    c1: float = 1e-4
    basin: float = 1e-3
    grad_clip: float = 0.0             # sinh family: elementwise clip of the model gradient in KL (klhr_sinh.py:158-161); 0 = off
    force_octet: bool = False          # use the general octet kernel even where the tile kernel applies
    force_tile: bool = False           # use the tile kernel even where the lane kernel applies
    x: np.ndarray = field(default=None, repr=False)
    w: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.family not in ("gauss", "sinh"):
            raise ValueError("family must be 'gauss' or 'sinh'")
        if not 1 <= self.N <= _lib.MAX_NODES:
            raise ValueError(f"N must be in 1..{_lib.MAX_NODES}")
        if self.x is None or self.w is None:
            self.x, self.w = gauss_hermite(self.N)

    def for_dtype(self, dtype):
        """Convergence thresholds scaled to the arithmetic type."""
        if dtype == torch.float32:
            fields = {k: v for k, v in self.__dict__.items() if not k.startswith("_")}
            return FitConfig(**{**fields, "gtol1": max(self.gtol1, 1e-4), "gtol2": max(self.gtol2, 2e-5)})
        return self

    def descriptor(self):
        """``klhr_fit_t`` for the C ABI (cached: a sampler launches thousands of times with the same settings)."""
        key = (self.family, self.N, self.initscale, self.tol, self.scale_clip, self.n1, self.n2, self.nb, self.kmax,
               self.overrelax_K, self.fix_d, self.gtol1, self.gtol2, self.step_cap, self.c1, self.basin, self.grad_clip,
               self.force_octet, self.force_tile, self.x.tobytes(), self.w.tobytes())
        cached = self.__dict__.get("_desc_cache")
        if cached is not None and cached[0] == key:
            return cached[1]
        d = self._build_descriptor()
        self.__dict__["_desc_cache"] = (key, d)
        return d

    def _build_descriptor(self):
        d = _lib.FitDesc(family=_lib.FAMILY_GAUSS if self.family == "gauss" else _lib.FAMILY_SINH,
                         n_nodes=self.N, n1=self.n1, n2=self.n2, nb=self.nb,
                         initscale=self.initscale, tol=self.tol, scale_clip=self.scale_clip,
                         gtol1=self.gtol1, gtol2=self.gtol2, step_cap=self.step_cap, c1=self.c1,
                         basin=self.basin, grad_clip=float(self.grad_clip), flags=(1 if self.force_octet else 0) | (2 if self.fix_d else 0) | (4 if self.force_tile else 0), kmax=int(self.kmax),
                         overrelax_K=int(self.overrelax_K))
        for i in range(self.N):
            d.x[i] = float(self.x[i])
            d.w[i] = float(self.w[i])
        return d

    @property
    def n_eta(self):
        return 2 if self.family == "gauss" else 4


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _require_cuda(t, name):
    if not (torch.is_tensor(t) and t.is_cuda and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous CUDA tensor")


class Trace:
    """Per-draw trace buffers [S, B, ...] allocated on the device."""

    def __init__(self, S, B, D, n_eta, dtype, device, variates=False, rho=True, slice_cap=0):
        z = lambda *shape, dt=dtype: torch.zeros(*shape, dtype=dt, device=device)
        self.eta = z(S, B, n_eta)
        self.zp = z(S, B)
        self.r = z(S, B)
        self.accept = z(S, B, dt=torch.int32)
        self.evals = z(S, B, dt=torch.int32)
        self.rho = z(S, B, D) if rho else None
        self.z_init = z(S, B) if variates else None
        self.z_prop = z(S, B) if variates else None
        self.u = z(S, B) if variates else None
        self.init4 = z(S, B, 4) if variates and n_eta == 4 else None
        self.or_r = z(S, B, dt=torch.int32) if variates else None
        self.or_v = z(S, B) if variates else None
        self.slice_u = z(S, B, slice_cap) if slice_cap > 0 and variates else None
        self.slice_n = z(S, B, dt=torch.int32) if slice_cap > 0 else None

    def descriptor(self):
        return _lib.TraceDesc(eta=_ptr(self.eta), zp=_ptr(self.zp), r=_ptr(self.r), accept=_ptr(self.accept),
                              evals=_ptr(self.evals), rho=_ptr(self.rho), z_init=_ptr(self.z_init),
                              z_prop=_ptr(self.z_prop), u=_ptr(self.u), init4=_ptr(self.init4),
                              or_r=_ptr(self.or_r), or_v=_ptr(self.or_v), slice_u=_ptr(self.slice_u),
                              slice_n=_ptr(self.slice_n))


def step_replay(model: BSModel, fit: FitConfig, theta, rho, z_init, z_prop, u, init4=None, trace=True,
                or_r=None, or_v=None):
    """One draw for every chain with host-injected direction and variates; ``theta`` (B, D)
    is advanced IN PLACE.  Returns a ``Trace`` (S = 1) or None.  With ``fit.overrelax_K > 0`` the
    binomial counts ``or_r`` (int32) and beta variates ``or_v`` of the over-relaxed proposal are
    injected too."""
    lib = _lib.load()
    for t, n in ((theta, "theta"), (rho, "rho"), (z_init, "z_init"), (z_prop, "z_prop"), (u, "u")):
        _require_cuda(t, n)
    B, D = theta.shape
    if D != model.dim():
        raise ValueError("theta does not match model.dim()")
    dtype, dev = theta.dtype, theta.device
    if fit.family == "sinh":
        if init4 is None:
            raise ValueError("sinh family needs init4")
        _require_cuda(init4, "init4")
    if fit.overrelax_K > 0:
        if or_r is None or or_v is None:
            raise ValueError("over-relaxed replay needs or_r and or_v")
        _require_cuda(or_r, "or_r")
        _require_cuda(or_v, "or_v")
        trace = True
    tr = Trace(1, B, D, fit.n_eta, dtype, dev, rho=False) if trace else None
    trd = tr.descriptor() if tr else None
    if fit.overrelax_K > 0:
        trd.or_r, trd.or_v = or_r.data_ptr(), or_v.data_ptr()
    md, fd = model.descriptor(dtype, dev), fit.descriptor()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.klhr_step_replay(C.byref(md), C.byref(fd), _dtype_code(dtype), theta.data_ptr(),
                                        rho.data_ptr(), z_init.data_ptr(), _ptr(init4), z_prop.data_ptr(),
                                        u.data_ptr(), C.byref(trd) if trd else None, B, st),
                   "klhr_step_replay")
    return tr


def kl_eval(model: BSModel, fit: FitConfig, theta, rho, eta, hessian=False):
    """The reference's ``KL(eta, rho)`` (klhr.py:106-120 / klhr_sinh.py:163-176) for every chain: returns
    ``(f (B,), grad (B, n))`` in the reference's coordinates, plus the Hessian ``(B, n, n)`` on request."""
    lib = _lib.load()
    for t, n in ((theta, "theta"), (rho, "rho"), (eta, "eta")):
        _require_cuda(t, n)
    B, D = theta.shape
    if D != model.dim() or tuple(rho.shape) != (B, D) or tuple(eta.shape) != (B, fit.n_eta):
        raise ValueError("kl_eval: theta (B, D), rho (B, D), eta (B, n_eta) expected")
    dtype, dev = theta.dtype, theta.device
    f = torch.empty(B, dtype=dtype, device=dev)
    g = torch.empty(B, fit.n_eta, dtype=dtype, device=dev)
    H = torch.empty(B, fit.n_eta, fit.n_eta, dtype=dtype, device=dev) if hessian else None
    md, fd = model.descriptor(dtype, dev), fit.descriptor()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.klhr_kl_eval(C.byref(md), C.byref(fd), _dtype_code(dtype), theta.data_ptr(), rho.data_ptr(),
                                    eta.data_ptr(), f.data_ptr(), g.data_ptr(), _ptr(H), B, st), "klhr_kl_eval")
    return (f, g, H) if hessian else (f, g)


@dataclass
class Direction:
    """Device-resident direction law (reference ``_random_direction``, klhr.py:143-153)."""
    mean_cols: torch.Tensor = None     # (n_cols, D) or None (zero mean)
    sd: torch.Tensor = None            # (D,) sqrt(_cov) or None (ones)
    cdf: torch.Tensor = None           # (n_cols,) cumulative column probabilities
    n_zero_cols: int = 0               # trailing all-zero columns not stored in mean_cols (klhr.py:64-66)

    def descriptor(self):
        n = 0 if self.mean_cols is None else int(self.mean_cols.shape[0]) + int(self.n_zero_cols)
        return _lib.DirectionDesc(mean_cols=_ptr(self.mean_cols), sd=_ptr(self.sd), cdf=_ptr(self.cdf),
                                  n_cols=n, n_zero_cols=int(self.n_zero_cols))


def run(model: BSModel, fit: FitConfig, theta, n_steps, seed, direction: Direction = None,
        chain_offset=0, draw_offset=0, *, shift=None, pooled_s1=None, pooled_s2=None, chain_s1=None,
        chain_s2=None, accept_count=None, evals_total=None, draws=None, thin=1, thin_offset=0,
        skip_accum_last=False, trace: Trace = None):
    """``n_steps`` draws for every chain, in place on ``theta`` (B, D), asynchronous on the
    current stream.  All keyword tensors are optional device accumulators (see
    ``klhr_accum_t`` in include/klhr_sm100.h)."""
    lib = _lib.load()
    _require_cuda(theta, "theta")
    B, D = theta.shape
    if D != model.dim():
        raise ValueError("theta does not match model.dim()")
    dtype, dev = theta.dtype, theta.device
    md, fd = model.descriptor(dtype, dev), fit.descriptor()
    dd = (direction or Direction()).descriptor()
    ad = _lib.AccumDesc(shift=_ptr(shift), pooled_s1=_ptr(pooled_s1), pooled_s2=_ptr(pooled_s2),
                        chain_s1=_ptr(chain_s1), chain_s2=_ptr(chain_s2), accept_count=_ptr(accept_count),
                        evals_total=_ptr(evals_total), draws=_ptr(draws), thin=int(thin),
                        skip_accum_last=1 if skip_accum_last else 0, thin_offset=int(thin_offset))
    trd = trace.descriptor() if trace is not None else None
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.klhr_run(C.byref(md), C.byref(fd), C.byref(dd), _dtype_code(dtype), theta.data_ptr(),
                                B, int(chain_offset), int(draw_offset), int(n_steps),
                                C.c_uint64(int(seed) & (2 ** 64 - 1)), C.byref(ad),
                                C.byref(trd) if trd else None, st),
                   "klhr_run")


@dataclass
class SliceConfig:
    """Reference ``Slice`` constructor arguments that reach ``_uni_slice`` (slice.py:14-39); ``m = inf`` only
    (the configuration the reference itself can run).  ``cap``: columns of the shrinkage-uniform trace."""
    w: float = 1.0
    lower: float = -np.inf
    upper: float = np.inf
    tol: float = 1e-12
    cap: int = 24

    def __post_init__(self):
        if not (self.w > 0 and np.isfinite(self.w)):
            raise ValueError("w must be positive and finite")
        if not (self.lower <= 0.0 <= self.upper):
            raise ValueError("need lower <= 0 <= upper: the current point is line coordinate 0")

    def descriptor(self):
        return _lib.SliceDesc(w=float(self.w), lower=float(self.lower), upper=float(self.upper),
                              tol=float(self.tol), cap=int(self.cap), reserved=0)


def slice_replay(model: BSModel, cfg: SliceConfig, theta, rho, e, u0, shrink_u):
    """One ``Slice._uni_slice`` (slice.py:84-146) for every chain with host-injected direction and variates;
    ``theta`` (B, D) is advanced IN PLACE.  ``shrink_u`` (B, cap) holds the uniforms behind the shrinkage
    proposals, NaN-padded.  Returns a ``Trace`` (zp = accepted line coordinate, evals, slice_n)."""
    lib = _lib.load()
    for t, n in ((theta, "theta"), (rho, "rho"), (e, "e"), (u0, "u0"), (shrink_u, "shrink_u")):
        _require_cuda(t, n)
    B, D = theta.shape
    if D != model.dim():
        raise ValueError("theta does not match model.dim()")
    if tuple(shrink_u.shape) != (B, cfg.cap):
        raise ValueError("shrink_u must have shape (B, cfg.cap)")
    dtype, dev = theta.dtype, theta.device
    tr = Trace(1, B, D, 0, dtype, dev, rho=False, slice_cap=cfg.cap)
    trd = tr.descriptor()
    md, sd = model.descriptor(dtype, dev), cfg.descriptor()
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.klhr_slice_replay(C.byref(md), C.byref(sd), _dtype_code(dtype), theta.data_ptr(),
                                         rho.data_ptr(), e.data_ptr(), u0.data_ptr(), shrink_u.data_ptr(),
                                         C.byref(trd), B, st), "klhr_slice_replay")
    return tr


def slice_run(model: BSModel, cfg: SliceConfig, theta, n_steps, seed, direction: Direction = None,
              chain_offset=0, draw_offset=0, *, shift=None, chain_s1=None, chain_s2=None, accept_count=None,
              evals_total=None, draws=None, thin=1, thin_offset=0, trace: Trace = None):
    """``n_steps`` slice-sampling draws along random directions for every chain, in place on ``theta``."""
    lib = _lib.load()
    _require_cuda(theta, "theta")
    B, D = theta.shape
    if D != model.dim():
        raise ValueError("theta does not match model.dim()")
    dtype, dev = theta.dtype, theta.device
    md, sd = model.descriptor(dtype, dev), cfg.descriptor()
    dd = (direction or Direction()).descriptor()
    ad = _lib.AccumDesc(shift=_ptr(shift), chain_s1=_ptr(chain_s1), chain_s2=_ptr(chain_s2),
                        accept_count=_ptr(accept_count), evals_total=_ptr(evals_total), draws=_ptr(draws),
                        thin=int(thin), skip_accum_last=0, thin_offset=int(thin_offset))
    trd = trace.descriptor() if trace is not None else None
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.klhr_slice_run(C.byref(md), C.byref(sd), C.byref(dd), _dtype_code(dtype), theta.data_ptr(),
                                      B, int(chain_offset), int(draw_offset), int(n_steps),
                                      C.c_uint64(int(seed) & (2 ** 64 - 1)), C.byref(ad),
                                      C.byref(trd) if trd else None, st), "klhr_slice_run")


def outer_scratch(theta):
    """Zeroed scratch planes for ``outer_accumulate`` on this (B, D): one (D*D + D) plane per slice of 1024 chains."""
    B, D = theta.shape
    n = int(_lib.load().klhr_outer_scratch_doubles(B, D))
    return torch.zeros(max(n, 1), dtype=torch.float64, device=theta.device)


def outer_accumulate(theta, shift, outer, s1=None, scratch=None):
    """Pooled second moments of (theta - shift).  Without ``scratch``: outer (D, D) += sum_c u_c u_c^T and
    s1 (D,) += sum_c u_c, combined with fp64 atomics.  With ``scratch`` (``outer_scratch(theta)``): every slice of
    1024 chains adds into its own scratch plane and ``outer`` / ``s1`` are left alone until ``outer_reduce`` folds
    the planes (bit-reproducible and independent of the sharding of the chains over ranks)."""
    lib = _lib.load()
    _require_cuda(theta, "theta")
    B, D = theta.shape
    with torch.cuda.device(theta.device):
        st = torch.cuda.current_stream(theta.device).cuda_stream
        _lib.check(lib.klhr_outer_accumulate(_dtype_code(theta.dtype), theta.data_ptr(), _ptr(shift),
                                             outer.data_ptr(), _ptr(s1), B, D, _ptr(scratch),
                                             scratch.numel() if scratch is not None else 0, st),
                   "klhr_outer_accumulate")


def outer_reduce(scratch, outer, s1, B, D):
    """outer += tree-sum of the scratch planes, s1 += tree-sum of their first-moment rows; zeroes the planes."""
    lib = _lib.load()
    with torch.cuda.device(outer.device):
        st = torch.cuda.current_stream(outer.device).cuda_stream
        _lib.check(lib.klhr_outer_reduce(scratch.data_ptr(), scratch.numel(), outer.data_ptr(), _ptr(s1), int(B), int(D), st),
                   "klhr_outer_reduce")


def launch_info(model: BSModel, fit: FitConfig, dtype=torch.float64, free_running=True, accumulate=False,
                device=None):
    """(threads per CTA, dynamic shared bytes, registers per thread, resident CTAs per SM)."""
    lib = _lib.load()
    dev = torch.device(device) if device is not None else model.device
    md, fd = model.descriptor(dtype, dev), fit.descriptor()
    t, s, r = C.c_int32(), C.c_int32(), C.c_int32()
    with torch.cuda.device(dev):
        n = lib.klhr_launch_info(C.byref(md), C.byref(fd), _dtype_code(dtype), int(free_running),
                                 int(accumulate), C.byref(t), C.byref(s), C.byref(r))
    if n <= 0:
        raise _lib.KLHRLibraryError(f"klhr_launch_info failed ({n}): {_lib.last_error()}")
    return dict(threads=t.value, smem=s.value, regs=r.value, ctas_per_sm=n)


PROBE_KINDS = {"fp64_fma": 0, "normals": 1, "fp32_fma": 2, "cvt_f64_f32": 3, "mufu": 4, "mul_wide_u32": 5,
               "f2d_int_pipe": 6, "fp64_dmma": 7}


def peak_probe(kind, device, iters=4096, ctas_per_sm=8, repeats=5):
    """Operations per second of one peak micro-kernel (``klhr_peak_probe``, csrc/klhr_probe.cu), best of
    ``repeats`` launches timed with CUDA events on the current stream.  For "fp64_fma"/"fp32_fma" one operation
    is one FMA (2 flop); for "normals" one standard normal of the step kernels' direction stream."""
    lib = _lib.load()
    dev = torch.device(device)
    out = torch.zeros(8, dtype=torch.float64, device=dev)
    ctas = torch.cuda.get_device_properties(dev).multi_processor_count * int(ctas_per_sm)
    best = 0.0
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream(dev).cuda_stream
        for r in range(repeats + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops = lib.klhr_peak_probe(PROBE_KINDS[kind], int(iters), ctas, out.data_ptr(), st)
            e1.record()
            if ops < 0:
                raise _lib.KLHRLibraryError(f"klhr_peak_probe failed ({ops})")
            torch.cuda.synchronize(dev)
            if r > 0:                                   # first launch: module load + clock ramp
                best = max(best, ops / (e0.elapsed_time(e1) * 1e-3))
    return best
