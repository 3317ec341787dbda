"""Generate tests/golden/*.npz from the UNMODIFIED reference -- authoring container only.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run as

    python oracle/make_golden.py [--reference /root/reference] [--out tests/golden]

It imports ``/root/reference/klhr.py`` and ``klhr_sinh.py`` untouched, with this
directory's ``bsmodel.py`` shim ahead of the reference on ``sys.path`` (BridgeStan is not
installable here), and swaps each sampler's public ``rng`` attribute (reference
``mcmc.py:9-12``) for a TAPE generator that records every variate the sampler consumes.
Per ``draw()`` (reference ``klhr.py:196-223`` / ``klhr_sinh.py:262-289``) the tape holds

    theta0, rho, z_init, [init4], z_prop, u          inputs of the step
    eta, zp, r, accept, theta1                       what the reference computed
    grad_evals_draw                                  nfev1 + N * nfev2 (klhr.py:132,140)

plus the adaptation state (``_mean``, ``_cov``, ``_eigvecs``, ``_eigvals``) after every
window closure.  ``multivariate_normal`` is served as ``m + sqrt(diag S) * z`` (same law;
NumPy's own routine routes z through an SVD permutation, SURVEY.md section 8a H3, which
is why replay injects rho and never z).

Nothing here is read on the GPU box: the fixtures are committed.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


class TapeRNG:
    """Generator look-alike recording the standard variates behind each call."""

    def __init__(self, seed):
        self._g = np.random.default_rng(seed)
        self.log = []  # (kind, payload)

    # reference call sites: klhr.py:91,129 ; klhr_sinh.py:191 ; klhr.py:180
    def normal(self, loc=0.0, scale=1.0, size=None):
        z = self._g.standard_normal(size)
        self.log.append(("normal", np.array(z, dtype=np.float64, copy=True)))
        return loc + scale * z

    # reference call site: klhr.py:188
    def uniform(self, low=0.0, high=1.0, size=None):
        u = self._g.random(size)
        self.log.append(("uniform", np.array(u, dtype=np.float64, copy=True)))
        self.last_uniform_value = low + (high - low) * u
        return self.last_uniform_value

    # reference call site: slice.py:90
    def exponential(self):
        e = self._g.standard_exponential()
        self.log.append(("exponential", np.array(e, dtype=np.float64)))
        return e

    # reference call site: klhr.py:147 ; searchsorted(cumsum(p), u, 'right') is what NumPy does
    def choice(self, a, p=None):
        u = self._g.random()
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        j = int(np.searchsorted(cdf, u, side="right"))
        self.log.append(("choice", np.array([u, j], dtype=np.float64)))
        return j

    # reference call site: klhr.py:152
    def multivariate_normal(self, mean, cov):
        z = self._g.standard_normal(np.size(mean))
        x = np.asarray(mean, dtype=np.float64) + np.sqrt(np.diag(cov)) * z
        self.log.append(("mvn", z.copy()))
        return x


def _tape_run(algo, M, family):
    """Drive ``algo.draw()`` M times, capturing inputs/outputs of every step."""
    rec = {k: [] for k in ("theta0", "rho", "z_init", "z_prop", "u", "eta", "zp", "r",
                           "accept", "theta1", "grad_evals_draw", "ujdir")}
    if family in ("sinh", "subsinh"):
        rec["init4"] = []
    closures = {"draw": [], "mean": [], "cov": [], "eigvecs": [], "eigvals": []}
    cur = {}

    orig_fit = algo.fit
    orig_mh = algo._metropolis_step
    logq = algo._logq if family == "gauss" else algo._log_q

    def fit(rho):
        cur["rho"] = np.array(rho, copy=True)
        ge0 = algo.grad_evals
        n0 = len(algo.rng.log)
        eta = eta_ret = orig_fit(rho)
        # variates consumed inside fit: 1 normal (stage-1 start) [+ normal(size=4) for sinh]
        used = algo.rng.log[n0:]
        cur["z_init"] = float(used[0][1])
        if family == "sinh":
            cur["init4"] = np.array(used[1][1], copy=True)
        if family == "subsinh":                       # normal(size=3): (.., .., e start); padded to 4
            cur["init4"] = np.append(np.array(used[1][1], copy=True), 0.0)
            eta = np.array([eta[0], eta[1], 0.0, eta[2]])      # stored as (m, log s, log d = 0, e)
        cur["eta"] = np.array(eta, copy=True)
        cur["nfev"] = algo.grad_evals - ge0
        return eta_ret

    def mh(eta, rho):
        theta0 = np.array(algo.theta, copy=True)
        n0 = len(algo.rng.log)
        out = orig_mh(eta, rho)
        used = algo.rng.log[n0:]
        z = float(np.ravel(used[0][1])[0])
        u = float(used[1][1])
        if family == "gauss":
            m, s = algo._unpack(eta)
            zp = m + s * z
        else:
            zp = float(np.ravel(algo._T(np.array([z]), eta))[0])
        thetap = zp * rho + theta0
        r = algo.model.log_density(thetap) - algo.model.log_density(theta0)
        with np.errstate(all="ignore"):
            r = r + float(np.ravel(logq(0, eta))[0]) - float(np.ravel(logq(zp, eta))[0])
        cur.update(theta0=theta0, z_prop=z, u=u, zp=zp, r=r,
                   accept=bool(np.log(u) < np.minimum(0, r)),
                   theta1=np.array(out, copy=True))
        return out

    algo.fit = fit
    algo._metropolis_step = mh

    for m in range(M):
        n0 = len(algo.rng.log)
        wa = algo._windowedadaptation
        will_close = (wa._warmup >= wa._windowsize and wa._num_windows > 0
                      and (algo._draw + 1) == wa._closures[wa._idx])
        algo.draw()
        first = algo.rng.log[n0]
        rec["ujdir"].append(float(first[1][0]) if first[0] == "choice" else -1.0)
        for k in rec:
            if k in ("grad_evals_draw", "ujdir"):
                continue
            rec[k].append(cur[k])
        rec["grad_evals_draw"].append(cur["nfev"])
        assert cur["accept"] == (not np.array_equal(cur["theta0"], cur["theta1"])) or cur["zp"] == 0
        if will_close:
            closures["draw"].append(algo._draw)
            closures["mean"].append(np.array(algo._mean, copy=True))
            closures["cov"].append(np.array(algo._cov, copy=True))
            closures["eigvecs"].append(np.array(algo._eigvecs, copy=True))
            closures["eigvals"].append(np.array(algo._eigvals, copy=True))

    out = {k: np.array(v) for k, v in rec.items()}
    # theta1[m] == theta0[m+1]; keep only the last one to halve the fixture size
    out["theta_last"] = out.pop("theta1")[-1]
    for k, v in closures.items():
        out["closure_" + k] = np.array(v)
    out["acceptance_probability"] = np.array(float(np.ravel(algo.acceptance_probability)[0]))
    out["grad_evals"] = np.array(int(algo.grad_evals))
    return out


SLICE_MAX_SHRINK = 24


def _tape_slice(algo, M):
    """Tape of the reference ``Slice`` (slice.py): per draw theta0, rho, the exponential e, the interval
    uniform u0, the shrinkage uniforms (padded with NaN), the accepted line coordinate x1 and the number
    of model value calls; adaptation state after every closure."""
    rec = {k: [] for k in ("theta0", "rho", "e", "u0", "shrink_u", "n_shrink", "x1", "evals", "ujdir")}
    closures = {"draw": [], "mean": [], "cov": [], "eigvecs": [], "eigvals": []}
    cur = {}
    orig = algo._uni_slice

    def uni(rho):
        n0 = len(algo.rng.log)
        c0 = algo.model.n_value_calls
        theta0 = np.array(algo.theta, copy=True)
        out = orig(rho)
        used = algo.rng.log[n0:]
        assert used[0][0] == "exponential" and all(k == "uniform" for k, _ in used[1:])
        su = np.full(SLICE_MAX_SHRINK, np.nan)
        sh = [float(v) for _, v in used[2:]]
        assert len(sh) <= SLICE_MAX_SHRINK
        su[:len(sh)] = sh
        cur.update(theta0=theta0, rho=np.array(rho, copy=True), e=float(used[0][1]), u0=float(used[1][1]),
                   shrink_u=su, n_shrink=len(sh), x1=float(algo.rng.last_uniform_value),
                   evals=algo.model.n_value_calls - c0)
        return out

    algo._uni_slice = uni
    for m in range(M):
        n0 = len(algo.rng.log)
        wa = algo._windowedadaptation
        will_close = (wa._warmup >= wa._windowsize and wa._num_windows > 0
                      and (algo._draw + 1) == wa._closures[wa._idx])
        algo.draw()
        first = algo.rng.log[n0]
        rec["ujdir"].append(float(first[1][0]) if first[0] == "choice" else -1.0)
        for k in rec:
            if k != "ujdir":
                rec[k].append(cur[k])
        if will_close:
            closures["draw"].append(algo._draw)
            closures["mean"].append(np.array(algo._mean, copy=True))
            closures["cov"].append(np.array(algo._cov, copy=True))
            closures["eigvecs"].append(np.array(algo._eigvecs, copy=True))
            closures["eigvals"].append(np.array(algo._eigvals, copy=True))
    out = {k: np.array(v) for k, v in rec.items()}
    out["theta_last"] = np.array(algo.theta, copy=True)
    for k, v in closures.items():
        out["closure_" + k] = np.array(v)
    out["acceptance_probability"] = np.array(float(algo.acceptance_probability))
    return out


# reference slice.py (N3): name, stan stem, data, draws, ctor kwargs
SLICE_CASES = [
    ("slice_normal_d2", "normal", {"D": 2}, 2_000, dict(seed=121)),
    ("slice_funnel_d2", "funnel", {"D": 1}, 2_000, dict(seed=122)),
    ("slice_funnel_d11", "funnel", {"D": 10}, 1_000, dict(seed=123, warmup=400)),
    ("slice_illnormal_d100", "ill-normal", {"D": 100}, 300, dict(seed=124, warmup=200)),
    ("slice_rosenbrock_d4_w3", "rosenbrock", {"D": 2}, 1_000, dict(seed=125, w=3.0)),
    ("slice_ark_t200_method2", "arK", "arK.json", 500, dict(seed=126, eigen_method_one=False, warmup=200)),
    ("stats_slice_funnel_d2_noadapt", "funnel", {"D": 1}, 30_000, dict(seed=127, warmup=0)),
]

CASES = [
    # name, stan stem, data, family, draws, ctor kwargs, tighten-gtol
    ("normal_d2_klhr", "normal", {"D": 2}, "gauss", 10_000, dict(seed=20261018), None),
    ("illnormal_d100_klhr", "ill-normal", {"D": 100}, "gauss", 400, dict(seed=11, warmup=200), None),
    ("illnormal_d100_klhr_tight", "ill-normal", {"D": 100}, "gauss", 200, dict(seed=12, warmup=100), 1e-12),
    ("funnel_d2_klhr", "funnel", {"D": 1}, "gauss", 3_000, dict(seed=21), None),
    ("funnel_d2_klhr_tight", "funnel", {"D": 1}, "gauss", 2_000, dict(seed=22), 1e-10),
    ("funnel_d11_klhr_tight", "funnel", {"D": 10}, "gauss", 1_000, dict(seed=23), 1e-10),
    ("funnel_d2_sinh", "funnel", {"D": 1}, "sinh", 2_000, dict(seed=31, overrelaxed=False), None),
    ("funnel_d2_sinh_tight", "funnel", {"D": 1}, "sinh", 1_500, dict(seed=32, overrelaxed=False), 1e-9),
    ("corrnormal_n50_klhr", "corr-normal", {"N": 50, "rho": 0.9}, "gauss", 400, dict(seed=41, warmup=200), None),
    ("ar1_n100_klhr", "ar1", {"N": 100}, "gauss", 400, dict(seed=51, warmup=200), None),
    ("ark_t200_klhr_tight", "arK", "arK.json", "gauss", 1_000, dict(seed=61), 1e-10),
    ("ark_t200_sinh", "arK", "arK.json", "sinh", 500, dict(seed=62, overrelaxed=False), None),
    ("rosenbrock_d4_klhr_tight", "rosenbrock", {"D": 2}, "gauss", 2_000, dict(seed=71), 1e-10),
    ("rosenbrock_d4_sinh", "rosenbrock", {"D": 2}, "sinh", 1_000, dict(seed=72, overrelaxed=False), None),
    ("normal_d2_klhr_method2", "normal", {"D": 2}, "gauss", 1_500,
     dict(seed=81, eigen_method_one=False), None),
    # direction covariance scaled by the gradient variances (klhr.py:205-206) and the method-two direction mean
    ("illnormal_d20_klhr_scaledir", "ill-normal", {"D": 20}, "gauss", 260,
     dict(seed=131, warmup=200, scale_dir_cov=True), None),
    ("funnel_d2_sinh_scaledir_method1", "funnel", {"D": 1}, "sinh", 160,
     dict(seed=132, warmup=100, scale_dir_cov=True, eigen_method_one=True, overrelaxed=False), None),
    # 3-parameter sinh-arcsinh variant (reference sub_klhr_sinh.py)
    ("funnel_d2_subsinh_tight", "funnel", {"D": 1}, "subsinh", 1_000, dict(seed=111, overrelaxed=False), 1e-9),
    ("rosenbrock_d4_subsinh", "rosenbrock", {"D": 2}, "subsinh", 400, dict(seed=112, overrelaxed=False), None),
    # the model of the reference's own self-tests (klhr.py:238) and relaxation experiment
    ("earnings_klhr_tight", "earnings", "earnings.json", "gauss", 500, dict(seed=101, warmup=200), 1e-10),
    ("earnings_sinh", "earnings", "earnings.json", "sinh", 250, dict(seed=102, warmup=100, overrelaxed=False), None),
    # long runs WITHOUT adaptation (warmup=0 -> isotropic direction law, identical on both
    # sides): only accept flags and thinned states are kept ("stats" tapes) for the
    # acceptance-rate and posterior parity tests.
    ("stats_funnel_d2_klhr_noadapt", "funnel", {"D": 1}, "gauss", 30_000, dict(seed=91, warmup=0), None),
    ("stats_funnel_d2_sinh_noadapt", "funnel", {"D": 1}, "sinh", 6_000,
     dict(seed=92, warmup=0, overrelaxed=False), None),
    ("stats_rosenbrock_d4_klhr_noadapt", "rosenbrock", {"D": 2}, "gauss", 20_000, dict(seed=93, warmup=0), None),
    # over-relaxed proposals (klhr.py:160-173; default ON in KLHRSINH, klhr_sinh.py:30): r and v come from
    # SciPy's global RNG, seeded here with np.random.seed(case seed) so the runs are reproducible
    ("stats_funnel_d2_klhr_overrelaxed", "funnel", {"D": 1}, "gauss", 20_000,
     dict(seed=94, warmup=0, overrelaxed=True), None),
    ("stats_funnel_d2_sinh_overrelaxed", "funnel", {"D": 1}, "sinh", 5_000,
     dict(seed=95, warmup=0, overrelaxed=True), None),
]

# seeded free runs of the unmodified reference WITHOUT the tape RNG (its own PCG64 stream, its own
# multivariate_normal): the port must reproduce the whole trajectory bit for bit from the seed
FREERUNS = [
    ("freerun_funnel_d2_klhr_adapt", "funnel", {"D": 1}, "gauss", 160, dict(seed=3, warmup=100)),
    ("freerun_funnel_d2_klhr_overrelaxed", "funnel", {"D": 1}, "gauss", 150, dict(seed=4, warmup=0, overrelaxed=True)),
    ("freerun_funnel_d2_sinh_overrelaxed", "funnel", {"D": 1}, "sinh", 120, dict(seed=5, warmup=0, overrelaxed=True)),
    ("freerun_illnormal_d20_klhr_adapt", "ill-normal", {"D": 20}, "gauss", 130, dict(seed=6, warmup=100)),
    # random-walk Metropolis (reference mh.py), stepsize of experiment_accuracy.py:69
    ("freerun_normal_d2_mh", "normal", {"D": 2}, "mh", 3000, dict(seed=7, stepsize=0.09)),
    ("freerun_funnel_d2_mh", "funnel", {"D": 1}, "mh", 3000, dict(seed=8, stepsize=0.9)),
    # slice sampling along adapted directions (reference slice.py)
    ("freerun_funnel_d2_slice_adapt", "funnel", {"D": 1}, "slice", 400, dict(seed=9, warmup=200)),
    ("freerun_illnormal_d20_slice_adapt", "ill-normal", {"D": 20}, "slice", 250, dict(seed=10, warmup=200)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=str(HERE.parent / "tests" / "golden"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()

    ref = Path(args.reference)
    sys.path.insert(0, str(ref))
    sys.path.insert(0, str(HERE))  # bare ``bsmodel`` -> the shim, ahead of the reference's

    import functools
    import bsmodel as shim
    import klhr as ref_klhr
    import klhr_sinh as ref_sinh
    import sub_klhr_sinh as ref_sub
    import scipy.optimize

    out_dir = Path(args.out)
    out_dir.mkdir(parents=True, exist_ok=True)

    for name, stem, data, family, M, kw in FREERUNS:
        if args.only and args.only not in name:
            continue
        model = shim.BSModel(stan_file=f"stan/{stem}.stan", data=data)
        if family == "mh":
            import mh as ref_mh
            algo = ref_mh.MH(model, kw["stepsize"], seed=kw["seed"])
        elif family == "slice":
            import slice as ref_slice
            algo = ref_slice.Slice(model, **kw)
        else:
            cls = ref_klhr.KLHR if family == "gauss" else ref_sinh.KLHRSINH
            (ref_klhr if family == "gauss" else ref_sinh).minimize = scipy.optimize.minimize
            algo = cls(model, **kw)
        np.random.seed(kw["seed"])                      # SciPy's global RNG (over-relaxed proposals)
        thetas = np.array([algo.draw() for _ in range(M)])
        meta = dict(case=name, model=stem, family=family, draws=M, ctor=kw, numpy=np.__version__,
                    scipy=scipy.__version__)
        np.savez_compressed(out_dir / f"{name}.npz", thetas=thetas,
                            acceptance_probability=np.array(float(np.ravel(algo.acceptance_probability)[0])),
                            grad_evals=np.array(int(getattr(algo, "grad_evals", 0))), meta_json=np.array(json.dumps(meta)),
                            data_json=np.array(json.dumps(data)))
        print(f"{name:36s} draws={M:6d} acc={float(np.ravel(algo.acceptance_probability)[0]):.3f}")

    for name, stem, data, M, kw in SLICE_CASES:
        if args.only and args.only not in name:
            continue
        import slice as ref_slice
        if isinstance(data, str):
            data = json.loads((ref / "stan" / data).read_text())
        model = shim.BSModel(stan_file=f"stan/{stem}.stan", data=data)
        kw = dict(kw)
        seed = kw.pop("seed")
        algo = ref_slice.Slice(model, seed=seed, **kw)
        algo.rng = TapeRNG(seed + 1000)
        tape = _tape_slice(algo, M)
        if name.startswith("stats_"):
            tape = {"theta_thin10": tape["theta0"][::10], "theta_last": tape["theta_last"],
                    "evals_mean": np.array(tape["evals"].mean()), "n_shrink_mean": np.array(tape["n_shrink"].mean()),
                    "acceptance_probability": tape["acceptance_probability"]}
        meta = dict(case=name, model=stem, family="slice", draws=M, ctor=dict(kw, seed=seed), numpy=np.__version__,
                    w=float(algo.w), tol=float(algo._tol), initscale=float(algo._initscale), J=int(algo.J),
                    eigen_method_one=bool(algo._eigen_method_one),
                    closures=list(algo._windowedadaptation._closures))
        tape["meta_json"] = np.array(json.dumps(meta))
        tape["data_json"] = np.array(json.dumps(data))
        np.savez_compressed(out_dir / f"{name}.npz", **tape)
        print(f"{name:32s} draws={M:6d} evals/draw={float(np.mean(tape.get('evals', tape.get('evals_mean')))):.2f}")

    for name, stem, data, family, M, kw, gtol in CASES:
        if args.only and args.only not in name:
            continue
        if isinstance(data, str):
            data = json.loads((ref / "stan" / data).read_text())
            data = {k: v for k, v in data.items() if k != "male"}     # unused by stan/earnings.stan
        model = shim.BSModel(stan_file=f"stan/{stem}.stan", data=data)
        mod = {"gauss": ref_klhr, "sinh": ref_sinh, "subsinh": ref_sub}[family]
        # "tight" tapes: same reference code, SciPy asked for a smaller gtol by rebinding
        # the module-level name the reference calls (klhr.py:5); sinh passes its own gtol
        # through options (klhr_sinh.py:199) so it is overridden there.
        if gtol is not None:
            def tight(fun, x0, gtol=gtol, **k):
                opts = dict(k.pop("options", None) or {})
                opts["gtol"] = gtol
                return scipy.optimize.minimize(fun, x0, options=opts, **k)
            mod.minimize = tight
        else:
            mod.minimize = scipy.optimize.minimize
        cls = {"gauss": ref_klhr.KLHR, "sinh": ref_sinh.KLHRSINH, "subsinh": ref_sub.SUBKLHRSINH}[family]
        seed = kw.pop("seed")
        kw = dict(kw)
        # construct with a plain seed (the ctor draws the start point), then tape
        algo = cls(model, seed=seed, **kw)
        np.random.seed(seed)                            # SciPy's global RNG (over-relaxed proposals)
        if name.startswith("stats_") and kw.get("overrelaxed"):
            # plain seeded run (no tape RNG: the over-relaxed step consumes no proposal normal)
            acc, th = [], []
            for _ in range(M):
                th.append(np.array(algo.theta, copy=True))
                prev = float(np.ravel(algo.acceptance_probability)[0])
                algo.draw()
                new = float(np.ravel(algo.acceptance_probability)[0])
                # the MH flag itself (klhr.py:188-193): an over-relaxed proposal can land exactly on the
                # current point (r == K - r), which is an ACCEPTED move that leaves theta unchanged
                acc.append(bool(round(prev + (new - prev) * algo._draw)))
            tape = {"accept": np.array(acc), "theta0": np.array(th), "theta_last": np.array(algo.theta),
                    "acceptance_probability": np.array(float(np.ravel(algo.acceptance_probability)[0])),
                    "grad_evals": np.array(int(algo.grad_evals))}
        else:
            algo.rng = TapeRNG(seed + 1000)
            tape = _tape_run(algo, M, family)
        if name.startswith("stats_"):
            tape = {"accept": tape["accept"], "theta_thin10": tape["theta0"][::10],
                    "acceptance_probability": tape["acceptance_probability"],
                    "grad_evals": tape["grad_evals"], "theta_last": tape["theta_last"]}
        tape["x_nodes"] = np.array(algo.x)
        tape["w_nodes"] = np.array(algo.w)
        meta = dict(case=name, model=stem, family=family, draws=M, ctor=dict(kw, seed=seed),
                    gtol=gtol, numpy=np.__version__, scipy=scipy.__version__,
                    tol=float(algo._tol), scale_clip=float(algo._scale_clip),
                    initscale=float(algo._initscale), J=int(algo.J),
                    eigen_method_one=bool(algo._eigen_method_one),
                    closures=list(algo._windowedadaptation._closures))
        dd = {k: v for k, v in data.items()}
        tape["meta_json"] = np.array(json.dumps(meta))
        tape["data_json"] = np.array(json.dumps(dd))
        np.savez_compressed(out_dir / f"{name}.npz", **tape)
        acc = float(tape["acceptance_probability"])
        print(f"{name:32s} draws={M:6d} acc={acc:.3f} grad_evals/draw={int(tape['grad_evals']) / M:.1f}")


if __name__ == "__main__":
    main()
