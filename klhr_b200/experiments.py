"""Experiment drivers on the batched sampler -- the numeric content of the reference's four click
scripts, without the plots (matplotlib is not a dependency here).

    python -m klhr_b200.experiments accuracy   [-M -w --windowsize --windowscale -l -J -r -v -s -o -e1] ALGO
    python -m klhr_b200.experiments ar1        ...
    python -m klhr_b200.experiments funnel     ...
    python -m klhr_b200.experiments relaxation --data stan/earnings.json ...

Flags and defaults are those of reference ``experiment_accuracy.py:14-25``, ``experiment_ar1.py:16-28``,
``experiment_funnel.py:15-26`` and ``experiment_relaxationtime.py:14-26``; ``ALGO`` is ``klhr``, ``klhr_sinh``,
``sub_klhr_sinh``, ``slice`` or (accuracy only, as the comparison arm) ``mh``.  Added: ``--chains`` (every metric is
averaged over that many independent chains), ``--data`` (Stan JSON file), ``--seed``, ``--out`` (JSON file),
``--draws-out`` / ``--draws-chains`` (parquet dump of the draws, the output the reference left commented out).
Each command prints one JSON object with the quantities the reference plots or prints under ``-v``
(acceptance rate, MSJD, RMSE of running means / variances, posterior summaries, gradient evaluations).
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import click
import numpy as np
import torch

from . import BSModel, KLHR, KLHRSINH, MH, SUBKLHRSINH, Slice

ALGOS = {"klhr": KLHR, "klhr_sinh": KLHRSINH, "sub_klhr_sinh": SUBKLHRSINH, "slice": Slice}


def common(fn):
    opts = [
        click.option("-M", "--iterations", "M", type=int, default=1_000, help="number of iterations"),
        click.option("-w", "--warmup", "warmup", type=int, default=100, help="number of warmup iterations"),
        click.option("--windowsize", type=int, default=25),
        click.option("--windowscale", type=int, default=2),
        click.option("-l", "--amnesia", "l", type=int, default=2, help="kept for CLI compatibility (pooled PCA has no amnesia)"),
        click.option("-J", "J", type=int, default=2, help="number of eigenvectors"),
        click.option("-r", "--replication", "rep", type=int, default=0),
        click.option("-v", "--verbose", is_flag=True),
        click.option("-s", "--scale_dir_cov", is_flag=True),
        click.option("-o", "--overrelaxed", is_flag=True),
        click.option("-e1", "--eigen_method_one", is_flag=True),
        click.option("--chains", type=int, default=4096, help="independent chains advanced together"),
        click.option("--data", type=click.Path(), default=None, help="Stan JSON data file"),
        click.option("--seed", type=int, default=None),
        click.option("--out", type=click.Path(), default=None, help="write the JSON summary here too"),
        click.option("--draws-out", "draws_out", type=click.Path(), default=None,
                     help="write the draws as parquet (one column per parameter, plus chain and iteration)"),
        click.option("--draws-chains", "draws_chains", type=int, default=16, help="chains kept in --draws-out"),
        click.argument("algorithm", type=str),
    ]
    for o in reversed(opts):
        fn = o(fn)
    return fn


def make_sampler(algorithm, model, kw):
    if algorithm not in ALGOS:
        raise click.UsageError(f"Unknown algorithm {algorithm}; available: {', '.join(ALGOS)}")
    return ALGOS[algorithm](model, seed=kw["seed"], chains=kw["chains"], warmup=kw["warmup"],
                            windowsize=kw["windowsize"], windowscale=kw["windowscale"], l=kw["l"], J=kw["J"],
                            scale_dir_cov=kw["scale_dir_cov"], overrelaxed=kw["overrelaxed"],
                            eigen_method_one=kw["eigen_method_one"])


def load_data(path, default):
    return json.loads(Path(path).read_text()) if path else default


def running_rmse(draws, truth_mean=0.0, truth_var=1.0):
    """RMSE over coordinates of the running mean / running variance (ddof = 1) after each iteration,
    averaged over chains: experiment_accuracy.py:93-97 for a batch.  draws: (M, B, D)."""
    M = draws.shape[0]
    n = torch.arange(1, M + 1, device=draws.device, dtype=torch.float64)[:, None, None]
    c1 = torch.cumsum(draws.double(), 0)
    c2 = torch.cumsum(draws.double() ** 2, 0)
    mean = c1 / n
    var = (c2 - n * mean * mean) / torch.clamp(n - 1, min=1)
    var = torch.where(n > 2, var, torch.ones_like(var))                # onlinemoments.py:20-23
    rm = torch.sqrt(((mean - truth_mean) ** 2).mean(-1)).mean(-1)
    rv = torch.sqrt(((var - truth_var) ** 2).mean(-1)).mean(-1)
    return rm, rv


def msjd(draws):
    return float((draws[1:] - draws[:-1]).double().norm(dim=-1).mean())


def dump_draws(draws, model, kw):
    """The parquet dump the reference scripts left commented out (experiment_ar1.py:93-94)."""
    if kw.get("draws_out"):
        from .output import write_draws
        write_draws(kw["draws_out"], draws, model.parameter_names(), chains=kw["draws_chains"])


def emit(summary, kw):
    text = json.dumps(summary)
    print(text)
    if kw.get("out"):
        Path(kw["out"]).write_text(text + "\n")


def checkpoints(M):
    return sorted({min(M, k) for k in (10, 100, 1000, 10_000, 100_000, M)})


@click.group()
def cli():
    pass


@cli.command()
@common
def accuracy(**kw):
    """stan/normal.stan: RMSE of running mean / variance vs (0, 1), KLHR-family vs random-walk MH
    (reference experiment_accuracy.py)."""
    data = load_data(kw["data"], {"D": 2})
    model = BSModel(stan_file="stan/normal.stan", data=data)
    M = kw["M"]
    out = {"experiment": "accuracy", "model": "normal", "D": model.dim(), "M": M, "chains": kw["chains"]}
    arms = {"mh": MH(model, 0.09, seed=kw["seed"], chains=kw["chains"])}          # experiment_accuracy.py:69
    if kw["algorithm"] != "mh":
        arms[kw["algorithm"]] = make_sampler(kw["algorithm"], model, kw)
    for name, algo in arms.items():
        draws = algo.sample(M)
        draws = draws if torch.is_tensor(draws) else torch.as_tensor(draws)[:, None, :]
        rm, rv = running_rmse(draws)
        if name == kw["algorithm"]:
            dump_draws(draws, model, kw)
        lp = model.log_density(draws[-1].to(model.device))
        out[name] = {"acceptance": algo.acceptance_probability, "msjd": msjd(draws),
                     "rmse_mean": {str(k): float(rm[k - 1]) for k in checkpoints(M)},
                     "rmse_var": {str(k): float(rv[k - 1]) for k in checkpoints(M)},
                     "mean_log_density_last": float(lp.double().mean())}
    emit(out, kw)


@cli.command()
@common
def ar1(**kw):
    """stan/ar1.stan: RMSE of post-warm-up means and variances vs (0, 1) (reference experiment_ar1.py:96-99)."""
    data = load_data(kw["data"], {"N": 100})
    model = BSModel(stan_file="stan/ar1.stan", data=data)
    algo = make_sampler(kw["algorithm"], model, kw)
    draws = algo.sample(kw["M"])
    draws = draws if torch.is_tensor(draws) else torch.as_tensor(draws)[:, None, :]
    dump_draws(draws, model, kw)
    post = draws[kw["warmup"]:].double()
    v = post.var(0, unbiased=True)
    m = post.mean(0)
    emit({"experiment": "ar1", "D": model.dim(), "M": kw["M"], "chains": kw["chains"],
          "acceptance": algo.acceptance_probability,
          "rmse_mean": float(torch.sqrt((m ** 2).mean(-1)).mean()),
          "rmse_var": float(torch.sqrt(((v - 1) ** 2).mean(-1)).mean()),
          "max_abs_mean_pooled": float(m.mean(0).abs().max()), "min_var_pooled": float(post.reshape(-1, model.dim()).var(0).min())}, kw)


@cli.command()
@common
def funnel(**kw):
    """stan/funnel.stan: the first coordinate against N(0, 3^2) (reference experiment_funnel.py:66-70)."""
    data = load_data(kw["data"], {"D": 1})
    model = BSModel(stan_file="stan/funnel.stan", data=data)
    algo = make_sampler(kw["algorithm"], model, kw)
    draws = algo.sample(kw["M"])
    draws = draws if torch.is_tensor(draws) else torch.as_tensor(draws)[:, None, :]
    dump_draws(draws, model, kw)
    x = draws[kw["warmup"]:, :, 0].double().flatten()
    xs, _ = torch.sort(x)
    cdf = 0.5 * (1 + torch.erf(xs / (3 * math.sqrt(2))))
    grid = torch.arange(1, xs.numel() + 1, device=xs.device, dtype=torch.float64) / xs.numel()
    emit({"experiment": "funnel", "dims": model.dim(), "M": kw["M"], "chains": kw["chains"],
          "acceptance": algo.acceptance_probability, "x_mean": float(x.mean()), "x_sd": float(x.std()),
          "x_sd_truth": 3.0, "ks_distance_to_N(0,3)": float((cdf - grid).abs().max())}, kw)


@cli.command()
@common
def relaxation(**kw):
    """stan/earnings.stan: how fast the chains reach the typical set from the N(0, 0.1^2) start
    (reference experiment_relaxationtime.py; needs --data stan/earnings.json)."""
    if not kw["data"]:
        raise click.UsageError("relaxation needs --data <path to earnings.json>")
    model = BSModel(stan_file="stan/earnings.stan", data_file=kw["data"])
    algo = make_sampler(kw["algorithm"], model, kw)
    draws = algo.sample(kw["M"])
    draws = draws if torch.is_tensor(draws) else torch.as_tensor(draws)[:, None, :]
    dump_draws(draws, model, kw)
    M, B, D = draws.shape
    lp = model.log_density(draws.reshape(-1, D).to(model.device)).reshape(M, B).double()
    tail = lp[max(kw["warmup"], M // 2):]
    level = tail.median() - 3 * tail.std()
    reached = (lp >= level).double().argmax(0).double()                # first iteration at typical-set level
    post = draws[kw["warmup"]:].double().reshape(-1, D)
    emit({"experiment": "relaxation", "M": M, "chains": B, "acceptance": algo.acceptance_probability,
          "msjd": msjd(draws), "iterations_to_typical_set": {"median": float(reached.median()),
                                                             "p90": float(reached.quantile(0.9))},
          "constrained_mean": [float(v) for v in model.constrain(post).mean(0)],
          "constrained_sd": [float(v) for v in model.constrain(post).std(0)],
          "names": model.parameter_names(), "grad_evals_per_chain_draw": algo.grad_evals / (B * (M - 1))}, kw)


if __name__ == "__main__":
    cli()
