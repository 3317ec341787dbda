"""The experiment drivers (numeric content of reference experiment_*.py) run end to end on the device."""
import json

import numpy as np
import pytest
from click.testing import CliRunner

pytestmark = pytest.mark.gpu


def _run(args):
    from klhr_b200.experiments import cli
    res = CliRunner().invoke(cli, args, catch_exceptions=False)
    assert res.exit_code == 0, res.output
    return json.loads(res.output.strip().splitlines()[-1])


def test_accuracy_experiment_klhr_beats_random_walk():
    out = _run(["accuracy", "-M", "400", "-w", "100", "--chains", "512", "--seed", "1", "-e1", "klhr"])
    k, m = out["klhr"], out["mh"]
    assert k["acceptance"] > 0.999 and 0.9 < m["acceptance"] < 1.0         # stepsize 0.09: almost always accepted
    assert k["rmse_mean"]["400"] < k["rmse_mean"]["10"]
    assert k["rmse_mean"]["400"] < 0.5 * m["rmse_mean"]["400"]              # the reference's point: KLHR mixes far faster
    assert k["msjd"] > 5 * m["msjd"]


def test_ar1_and_funnel_experiments(tmp_path):
    a = _run(["ar1", "-M", "600", "-w", "200", "--chains", "256", "--seed", "2", "-e1", "--draws-out",
              str(tmp_path / "ar1.parquet"), "--draws-chains", "3", "klhr"])
    import pyarrow.parquet as pq
    t = pq.read_table(tmp_path / "ar1.parquet")
    assert t.num_rows == 600 * 3 and t.column_names[:3] == ["chain", "iteration", "y.1"] and t.num_columns == 102
    assert a["D"] == 100 and a["acceptance"] > 0.999 and a["max_abs_mean_pooled"] < 0.2
    f = _run(["funnel", "-M", "1500", "-w", "500", "--chains", "1024", "--seed", "3", "klhr_sinh"])
    assert abs(f["x_sd"] - 3.0) < 0.35 and abs(f["x_mean"]) < 0.3 and f["ks_distance_to_N(0,3)"] < 0.06
    s = _run(["funnel", "-M", "600", "-w", "300", "--chains", "256", "--seed", "4", "-o", "sub_klhr_sinh"])
    assert 0.8 < s["acceptance"] <= 1.0
    sl = _run(["funnel", "-M", "1500", "-w", "500", "--chains", "1024", "--seed", "5", "-e1", "slice"])
    assert sl["acceptance"] == 1.0 and abs(sl["x_sd"] - 3.0) < 0.35 and sl["ks_distance_to_N(0,3)"] < 0.06


def test_relaxation_experiment_on_synthetic_earnings(tmp_path):
    rng = np.random.default_rng(0)
    h = 66 + 4 * rng.normal(size=300)
    e = 20_000 + 1_500 * (h - 66) + 15_000 * rng.normal(size=300)
    p = tmp_path / "earnings.json"
    p.write_text(json.dumps({"N": 300, "earn": e.tolist(), "height": h.tolist()}))
    out = _run(["relaxation", "-M", "400", "-w", "200", "--windowsize", "50", "--chains", "256", "--seed", "5",
                "--data", str(p), "-e1", "klhr"])
    assert out["names"] == ["beta.1", "beta.2", "sigma", "s"]
    assert out["iterations_to_typical_set"]["median"] < 200               # reference PNG: ~60-70 iterations
    # sigma leaves its start value exp(0.1 z) ~ 1 for the dollar scale of the residuals within the run; the
    # hierarchical (beta, s) funnel of this model keeps beta shrunk (the reference chain does the same:
    # tests/golden/earnings_klhr_tight.npz sits at s ~ e^-48), so no OLS recovery is asserted
    assert 3_000 < out["constrained_mean"][2] < 40_000
    assert all(np.isfinite(out["constrained_mean"])) and out["grad_evals_per_chain_draw"] > 10
