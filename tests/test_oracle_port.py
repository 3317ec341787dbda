"""The single-chain port (oracle/ref_port.py) must reproduce tapes of the UNMODIFIED
reference bit for bit -- this is what pins the oracle."""
import numpy as np
import pytest

from oracle.bsmodel import BSModel
from oracle.ref_port import replay_port

CASES = [("normal_d2_klhr", 1100), ("normal_d2_klhr_method2", 400), ("illnormal_d100_klhr", 220),
         ("funnel_d2_klhr", 200), ("funnel_d2_sinh", 60), ("corrnormal_n50_klhr", 210),
         ("ar1_n100_klhr", 60), ("ark_t200_sinh", 40), ("rosenbrock_d4_sinh", 40), ("rosenbrock_d4_subsinh", 60),
         ("illnormal_d20_klhr_scaledir", 260), ("funnel_d2_sinh_scaledir_method1", 160)]


@pytest.mark.parametrize("name,n", CASES)
def test_port_is_bit_exact_on_reference_tape(tapes, name, n):
    t, meta, data = tapes(name)
    model = BSModel(stan_file=meta["model"] + ".stan", data=data)
    kw = {k: v for k, v in meta["ctor"].items() if k != "seed"}
    out = replay_port(t, model, meta["family"], n=n, **kw)
    if meta["family"] == "subsinh":                      # tape stores (m, log s, 0, e)
        out["eta"] = np.insert(out["eta"], 2, 0.0, axis=1)
    assert np.array_equal(out["eta"], t["eta"][:n])
    assert np.array_equal(out["zp"], t["zp"][:n])
    assert np.array_equal(out["r"], t["r"][:n], equal_nan=True)
    assert np.array_equal(out["accept"], t["accept"][:n])
    assert np.array_equal(out["theta"][:-1], t["theta0"][1:n])
    s = out["sampler"]
    k = int((t["closure_draw"] <= n).sum())
    if k:   # adaptation state after the last closure inside the replayed range
        assert np.array_equal(s.dir_mean, t["closure_mean"][k - 1])
        assert np.array_equal(s.dir_cov, t["closure_cov"][k - 1])
        assert np.array_equal(s.eigvecs, t["closure_eigvecs"][k - 1])
        assert np.array_equal(s.eigvals, t["closure_eigvals"][k - 1])


@pytest.mark.parametrize("name", ["freerun_funnel_d2_klhr_adapt", "freerun_funnel_d2_klhr_overrelaxed",
                                  "freerun_funnel_d2_sinh_overrelaxed", "freerun_illnormal_d20_klhr_adapt"])
def test_port_reproduces_seeded_free_runs_of_the_reference(name):
    """No tape RNG here: the unmodified reference ran from a seed with its own PCG64 stream,
    multivariate_normal, window adaptation and (where enabled) over-relaxed proposals drawing from
    SciPy's global RNG; the port reproduces the whole trajectory bit for bit."""
    import json
    from conftest import GOLDEN
    from oracle.ref_port import ChainSampler
    t = dict(np.load(GOLDEN / f"{name}.npz"))
    meta, data = json.loads(str(t["meta_json"])), json.loads(str(t["data_json"]))
    model = BSModel(stan_file=meta["model"] + ".stan", data=data)
    s = ChainSampler(model, family=meta["family"], **meta["ctor"])
    np.random.seed(meta["ctor"]["seed"])
    out = np.array([s.draw() for _ in range(meta["draws"])])
    assert np.array_equal(out, t["thetas"])
    assert float(np.ravel(s.acceptance_probability)[0]) == float(t["acceptance_probability"])
    assert s.grad_evals == int(t["grad_evals"])


@pytest.mark.parametrize("name", ["freerun_normal_d2_mh", "freerun_funnel_d2_mh"])
def test_mh_port_reproduces_reference(name):
    import json
    from conftest import GOLDEN
    from oracle.ref_port import MHChain
    t = dict(np.load(GOLDEN / f"{name}.npz"))
    meta, data = json.loads(str(t["meta_json"])), json.loads(str(t["data_json"]))
    s = MHChain(BSModel(stan_file=meta["model"] + ".stan", data=data), meta["ctor"]["stepsize"], seed=meta["ctor"]["seed"])
    out = np.array([s.draw() for _ in range(meta["draws"])])
    assert np.array_equal(out, t["thetas"])
    assert float(np.ravel(s.acceptance_probability)[0]) == float(t["acceptance_probability"])


SLICE_TAPES = ["slice_normal_d2", "slice_funnel_d2", "slice_funnel_d11", "slice_illnormal_d100",
               "slice_rosenbrock_d4_w3", "slice_ark_t200_method2"]


@pytest.mark.parametrize("name", SLICE_TAPES)
def test_slice_port_is_bit_exact_on_reference_tape(tapes, name):
    """Tape of the unmodified reference ``Slice`` (slice.py) replayed through ``SliceChain``: accepted line
    coordinates, trajectory, value-call counts and adaptation state agree to the last bit."""
    from oracle.ref_port import replay_slice_port
    t, meta, data = tapes(name)
    model = BSModel(stan_file=meta["model"] + ".stan", data=data)
    kw = {k: v for k, v in meta["ctor"].items() if k != "seed"}
    n = min(400, len(t["e"]))
    out = replay_slice_port(t, model, n=n, **kw)
    assert np.array_equal(out["x1"], t["x1"][:n])
    assert np.array_equal(out["n_shrink"], t["n_shrink"][:n])
    assert np.array_equal(out["evals"], t["evals"][:n])
    assert np.array_equal(out["theta"][:-1], t["theta0"][1:n])
    s = out["sampler"]
    k = int((t["closure_draw"] <= n).sum())
    assert k >= 1
    assert np.array_equal(s.dir_mean, t["closure_mean"][k - 1])
    assert np.array_equal(s.dir_cov, t["closure_cov"][k - 1])
    assert np.array_equal(s.eigvecs, t["closure_eigvecs"][k - 1])
    assert np.array_equal(s.eigvals, t["closure_eigvals"][k - 1])


@pytest.mark.parametrize("name", ["freerun_funnel_d2_slice_adapt", "freerun_illnormal_d20_slice_adapt"])
def test_slice_port_reproduces_seeded_free_runs_of_the_reference(name):
    import json
    from conftest import GOLDEN
    from oracle.ref_port import SliceChain
    t = dict(np.load(GOLDEN / f"{name}.npz"))
    meta, data = json.loads(str(t["meta_json"])), json.loads(str(t["data_json"]))
    s = SliceChain(BSModel(stan_file=meta["model"] + ".stan", data=data), **meta["ctor"])
    out = np.array([s.draw() for _ in range(meta["draws"])])
    assert np.array_equal(out, t["thetas"])
    assert float(s.acceptance_probability) == float(t["acceptance_probability"]) == 1.0


def test_quadrature_table(tapes):
    # values after the reference's normalisation, klhr.py:46-49
    from oracle.ref_port import gauss_hermite_probabilists
    x, w = gauss_hermite_probabilists(8)
    t, _, _ = tapes("normal_d2_klhr")
    assert np.array_equal(x, t["x_nodes"]) and np.array_equal(w, t["w_nodes"])
    assert np.isclose(w.sum(), 1) and np.isclose((w * x ** 2).sum(), 1) and np.isclose((w * x ** 4).sum(), 3)
    assert np.allclose(np.abs(x[4:]), [0.5390798113513752, 1.6365190424351082, 2.802485861287542, 4.1445471861258945])
