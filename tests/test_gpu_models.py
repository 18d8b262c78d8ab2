"""GPU parity on the reference's remaining test fixtures (tests/test_data.js, tests/test_mcmc_js.R): every stepper kind,
draw for draw against the oracle under the matched Philox stream. These models contain no factorised plate, so the
comparison is bit-exact -- including the int stepper's accept decisions and the binary stepper's choices."""
import numpy as np
import pytest

import models

pytestmark = pytest.mark.gpu


def _run(gpu_pkg, orc, log_post, c_model, params, data_py, data_c, chains, seed, burn, sample, options=None, comp_options=None, thin=1):
    opts = {"chains": chains, "seed": seed, "thin": thin}
    opts.update(options or {})
    s = gpu_pkg.mcmc.AmwgSampler(params, log_post, data_py, opts)
    s.burn(burn)
    got = s.sample(sample)
    ref = orc.run_model(c_model, data_c, params, chains=chains, seed=seed, burn=burn, sample=sample, thin=thin, comp_options=comp_options)
    return s, got, ref


def _same(a, b):
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def test_real_metropolis_stepper(gpu_pkg, orc):
    """tests/test_mcmc_js.R:55-65: one real parameter, target N(10, 5)."""
    s, got, ref = _run(gpu_pkg, orc, models.norm_dens(gpu_pkg.ld), "norm_dens", {"x": {"type": "real"}}, None, None, 256, 1, 300, 200)
    assert _same(got["x"], ref["x"])
    assert abs(got["x"].mean() - 10) < 0.5


def test_int_metropolis_stepper_accept_indices(gpu_pkg, orc):
    """tests/test_mcmc_js.R:67-81: int parameter with lower bound 0, target Pois(10). The int stepper's proposals
    (Math.round of rnorm), bounds rejections and accept decisions are bit-equal to the oracle: the accept indices
    (sweeps at which the state changed) are identical for every chain."""
    s, got, ref = _run(gpu_pkg, orc, models.poisson_dens(gpu_pkg.ld), "poisson_dens", {"x": {"type": "int", "lower": 0}}, None, None, 512, 2, 100, 400)
    assert _same(got["x"], ref["x"])
    moved_gpu, moved_ref = np.diff(got["x"], axis=0) != 0, np.diff(ref["x"], axis=0) != 0
    assert np.array_equal(moved_gpu, moved_ref) and moved_gpu.sum() > 10000
    assert np.all(got["x"] == np.round(got["x"])) and got["x"].min() >= 0
    info = s.info()["steppers"][0]["x"]
    refs = [orc.OracleSampler("poisson_dens", None, {"x": {"type": "int", "lower": 0}}, seed=2, chain=c) for c in range(4)]
    for q in refs:
        q.burn(500)
    assert np.array_equal(info["prop_log_scale"][:4], np.array([q.info()[0, 0] for q in refs]))
    assert np.array_equal(info["acceptance_count"][:4], np.array([q.info()[0, 2] for q in refs], dtype=np.int32))
    assert info["batch_count"] == 10


def test_multidim_real_and_int_steppers(gpu_pkg, orc):
    """tests/test_mcmc_js.R:83-121: 2x2 real and int parameters, per-component option arrays, random top-level order."""
    pls = [[np.log(50), np.log(5)], [np.log(0.5), np.log(0.05)]]
    params = {"x": {"type": "real", "dim": [2, 2], "init": [[1000, 10], [0.1, 0.001]]}}
    s, got, ref = _run(gpu_pkg, orc, models.multivar_norm_dens(gpu_pkg.ld), "multivar_norm_dens", params, None, None, 128, 3, 200, 150,
                       options={"prop_log_scale": pls}, comp_options={"x": {"prop_log_scale": np.array(pls).reshape(-1)}})
    assert got["x"].shape == (150, 128, 2, 2) and _same(got["x"], ref["x"])
    params = {"x": {"type": "int", "dim": [2, 2], "lower": 0, "init": [[1, 10], [1000, 100000]]}}
    tar = [[0.2, 0.3], [0.4, 0.5]]
    s, got, ref = _run(gpu_pkg, orc, models.multivar_poisson_dens(gpu_pkg.ld), "multivar_poisson_dens", params, None, None, 128, 4, 95, 120,
                       options={"target_accept_rate": tar, "batch_size": 10},
                       comp_options={"x": {"target_accept_rate": np.array(tar).reshape(-1), "batch_size": 10.0}})
    assert _same(got["x"], ref["x"])
    assert s.info()["steppers"][0]["x"]["batch_count"].tolist() == [[21, 21], [21, 21]]


def test_binary_steppers(gpu_pkg, orc):
    """tests/test_mcmc_js.R:123-142: P(x=1) = 0.85; 2x2 binary with P(x1=1) = 1/1.3, P(x4=1) = 1/1.5."""
    s, got, ref = _run(gpu_pkg, orc, models.bern_dens(gpu_pkg.ld), "bern_dens", {"x": {"type": "binary"}}, None, None, 2048, 5, 2, 60)
    assert _same(got["x"], ref["x"]) and abs(got["x"][1:].mean() - 0.85) < 0.01
    s, got, ref = _run(gpu_pkg, orc, models.multi_bern_dens(gpu_pkg.mcmc), "multi_bern_dens", {"x": {"type": "binary", "dim": [2, 2]}}, None, None, 2048, 6, 5, 40)
    assert _same(got["x"], ref["x"])
    assert abs(got["x"][:, :, 0, 0].mean() - 1 / 1.3) < 0.02 and abs(got["x"][:, :, 1, 1].mean() - 1 / 1.5) < 0.02


def test_complex_model_mixes_all_stepper_kinds(gpu_pkg, orc):
    """tests/test_mcmc_js.R:205-222, 238-254: real p1 in [0,1] + int n1 >= 1 + binary m, negative-binomial likelihood, global and
    per-parameter option override, thin(10)."""
    x = [float(v) for v in np.random.default_rng(7).negative_binomial(21, 0.5, 12)]
    opts = {"max_adaptation": 0.5, "params": {"p1": {"max_adaptation": 0.1}}}
    s, got, ref = _run(gpu_pkg, orc, models.complex_model_post(gpu_pkg.ld, gpu_pkg.mcmc), "complex", models.PARAMS_COMPLEX, x, {"x": np.array(x)},
                       256, 8, 120, 300, options=opts, thin=10,
                       comp_options={"p1": {"max_adaptation": 0.1}, "n1": {"max_adaptation": 0.5}})
    assert got["p1"].shape == (30, 256)                                   # sample(300) with thin 10 (R:245 analogue)
    for k in ("p1", "n1", "m"):
        assert _same(got[k], ref[k]), k
    assert set(np.unique(got["m"])) <= {0.0, 1.0} and np.all(got["n1"] >= 1)


def test_hierarchical_binomial(gpu_pkg, orc):
    """tests/test_mcmc_js.R:256-267: p dim [1,6] in [0,1], interleaved ld.norm(logit) / ld.binom terms, thin."""
    d = models.BINOM_DATA
    s, got, ref = _run(gpu_pkg, orc, models.hierarchical_binomial_post(gpu_pkg.ld, gpu_pkg.mcmc), "hier_binom", models.PARAMS_HIER_BINOM, d,
                       {"x": np.array(d["x"], float), "n": np.array(d["n"], float)}, 128, 9, 200, 400, thin=100)
    assert got["p"].shape == (4, 128, 1, 6)
    for k in ("p", "mu_logit_p", "sigma_logit_p"):
        assert _same(got[k], ref[k]), k
