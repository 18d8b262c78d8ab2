"""Shared helpers for the golden-vector tests (tests/golden/reference_js.json, made by oracle/minijs/make_golden.py by executing
the unmodified reference JS)."""
import json
import os

import numpy as np

import models
from conftest import NORM_DATA

HERE = os.path.dirname(os.path.abspath(__file__))
INF = float("inf")


def load():
    with open(os.path.join(HERE, "golden", "reference_js.json")) as f:
        return json.load(f)


def unhex(v):
    """hex-string doubles -> floats, recursively."""
    if isinstance(v, str):
        if v.startswith(("0x", "-0x")) or v in ("inf", "-inf", "nan"):
            return float.fromhex(v)
        return v
    if isinstance(v, list):
        return [unhex(x) for x in v]
    if isinstance(v, dict):
        return {k: unhex(x) for k, x in v.items()}
    return v


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


# golden case -> (oracle C model, Python log_post factory, params, data for the product, data for the oracle)
PARAMS_BY_NAME = {"params1": models.PARAMS1, "params_complex_model": models.PARAMS_COMPLEX, "params_hierarchical_binomial": models.PARAMS_HIER_BINOM}
DATA_BY_NAME = {"norm_data": NORM_DATA, "binom_data": models.BINOM_DATA}


def resolve_case(case, pkg):
    ld, mcmc = pkg.ld, pkg.mcmc
    lp = case["log_post"]
    table = {
        "readme_norm_post": ("norm_readme", models.norm_post_readme(ld)),
        "norm_post": ("norm_test", models.norm_post_test(ld)),
        "readme_beta_bern": ("beta_bern", models.beta_bern(ld)),
        "spike_bern": ("spike_bern", models.spike_bern(ld, mcmc)),
        "norm_dens": ("norm_dens", models.norm_dens(ld)),
        "poisson_dens": ("poisson_dens", models.poisson_dens(ld)),
        "multivar_norm_dens": ("multivar_norm_dens", models.multivar_norm_dens(ld)),
        "multivar_poisson_dens": ("multivar_poisson_dens", models.multivar_poisson_dens(ld)),
        "bern_dens": ("bern_dens", models.bern_dens(ld)),
        "multi_bern_dens": ("multi_bern_dens", models.multi_bern_dens(mcmc)),
        "complex_model_post": ("complex", models.complex_model_post(ld, mcmc)),
        "hierarchical_binomial_post": ("hier_binom", models.hierarchical_binomial_post(ld, mcmc)),
        "hier_norm_post": ("hier_norm", models.hier_norm_post(ld)),
        "pois_reg_post": ("pois_reg", models.pois_reg_post(ld, mcmc)),
    }
    c_model, py_model = table[lp]
    params = PARAMS_BY_NAME[case["params"]] if isinstance(case["params"], str) else case["params"]
    data = DATA_BY_NAME[case["data"]] if isinstance(case["data"], str) else case["data"]
    if isinstance(data, dict):
        data_c = {k: np.asarray(v, dtype=np.float64) for k, v in data.items()}
    elif data is None:
        data_c = None
    else:
        data_c = {"x": np.asarray(data, dtype=np.float64)} if c_model == "complex" else np.asarray(data, dtype=np.float64)
    return c_model, py_model, params, data, data_c


def flat_info(info_obj):
    """golden final_info[param] (an info object, or nested arrays of them) -> list of per-component dicts, row-major."""
    if isinstance(info_obj, list):
        out = []
        for x in info_obj:
            out.extend(flat_info(x))
        return out
    return [unhex(info_obj)] if info_obj else []


# ---- stand-alone stepper scenarios (G["steppers"]) --------------------------------------------------------------------
NB12 = [float(v) for v in np.random.default_rng(7).negative_binomial(21, 0.5, 12)]


def stepper_setup(case, pkg):
    """-> dict(c_model, params, state, options, posterior(state) -> zero-arg closure, data_c) mirroring the JS text in the case."""
    ld, mcmc = pkg.ld, pkg.mcmc
    cls = case["class"]
    inf = INF
    table = {
        "RealMetropolisStepper": dict(c_model="norm_dens", params={"x": {"lower": -inf, "upper": inf, "dim": [1]}}, state={"x": 0.0}, options=None,
                                      model=models.norm_dens(ld), type="real"),
        "IntMetropolisStepper": dict(c_model="poisson_dens", params={"x": {"lower": 0, "upper": inf, "dim": [1]}}, state={"x": 1.0}, options=None,
                                     model=models.poisson_dens(ld), type="int"),
        "MultiRealComponentMetropolisStepper": dict(c_model="multivar_norm_dens", params={"x": {"lower": -inf, "upper": inf, "dim": [2, 2]}},
                                                    state={"x": [[0.0, 0.0], [0.0, 0.0]]}, options={"max_adaptation": 0.2, "prop_log_scale": [[10, 0], [-10, 5]]},
                                                    model=models.multivar_norm_dens(ld), type="real"),
        "MultiIntComponentMetropolisStepper": dict(c_model="multivar_poisson_dens", params={"x": {"lower": 0, "upper": inf, "dim": [2, 2]}},
                                                   state={"x": [[0.0, 0.0], [0.0, 0.0]]},
                                                   options={"batch_size": 10, "target_accept_rate": [[0.22, 0.22], [0.75, 0.10]], "prop_log_scale": [[1, 10], [30, 1]]},
                                                   model=models.multivar_poisson_dens(ld), type="int"),
        "BinaryStepper": dict(c_model="bern_dens", params={"x": {"type": "binary"}}, state={"x": 0.0}, options=None, model=models.bern_dens(ld), type="binary"),
        "BinaryComponentStepper": dict(c_model="multi_bern_dens", params={"x": {"type": "binary", "dim": [2, 2]}}, state={"x": [[0.0, 0.0], [0.0, 0.0]]},
                                       options=None, model=models.multi_bern_dens(mcmc), type="binary"),
        "AmwgStepper": dict(c_model="complex", params=mcmc.complete_params(models.PARAMS_COMPLEX), state={"p1": 0.5, "n1": 1.0, "m": 1.0},
                            options={"max_adaptation": 0.5, "params": {"p1": {"max_adaptation": 0.1}}}, model=None, type=None),
    }
    su = dict(table[cls])
    su["data_c"] = {"x": np.asarray(NB12)} if cls == "AmwgStepper" else None
    return su
