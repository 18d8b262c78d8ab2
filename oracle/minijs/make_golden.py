#!/usr/bin/env python
"""Generate tests/golden/reference_js.json by EXECUTING the unmodified reference sources
(/root/reference/mcmc.js, distributions.js, tests/test_data.js) in oracle/minijs, with

    Math.random := the Philox stream (seed, chain) of DESIGN.md "RNG contract"   (oracle/liboracle.so: orc_stream_uniform)
    Math.log / Math.exp := the fdlibm algorithms V8 ports                         (oracle/liboracle.so: orc_log / orc_exp)

Run here (the build container has /root/reference; the GPU box does not):   python oracle/minijs/make_golden.py
Doubles are stored as C99 hex strings so the comparison in tests/test_golden.py is bit for bit.
"""
from __future__ import annotations

import json
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import oracle.oracle as orc  # noqa: E402
from oracle.minijs.minijs import Interpreter, JSArray, JSFunction, JSObject, JSThrow, to_js, undefined  # noqa: E402

REF = "/root/reference"


class Stream:
    def __init__(self, L, seed, chain):
        self.L, self.seed, self.chain, self.n = L, seed, chain, 0

    def __call__(self):
        u = self.L.orc_stream_uniform(self.seed, self.chain, self.n)
        self.n += 1
        return u


def hexify(v):
    """JS value -> JSON-able, doubles as hex strings."""
    if isinstance(v, JSArray): return [hexify(x) for x in v.items]
    if isinstance(v, JSFunction): return "<function>"
    if isinstance(v, JSObject): return {k: hexify(x) for k, x in v.props.items()}
    if v is undefined: return "<undefined>"
    if isinstance(v, bool) or v is None or isinstance(v, str): return v
    if isinstance(v, float): return float(v).hex()
    return v


def new_engine(seed=0, chain=0):
    L = orc.lib()
    st = Stream(L, seed, chain)
    it = Interpreter(math_log=L.orc_log, math_exp=L.orc_exp, random=st)
    mod = it.new_object()
    mod.put("exports", it.new_object())
    it.set_global("module", mod)
    it.run(open(os.path.join(REF, "mcmc.js")).read())                   # UMD wrapper takes the CommonJS branch (mcmc.js:13-17)
    it.set_global("mcmc", mod.get("exports"))
    mod.put("exports", it.new_object())
    it.run(open(os.path.join(REF, "distributions.js")).read())
    it.set_global("ld", mod.get("exports"))
    it.run(open(os.path.join(REF, "tests", "test_data.js")).read())    # fixtures become globals, as in the R driver (test_mcmc_js.R:33-36)

    def clone(this, a):                                                 # snapshot of a live object (AmwgStepper.step returns `state` itself)
        def cp(v):
            if isinstance(v, JSArray): return it.new_array([cp(x) for x in v.items])
            if isinstance(v, JSObject) and not isinstance(v, JSFunction):
                o = it.new_object()
                for k2, x in v.props.items(): o.put(k2, cp(x))
                return o
            return v
        return cp(a[0])
    it.set_global("JSON_clone", it.make_native("JSON_clone", clone))
    return it, st


# the models of README.md, as JavaScript text (README.md:26-36 and :149-164), and the config-3 spike model
README_JS = """
var readme_norm_post = function(state, data) {
  var log_post = 0;
  log_post += ld.norm(state.mu, 0, 100);
  log_post += ld.unif(state.sigma, 0, 100);
  for(var i = 0; i < data.length; i++) {
    log_post += ld.norm(data[i], state.mu, state.sigma);
  }
  return log_post;
};
var readme_beta_bern = function(state, data) {
  var log_post = 0;
  log_post += ld.beta(state.theta, 2, 2);
  var n = data.x.length;
  for(var i = 0; i < n; i++) {
    log_post += ld.bern(data.x[i], state.theta)
  }
  return log_post;
};
var spike_bern = function(state, data) {
  var theta = state.theta, m = state.m;
  var log_post = 0;
  log_post += ld.beta(theta, 2, 2);
  log_post += ld.bern(m, 0.5);
  for(var i = 0; i < data.x.length; i++) {
    if(m === 0) { log_post += ld.bern(data.x[i], 0.5); } else { log_post += ld.bern(data.x[i], theta); }
  }
  return log_post;
};
var hier_norm_post = function(state, data) {
  var log_post = 0;
  for(var j = 0; j < state.mu.length; j++) { log_post += ld.norm(state.mu[j], 0, 100); }
  log_post += ld.unif(state.sigma, 0, 100);
  for(var i = 0; i < data.y.length; i++) { log_post += ld.norm(data.y[i], state.mu[data.g[i]], state.sigma); }
  return log_post;
};
var pois_reg_post = function(state, data) {
  var log_post = 0;
  for(var k = 0; k < state.beta.length; k++) { log_post += ld.norm(state.beta[k], 0, 10); }
  for(var i = 0; i < data.y.length; i++) {
    var eta = 0;
    for(var k = 0; k < state.beta.length; k++) { eta += data.X[i][k] * state.beta[k]; }
    log_post += ld.pois(data.y[i], Math.exp(eta));
  }
  return log_post;
};
"""

PRESIDENTS = [183, 192, 182, 183, 177, 185, 188, 188, 182, 185]
Y8 = [1, 0, 1, 1, 0, 1, 1, 1]
Y40 = [int(v) for v in (np.random.default_rng(40).random(40) < 0.7)]
NB12 = [int(v) for v in np.random.default_rng(7).negative_binomial(21, 0.5, 12)]
# BASELINE configs 4 and 5 in miniature: hierarchical Normal (3 groups x 4) and Poisson regression (N = 12, K = 2)
_rg = np.random.default_rng(45)
HIER_G = [j for j in range(3) for _ in range(4)]
HIER_Y = [float(np.round(v, 3)) for v in (np.array([95.0, 100.0, 108.0])[HIER_G] + _rg.normal(0, 5, 12))]
POIS_X = [[1.0, float(np.round(v, 3))] for v in _rg.normal(0, 0.5, 12)]
POIS_Y = [int(v) for v in _rg.poisson(np.exp(np.array(POIS_X) @ np.array([0.8, -0.6])))]

# (name, log_post JS expression, params (Python), data (Python), options (Python), script of calls)
SAMPLER_CASES = [
    ("readme_normal", "readme_norm_post", {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}, PRESIDENTS, None,
     [("burn", 100), ("sample", 150)]),
    ("test_norm_post_derived_thin", "norm_post", "params1", "norm_data", {"thin": 3},
     [("burn", 40), ("sample", 100), ("monitor", ["sigma", "var"]), ("sample", 10), ("monitor", None), ("thin", 1), ("sample", 5)]),
    ("readme_beta_bernoulli", "readme_beta_bern", {"theta": {"type": "real", "lower": 0, "upper": 1}}, {"x": Y8}, {"thin": 2},
     [("sample", 200)]),
    ("spike_bernoulli_binary", "spike_bern", {"theta": {"type": "real", "lower": 0, "upper": 1}, "m": {"type": "binary"}}, {"x": Y40}, None,
     [("burn", 60), ("sample", 120)]),
    ("real_stepper_norm_dens", "norm_dens", {"x": {"type": "real"}}, None, None, [("burn", 120), ("sample", 150)]),
    ("int_stepper_poisson", "poisson_dens", {"x": {"type": "int", "lower": 0}}, None, None, [("burn", 120), ("sample", 250)]),
    ("multidim_real", "multivar_norm_dens", {"x": {"type": "real", "dim": [2, 2], "init": [[1000, 10], [0.1, 0.001]]}}, None,
     {"prop_log_scale": [[math.log(50), math.log(5)], [math.log(0.5), math.log(0.05)]]}, [("burn", 110), ("sample", 100)]),
    ("multidim_int", "multivar_poisson_dens", {"x": {"type": "int", "dim": [2, 2], "lower": 0, "init": [[1, 10], [1000, 100000]]}}, None,
     {"target_accept_rate": [[0.2, 0.3], [0.4, 0.5]], "batch_size": 10}, [("burn", 95), ("sample", 100)]),
    ("binary_stepper", "bern_dens", {"x": {"type": "binary"}}, None, None, [("sample", 200)]),
    ("binary_component_stepper", "multi_bern_dens", {"x": {"type": "binary", "dim": [2, 2]}}, None, None, [("burn", 5), ("sample", 150)]),
    ("complex_model_options_thin", "complex_model_post", "params_complex_model", NB12,
     {"max_adaptation": 0.5, "params": {"p1": {"max_adaptation": 0.1}}}, [("thin", 10), ("burn", 100), ("sample", 300)]),
    ("hierarchical_binomial", "hierarchical_binomial_post", "params_hierarchical_binomial", "binom_data", None,
     [("thin", 100), ("burn", 100), ("sample", 600)]),
    ("adaptation_toggle", "readme_beta_bern", {"theta": {"type": "real", "lower": 0, "upper": 1}}, {"x": Y40}, None,
     [("burn", 175), ("stop_adaptation",), ("burn", 60), ("start_adaptation",), ("burn", 40), ("sample", 10)]),
    ("options_or_quirk", "readme_norm_post", {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}, PRESIDENTS,
     {"is_adapting": True, "prop_log_scale": 2, "batch_size": 7, "params": {"mu": {"is_adapting": False, "prop_log_scale": 0, "batch_size": 0}}},
     [("burn", 30), ("sample", 20)]),
    ("config4_shape_hierarchical_normal", "hier_norm_post", {"mu": {"type": "real", "dim": [3]}, "sigma": {"type": "real", "lower": 0}},
     {"y": HIER_Y, "g": HIER_G}, None, [("burn", 120), ("sample", 80)]),
    ("config5_shape_poisson_regression", "pois_reg_post", {"beta": {"type": "real", "dim": [2]}}, {"y": POIS_Y, "X": POIS_X}, None,
     [("burn", 120), ("sample", 80)]),
]


# stand-alone steppers, as tests/test_mcmc_js.R:55-142 drives them (+ an AmwgStepper over the "complex" model, R:205-222)
STEPPER_CASES = [
    ("RealMetropolisStepper", "{x: {lower: -Infinity, upper: Infinity, dim:[1]}}", "{x: 0}", "norm_dens(state)", None,
     [("step", 200)]),
    ("IntMetropolisStepper", "{x: {lower: 0, upper: Infinity, dim:[1]}}", "{x: 1}", "poisson_dens(state)", None,
     [("step", 300)]),
    ("MultiRealComponentMetropolisStepper", "{x: {lower: -Infinity, upper: Infinity, dim: [2, 2]}}", "{x: [[0, 0], [0, 0]]}", "multivar_norm_dens(state)",
     "{max_adaptation: 0.2, prop_log_scale: [[10,0],[-10, 5]]}", [("step", 100), ("stop_adaptation",), ("step", 100), ("start_adaptation",), ("step", 200)]),
    ("MultiIntComponentMetropolisStepper", "{x: {lower: 0, upper: Infinity, dim: [2, 2]}}", "{x: [[0, 0], [0, 0]]}", "multivar_poisson_dens(state)",
     "{batch_size: 10, target_accept_rate: [[0.22, 0.22],[0.75, 0.10]], prop_log_scale: [[1,10],[30, 1]]}",
     [("step", 100), ("stop_adaptation",), ("step", 100), ("start_adaptation",), ("step", 200)]),
    ("BinaryStepper", "{x: {type: 'binary'}}", "{x: 0}", "bern_dens(state)", None, [("step", 200)]),
    ("BinaryComponentStepper", "{x: {type: 'binary', dim: [2,2]}}", "{x: [[0, 0], [0, 0]]}", "multi_bern_dens(state)", None, [("step", 150)]),
    ("AmwgStepper", "mcmc.complete_params(params_complex_model, mcmc.param_init_fixed)", "{p1: 0.5, n1: 1, m: 1}", "complex_model_post(state, X)",
     "{max_adaptation: 0.5, params: {p1: {max_adaptation: 0.1}}}", [("step", 150)]),
]


def run_stepper_case(case, seed, chain):
    cls, params_js, state_js, post_js, options_js, script = case
    it, st = new_engine(seed, chain)
    it.set_global("X", to_js(it, NB12))
    it.run(f"var parameters = {params_js}; var state = {state_js}; var posterior = function() {{ return {post_js}; }};"
           f"var options = {options_js if options_js else 'undefined'};"
           f"var stepper = new mcmc.{cls}(parameters, state, posterior, options);")
    stepper = it.get_global("stepper")
    out = {"class": cls, "seed": seed, "chain": chain, "params_js": params_js, "state_js": state_js, "posterior_js": post_js,
           "options_js": options_js, "script": [list(x) for x in script], "results": []}
    for step in script:
        if step[0] == "step":
            it.run(f"var out = []; for (var k = 0; k < {step[1]}; k++) {{ var r = stepper.step(); out.push(r === state ? JSON_clone(r) : r); }}")
            out["results"].append(hexify(it.get_global("out")))
        elif step[0] == "stop_adaptation":
            it.run("stepper.stop_adaptation();")
        else:
            it.run("stepper.start_adaptation();")
    if cls == "AmwgStepper":                         # info() keyed through each substepper's own name (shuffle quirk, see stepper_info)
        info = {}
        for sub in stepper.get("substeppers").items:
            info[sub.get("param_name")] = hexify(it.get_prop(sub, "info").call(sub, []))
        out["final_info"] = info
    else:
        out["final_info"] = hexify(it.get_prop(stepper, "info").call(stepper, []))
    out["final_state"] = hexify(it.get_global("state"))
    out["uniforms_consumed"] = st.n
    return out


def stepper_info(it, sampler):
    """AmwgStepper.info() of the (single) stepper, with names re-attributed through the substepper's own param_name
    (the reference pairs param_names[i] with the shuffled substeppers[i])."""
    st = sampler.get("steppers").items[0]
    out = {}
    for sub in st.get("substeppers").items:
        name = sub.get("param_name")
        info = it.get_prop(sub, "info").call(sub, [])
        out[name] = hexify(info)
    return out


def run_sampler_case(case, seed, chain):
    name, lp, params, data, options, script = case
    it, st = new_engine(seed, chain)
    it.run(README_JS)
    it.set_global("P", it.get_global(params) if isinstance(params, str) else to_js(it, params))
    it.set_global("D", it.get_global(data) if isinstance(data, str) else (to_js(it, data) if data is not None else undefined))
    it.set_global("O", to_js(it, options) if options is not None else undefined)
    it.run(f"var sampler = new mcmc.AmwgSampler(P, {lp}, D, O);")
    sampler = it.get_global("sampler")
    out = {"name": name, "seed": seed, "chain": chain, "log_post": lp,
           "params": params, "data": data, "options": options, "script": [list(s) for s in script], "results": []}
    out["completed_params"] = hexify(sampler.get("params"))
    for step in script:
        op = step[0]
        if op == "burn":
            it.run(f"sampler.burn({step[1]});")
        elif op == "sample":
            it.run(f"var smp = sampler.sample({step[1]});")
            out["results"].append({"sample": step[1], "draws": hexify(it.get_global("smp"))})
        elif op == "thin":
            it.run(f"sampler.thin({step[1]});")
        elif op == "monitor":
            it.set_global("M", to_js(it, step[1]) if step[1] is not None else None)
            it.run("sampler.monitor(M);")
        elif op == "stop_adaptation":
            it.run("sampler.stop_adaptation();")
        elif op == "start_adaptation":
            it.run("sampler.start_adaptation();")
    out["final_state"] = hexify(sampler.get("state"))
    out["final_info"] = stepper_info(it, sampler)
    out["uniforms_consumed"] = st.n
    return out


def main():
    L = orc.lib()
    G = {"_about": "generated by oracle/minijs/make_golden.py from the unmodified reference sources; doubles are hex strings"}
    it, st = new_engine(9, 3)
    # complete_params on the reference's own fixtures + the expected values the reference's test asserts (test_mcmc_js.R:39-46)
    it.run("var cp1 = mcmc.complete_params(params1, mcmc.param_init_fixed); var cp2 = mcmc.complete_params(params2, mcmc.param_init_fixed);")
    G["complete_params"] = {"params1": hexify(it.get_global("cp1")), "params1_expected": hexify(it.get_global("params1_completed")),
                            "params2": hexify(it.get_global("cp2")), "params2_expected": hexify(it.get_global("params2_completed"))}
    # param_init_fixed table (mcmc.js:313-341)
    tab = []
    inf = math.inf
    for t, lo, hi in [("real", -inf, inf), ("real", -inf, 3.0), ("real", 2.0, inf), ("real", 0.0, 1.0), ("int", -inf, inf), ("int", -inf, 3.0),
                      ("int", 2.0, inf), ("int", 0.0, 5.0), ("int", -3.0, -2.0), ("int", 0.0, 1.0), ("binary", 0.0, 1.0)]:
        it.set_global("a1", t); it.set_global("a2", lo); it.set_global("a3", hi)
        it.run("var r = mcmc.param_init_fixed(a1, a2, a3);")
        tab.append([t, float(lo).hex(), float(hi).hex(), float(it.get_global("r")).hex()])
    thrown = []
    for t, lo, hi in [("real", 2.0, 1.0), ("complex", -inf, inf)]:
        it.set_global("a1", t); it.set_global("a2", lo); it.set_global("a3", hi)
        try:
            it.run("mcmc.param_init_fixed(a1, a2, a3);")
        except JSThrow as e:
            thrown.append([t, float(lo).hex(), float(hi).hex(), e.value])
    G["param_init_fixed"] = {"values": tab, "throws": thrown}
    # ld.* on a grid of inputs (distributions.js has no tests of its own)
    rng = np.random.default_rng(123)
    ld_cases = {}
    arg_sets = {
        "norm": [(183, 180, 5), (0.5, 0, 100)] + [(rng.normal(0, 50), rng.normal(0, 50), rng.uniform(0.01, 100)) for _ in range(40)],
        "unif": [(1, 0, 100), (101, 0, 100), (-1, 0, 100)] + [(rng.uniform(-1, 2), 0, 1) for _ in range(10)],
        "beta": [(0.3, 2, 2), (0.3, 1, 1), (1.5, 2, 2)] + [(rng.uniform(0, 1), rng.uniform(0.5, 5), rng.uniform(0.5, 5)) for _ in range(40)],
        "bern": [(1, 0.85), (0, 0.85), (0.5, 0.3)] + [(float(rng.integers(0, 2)), rng.uniform(0, 1)) for _ in range(20)],
        "pois": [(3, 10), (-1, 3), (0, 0.1)] + [(float(rng.integers(0, 60)), rng.uniform(0.01, 40)) for _ in range(40)],
        "lgamma": [(0.5,), (10,), (100.5,)] + [(rng.uniform(0.01, 200),) for _ in range(40)],
        "lfactorial": [(-1,), (0,), (5,)] + [(float(rng.integers(0, 100)),) for _ in range(20)],
        "lchoose": [(float(rng.integers(1, 60)), float(rng.integers(0, 30))) for _ in range(20)],
        "lbeta": [(rng.uniform(0.1, 30), rng.uniform(0.1, 30)) for _ in range(20)],
        "cauchy": [(rng.normal(0, 5), rng.normal(0, 5), rng.uniform(0.1, 5)) for _ in range(20)],
        "laplace": [(rng.normal(0, 5), rng.normal(0, 5), rng.uniform(0.1, 5)) for _ in range(20)],
        "gamma": [(0, 1, 2), (-1, 2, 2)] + [(rng.uniform(0, 20), rng.uniform(0.2, 9), rng.uniform(0.2, 9)) for _ in range(20)],
        "invgamma": [(0, 1, 2)] + [(rng.uniform(0.01, 20), rng.uniform(0.2, 9), rng.uniform(0.2, 9)) for _ in range(20)],
        "lnorm": [(0, 1, 2)] + [(rng.uniform(0.01, 20), rng.normal(0, 2), rng.uniform(0.2, 3)) for _ in range(20)],
        "pareto": [(rng.uniform(0.1, 20), rng.uniform(0.2, 9), rng.uniform(0.2, 9)) for _ in range(20)],
        "logis": [(rng.normal(0, 5), rng.normal(0, 5), rng.uniform(0.1, 5)) for _ in range(20)],
        "exp": [(-1, 2)] + [(rng.uniform(0, 20), rng.uniform(0.1, 5)) for _ in range(20)],
        "binom": [(3, 10, 0), (0, 10, 0), (11, 10, .5)] + [(float(rng.integers(0, 30)), float(rng.integers(1, 30)), rng.uniform(0.01, .99)) for _ in range(30)],
        "nbinom": [(-1, 3, .5)] + [(float(rng.integers(0, 30)), float(rng.integers(1, 30)), rng.uniform(0.01, 0.99)) for _ in range(30)],
        "hyper": [(float(rng.integers(0, 10)), float(rng.integers(10, 30)), float(rng.integers(10, 30)), float(rng.integers(5, 10))) for _ in range(20)],
        "t": [(rng.normal(0, 5), rng.normal(0, 5), rng.uniform(0.1, 5), rng.uniform(1, 30)) for _ in range(20)],
        "weibull": [(rng.uniform(0, 20), rng.uniform(0.2, 5), rng.uniform(0.2, 5)) for _ in range(20)],
    }
    for fname, argl in arg_sets.items():
        rows = []
        for args in argl:
            args = [float(a) for a in args]
            it.set_global("A", to_js(it, args))
            it.run(f"var r = ld.{fname}.apply(null, A);")
            rows.append([[a.hex() for a in args], float(it.get_global("r")).hex()])
        ld_cases[fname] = rows
    for fname, pyargs in (("bivarnorm", [[1.0, 2.0], [0.5, 1.5], [1.2, 0.7], 0.3]), ("dirichlet", [[0.2, 0.3, 0.5], [1.5, 2.0, 3.0]]),
                          ("cat", [2.0, [0.2, 0.5, 0.3]])):
        it.set_global("A", to_js(it, pyargs))
        it.run(f"var r = ld.{fname}.apply(null, A);")
        ld_cases[fname] = [[json.loads(json.dumps(pyargs)), float(it.get_global("r")).hex()]]
    G["ld"] = ld_cases
    # helpers: rnorm / runif / runif_discrete / a shuffle through nested_array_random_apply is covered by the samplers
    it2, st2 = new_engine(9, 3)
    it2.run("var r = []; for (var i = 0; i < 300; i++) r.push(mcmc.rnorm(10, 5)); var u = []; for (var i = 0; i < 20; i++) u.push(mcmc.runif(2, 5));"
            "var d = []; for (var i = 0; i < 20; i++) d.push(mcmc.runif_discrete(1, 6));")
    G["helpers"] = {"seed": 9, "chain": 3, "rnorm_10_5": hexify(it2.get_global("r")), "runif_2_5": hexify(it2.get_global("u")),
                    "runif_discrete_1_6": hexify(it2.get_global("d")), "uniforms_consumed": st2.n}
    # error strings the reference throws (mcmc.js:299-300, 867)
    errs = {}
    it3, _ = new_engine(0, 0)
    for label, js in (("option_dim", 'new mcmc.AmwgSampler({x: {type: "real", dim: [2, 2]}}, multivar_norm_dens, null, {prop_log_scale: [1, 2, 3]});'),
                      ("bad_type", 'new mcmc.AmwgSampler({q: {type: "complex"}}, norm_dens);'),
                      ("bad_bounds", 'new mcmc.AmwgSampler({q: {type: "real", lower: 2, upper: 1}}, norm_dens);')):
        try:
            it3.run(js)
            errs[label] = None
        except JSThrow as e:
            errs[label] = e.value
    G["throws"] = errs
    # samplers
    cases = []
    for k, case in enumerate(SAMPLER_CASES):
        for chain in (0, 5):
            print("running", case[0], "chain", chain, flush=True)
            cases.append(run_sampler_case(case, seed=100 + k, chain=chain))
    G["samplers"] = cases
    steppers = []
    for k, case in enumerate(STEPPER_CASES):
        for chain in (0, 3):
            print("running stepper", case[0], "chain", chain, flush=True)
            steppers.append(run_stepper_case(case, seed=200 + k, chain=chain))
    G["steppers"] = steppers
    out = os.path.join(ROOT, "tests", "golden", "reference_js.json")
    with open(out, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
