"""The JavaScript host driving the GPU: js/mcmc.js + js/amwg_trace.js run under the repo's ES5 interpreter with a native binding
that makes the calls of js/amwg_napi.cc (tests/js_host.py). The models are JAVASCRIPT SOURCE TEXT -- the reference's README and
test fixtures -- recorded from `log_post.toString()`; every golden scenario of tests/golden/reference_js.json (made by executing
the unmodified reference JS) must be reproduced draw for draw through this path."""
import copy

import numpy as np
import pytest

import golden_util as gu
from js_host import DeviceNative, JsHost, to_py
from oracle.minijs.minijs import to_js

pytestmark = pytest.mark.gpu
G = gu.load()

# the model closures as JavaScript text (README.md:26-36, :149-164; tests/test_data.js:80-211 restated; the config-shaped models of
# oracle/minijs/make_golden.py)
JS_MODELS = r"""
var readme_norm_post = function(state, data) {
  var log_post = 0;
  log_post += ld.norm(state.mu, 0, 100);
  log_post += ld.unif(state.sigma, 0, 100);
  for(var i = 0; i < data.length; i++) {
    log_post += ld.norm(data[i], state.mu, state.sigma);
  }
  return log_post;
};
var norm_post = function(par, data) {
  var mu = par.mu;
  var sigma = par.sigma;
  var log_post = 0;
  log_post += ld.norm(mu, 0, 100);
  log_post += ld.unif(sigma, 0, 100);
  for(var i = 0; i < data.length; i++) {
    log_post += ld.norm(data[i], mu, sigma);
  }
  par.var = sigma * sigma;
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
var norm_dens = function(par) { return ld.norm(par.x, 10, 5); };
var poisson_dens = function(par) { return ld.pois(par.x, 10); };
var bern_dens = function(par) { return ld.bern(par.x, 0.85); };
var multivar_norm_dens = function(par) {
  x1 = par.x[0][0]; x2 = par.x[0][1]; x3 = par.x[1][0]; x4 = par.x[1][1];          // sloppy-mode implicit globals, as in the fixture
  return ld.norm(x1, 1000, 50) + ld.norm(x2, 10, 5) + ld.norm(x3, 0.1, 0.5) + ld.norm(x4, 0.001, 0.05);
};
var multivar_poisson_dens = function(par) {
  x1 = par.x[0][0]; x2 = par.x[0][1]; x3 = par.x[1][0]; x4 = par.x[1][1];
  return ld.pois(x1, 0.1) + ld.pois(x2, 10) + ld.pois(x3, 1000) + ld.pois(x4, 100000);
};
var multi_bern_dens = function(par) {
  x1 = par.x[0][0]; x2 = par.x[0][1]; x3 = par.x[1][0]; x4 = par.x[1][1];
  return Math.log(x1 * x2 * 0.85 + (1 - x1 * x2) * 0.15) + Math.log(x3 * x4 * 0.75 + (1 - x3 * x4) * 0.25);
};
var complex_model_post = function(par, x) {
  var p1 = par.p1, n1 = par.n1, m = par.m;
  var log_post = 0;
  log_post += ld.bern(m, 0.4);
  log_post += ld.beta(p1, 2, 2);
  log_post += ld.nbinom(n1, 2, 0.1);
  for(var i = 0; i < x.length; i++) {
    if(m === 0) {
      log_post += ld.nbinom(x[i], 21, 0.5);
    } else {
      log_post += ld.nbinom(x[i], n1, p1);
    }
  }
  return log_post;
};
var logit = function(p) { return Math.log(p / (1 - p)); };
var hierarchical_binomial_post = function(par, d) {
  var p = par.p[0];
  var mu_logit_p = par.mu_logit_p, sigma_logit_p = par.sigma_logit_p;
  var log_post = 0;
  log_post += ld.norm(mu_logit_p, 0, 10);
  log_post += ld.norm(sigma_logit_p, 0, 10);
  for(var i = 0; i < d.x.length; i++) {
    log_post += ld.norm(logit(p[i]), mu_logit_p, sigma_logit_p);
    log_post += ld.binom(d.x[i], d.n[i], p[i]);
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


@pytest.fixture(scope="module")
def host(gpu_pkg):
    h = JsHost(native=DeviceNative(gpu_pkg))
    h.it.set_global("mcmc", h.load("mcmc"))
    h.it.set_global("ld", h.load("distributions"))
    h.run(JS_MODELS)
    return h


def _nested(v):
    return np.asarray(v, dtype=np.float64)


@pytest.mark.parametrize("case", G["samplers"], ids=lambda c: f"{c['name']}-chain{c['chain']}")
def test_js_models_reproduce_the_reference_draw_for_draw(case, host, gpu_pkg):
    it = host.it
    _c, _py, params, data, _dc = gu.resolve_case(case, gpu_pkg)
    opts = copy.deepcopy(case["options"]) or {}
    opts.update({"seed": case["seed"], "first_chain": case["chain"], "chains": 1, "faithful": True})
    it.set_global("the_params", to_js(it, gu.unhex(copy.deepcopy(params))))
    it.set_global("the_data", to_js(it, data))
    it.set_global("the_options", to_js(it, opts))
    host.run("var S = new mcmc.AmwgSampler(the_params, %s, the_data, the_options);" % case["log_post"])
    results = iter(case["results"])
    for step in case["script"]:
        op = step[0]
        if op == "burn":
            host.run("S.burn(%d);" % step[1])
        elif op == "thin":
            host.run("S.thin(%d);" % step[1])
        elif op == "monitor":
            it.set_global("the_monitor", to_js(it, step[1]))
            host.run("S.monitor(the_monitor);")
        elif op == "stop_adaptation":
            host.run("S.stop_adaptation();")
        elif op == "start_adaptation":
            host.run("S.start_adaptation();")
        elif op == "sample":
            want = gu.unhex(next(results)["draws"])
            host.run("var draws = S.sample(%d);" % step[1])
            got = to_py(host.get("draws"))
            assert list(got.keys()) == list(want.keys())
            for k in want:
                assert gu.same(_nested(got[k]), _nested(want[k])), (case["name"], k)
    host.run("var final_state = S.state(); var final_info = S.info().steppers[0]; S.close();")
    st = to_py(host.get("final_state"))
    for name, want in gu.unhex(case["final_state"]).items():
        assert gu.same(_nested(st[name]).reshape(-1), _nested(want).reshape(-1)), name
    info = to_py(host.get("final_info"))
    for name in params:
        want = gu.flat_info(case["final_info"].get(name, {}))
        if not want:
            assert info[name] == {}
            continue
        for key in ("prop_log_scale", "acceptance_count", "iterations_since_adaption", "batch_count"):
            assert gu.same(_nested(info[name][key]).reshape(-1), [w[key] for w in want]), (name, key)


def test_js_ld_and_helpers_on_the_device(host):
    """ld.* called with plain numbers from JavaScript and the exported RNG helpers: the golden values of the reference."""
    it = host.it
    for fname, rows in G["ld"].items():
        for r in rows[:6]:
            args = gu.unhex(r[0])
            it.set_global("the_args", to_js(it, args))
            host.run("var ld_val = ld.%s.apply(ld, the_args);" % fname)
            got, want = float(host.get("ld_val")), float.fromhex(r[1])
            tol = 8 * np.spacing(abs(want)) if fname in ("t", "weibull", "bivarnorm", "dirichlet", "cat") else 0.0
            assert got == want or (np.isnan(got) and np.isnan(want)) or abs(got - want) <= tol, (fname, args, got, want)
    h = G["helpers"]
    host.run("mcmc.set_random_stream(%d, %d, 0); var r10 = []; for (var i = 0; i < 25; i++) { r10.push(mcmc.rnorm(10, 5)); }" % (h["seed"], h["chain"]))
    assert gu.same(to_py(host.get("r10")), [float.fromhex(v) for v in h["rnorm_10_5"][:25]])


def test_js_many_chains_and_specialised_kernel(host, monkeypatch):
    """chains > 1 from JavaScript: an extra chain axis, and a model big enough to run the run-time specialised sweep."""
    it = host.it
    x = np.random.default_rng(12).normal(50.0, 3.0, 128)
    it.set_global("the_data", to_js(it, [float(v) for v in x]))
    monkeypatch.setenv("AMWG_JIT", "1")
    host.run("""
      var S2 = new mcmc.AmwgSampler({mu: {type: "real"}, sigma: {type: "real", lower: 0}}, readme_norm_post, the_data, {chains: 512, seed: 4});
      var kern = S2.sweep_kernel();
      S2.burn(600);
      var d2 = S2.sample(20);
      S2.close();
    """)
    assert to_py(host.get("kern")).startswith("specialised")
    d = to_py(host.get("d2"))
    mu = np.asarray(d["mu"])
    assert mu.shape == (20, 512)
    assert abs(mu.mean() - x.mean()) < 0.05 and abs(np.asarray(d["sigma"]).mean() - x.std(ddof=1)) < 0.08


@pytest.mark.parametrize("case", G["steppers"], ids=lambda c: f"{c['class']}-chain{c['chain']}")
def test_js_standalone_steppers_match_the_reference(case, host):
    """mcmc.RealMetropolisStepper & co from JavaScript, constructed and stepped with the very JS text the golden generator ran
    against the unmodified reference (tests/test_mcmc_js.R:55-142 use the steppers this way)."""
    it = host.it
    it.set_global("X", to_js(it, [float(v) for v in gu.NB12]))
    host.run("""
      var params_complex_model = {p1: {type: "real", lower: 0, upper: 1}, n1: {type: "int", lower: 1, init: 1}, m: {type: "binary"}};
      var state = %s;
      var stepper_options = %s || {};
      stepper_options.seed = %d; stepper_options.first_chain = %d; stepper_options.faithful = true;
      var stepper = new mcmc.%s(%s, state, function () { return %s; }, stepper_options);
    """ % (case["state_js"], case["options_js"] or "null", case["seed"], case["chain"], case["class"], case["params_js"], case["posterior_js"]))
    results = iter(case["results"])
    for step in case["script"]:
        if step[0] == "step":
            host.run("var outs = []; for (var i = 0; i < %d; i++) { outs.push(JSON_clone(stepper.step())); }" % step[1])
            got, want = to_py(host.get("outs")), gu.unhex(next(results))
            if case["class"] == "AmwgStepper":
                for k in want[0]:
                    assert gu.same([g[k] for g in got], [w[k] for w in want]), k
            else:
                assert gu.same(_nested(got), _nested(want))
        elif step[0] == "stop_adaptation":
            host.run("stepper.stop_adaptation();")
        elif step[0] == "start_adaptation":
            host.run("stepper.start_adaptation();")
    st = to_py(host.get("state"))
    for name, want in gu.unhex(case["final_state"]).items():
        assert gu.same(_nested(st[name]).reshape(-1), _nested(want).reshape(-1)), name
