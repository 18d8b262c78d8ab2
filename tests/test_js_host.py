"""The JavaScript host side (js/*.js), run without Node by the repo's ES5 interpreter (tests/js_host.py).

CPU: js/amwg_rewrite.js + js/amwg_trace.js record the reference's OWN JavaScript test models from their source text
(/root/reference is not needed: the model texts are the fixtures of oracle/minijs/make_golden.py and tests/test_data.js restated
here as strings) and must produce, word for word, the program the Python tracer produces for the same model; js/mcmc.js must hand
the native layer the same model descriptor (parameters, inits, per-component options) as the Python host."""
import numpy as np
import pytest

import models
from js_host import JsHost, RecordingNative, to_py
from oracle.minijs.minijs import JSThrow, to_js

PRESIDENTS = [183, 192, 182, 183, 177, 185, 188, 188, 182, 185]

JS_MODELS = r"""
var norm_post_readme = function(state, data) {          // README.md:26-36
  var log_post = 0;
  log_post += ld.norm(state.mu, 0, 100);
  log_post += ld.unif(state.sigma, 0, 100);
  for(var i = 0; i < data.length; i++) {
    log_post += ld.norm(data[i], state.mu, state.sigma);
  }
  return log_post;
};
var norm_post = function(par, data) {                   // tests/test_data.js:80-91
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
var beta_bern = function(state, data) {                 // README.md:149-164
  var log_post = 0;
  log_post += ld.beta(state.theta, 2, 2);
  var n = data.x.length;
  for(var i = 0; i < n; i++) {
    log_post += ld.bern(data.x[i], state.theta)
  }
  return log_post;
};
var spike_bern = function(state, data) {                // BASELINE config 3, a literal `if` on the binary parameter
  var theta = state.theta, m = state.m;
  var log_post = 0;
  log_post += ld.beta(theta, 2, 2);
  log_post += ld.bern(m, 0.5);
  for(var i = 0; i < data.x.length; i++) {
    if(m === 0) { log_post += ld.bern(data.x[i], 0.5); } else { log_post += ld.bern(data.x[i], theta); }
  }
  return log_post;
};
var complex_model_post = function(par, x) {             // tests/test_data.js:154-171
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
var logit = function(p) { return Math.log(p / (1 - p)); };      // tests/test_data.js:195-197 (a global helper)
var binomial_post = function(par, d) {                  // tests/test_data.js:199-211
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
var hier_norm_post = function(state, data) {            // BASELINE config 4
  var log_post = 0;
  for(var j = 0; j < state.mu.length; j++) { log_post += ld.norm(state.mu[j], 0, 100); }
  log_post += ld.unif(state.sigma, 0, 100);
  for(var i = 0; i < data.y.length; i++) { log_post += ld.norm(data.y[i], state.mu[data.g[i]], state.sigma); }
  return log_post;
};
"""


@pytest.fixture(scope="module")
def host():
    rec = RecordingNative()
    h = JsHost(native=rec)
    h.recorder = rec
    h.it.set_global("amwg_trace", h.load("amwg_trace"))
    h.it.set_global("mcmc", h.load("mcmc"))
    h.run(JS_MODELS)
    return h


def _py_program(pkg, log_post, params, data, **kw):
    mcmc = pkg.mcmc
    cp = mcmc.complete_params(params)
    offsets, n = {}, 0
    for name, p in cp.items():
        offsets[name] = n
        n += int(np.prod(p["dim"]))
    prog, _ = pkg.tracer.trace(log_post, cp, offsets, n, data, **kw)
    return prog


def _js_program(host, fn_name, params, data, options=None):
    it = host.it
    it.set_global("the_params", to_js(it, params))
    it.set_global("the_data", to_js(it, data))
    it.set_global("the_options", to_js(it, options or {}))
    host.run("""
      var cp = mcmc.complete_params(the_params), names = [], offsets = {}, n_comp = 0, k, j, n;
      for (k in cp) { if (cp.hasOwnProperty(k)) { names.push(k); offsets[k] = n_comp; n = 1; for (j = 0; j < cp[k].dim.length; j++) { n *= cp[k].dim[j]; } n_comp += n; } }
      var the_prog = amwg_trace.trace(%s, names, cp, offsets, n_comp, the_data, the_options);
    """ % fn_name)
    return to_py(host.get("the_prog"))


def _same_program(jp, pp):
    assert [int(x) for x in jp["code"]] == [int(x) for x in pp.code]
    assert np.array_equal(np.asarray(jp["consts"], dtype=np.float64).view(np.uint64) if jp["consts"] else np.empty(0, np.uint64),
                          np.asarray(pp.consts, dtype=np.float64).view(np.uint64)) or \
        np.array_equal(np.asarray(jp["consts"]), np.asarray(pp.consts), equal_nan=True)
    for k in ("comp_prog", "touch_off", "touch_terms", "fold_prog", "fold_dst", "block_params", "term_block_comp",
              "variant_comps", "variant_logpost", "variant_derived"):
        assert [int(x) for x in jp[k]] == [int(x) for x in getattr(pp, k)], k
    for k in ("logpost_prog", "derived_prog", "n_terms", "stat_prog"):
        assert int(jp[k]) == int(getattr(pp, k)), k
    assert jp["summary"] == pp.summary
    assert len(jp["plates"]) == len(pp.plates)
    for a, b in zip(jp["plates"], pp.plates):
        assert int(a["kind"]) == b["kind"] and int(a["n"]) == b["n"] and [int(x) for x in a["col"]] == list(b["col"]) and [int(x) for x in a["iparam"]] == list(b["iparam"])
    assert len(jp["columns"]) == len(pp.columns)
    for a, b in zip(jp["columns"], pp.columns):
        assert np.array_equal(np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1))


NB12 = [float(v) for v in np.random.default_rng(7).negative_binomial(21, 0.5, 12)]
Y40 = [float(v) for v in (np.random.default_rng(256).random(40) < 0.7)]
HIER = {"y": [float(v) for v in np.random.default_rng(5).normal(100, 5, 96)], "g": [float(v) for v in np.repeat(np.arange(6), 16)]}

CASES = [
    ("norm_post_readme", lambda pkg: models.norm_post_readme(pkg.ld), models.PARAMS_NORM, [float(v) for v in PRESIDENTS], {}),
    ("norm_post_readme", lambda pkg: models.norm_post_readme(pkg.ld), models.PARAMS_NORM, [float(v) for v in PRESIDENTS] * 8, {}),
    ("norm_post_readme", lambda pkg: models.norm_post_readme(pkg.ld), models.PARAMS_NORM, [float(v) for v in PRESIDENTS] * 8, {"faithful": True}),
    ("norm_post", lambda pkg: models.norm_post_test(pkg.ld), models.PARAMS1, [100.0, 62, 96, 122, 141, 144, 74, 73, 78, 128], {}),
    ("beta_bern", lambda pkg: models.beta_bern(pkg.ld), models.PARAMS_THETA, {"x": Y40}, {}),
    ("spike_bern", lambda pkg: models.spike_bern_literal(pkg.ld), models.PARAMS_SPIKE, {"x": Y40}, {}),
    ("complex_model_post", lambda pkg: models.complex_model_post_literal(pkg.ld), models.PARAMS_COMPLEX, NB12, {}),
    ("binomial_post", lambda pkg: models.hierarchical_binomial_post(pkg.ld, pkg.mcmc), models.PARAMS_HIER_BINOM,
     {"x": [5.0, 6, 9, 14, 13, 20], "n": [10.0, 10, 20, 20, 30, 30]}, {}),
    ("hier_norm_post", lambda pkg: models.hier_norm_post(pkg.ld), {"mu": {"type": "real", "dim": [6]}, "sigma": {"type": "real", "lower": 0}}, HIER, {}),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: c[0] + ("-faithful" if c[4].get("faithful") else "") + "-%d" % (len(c[3]) if isinstance(c[3], list) else 0))
def test_js_source_lowers_to_the_same_program_as_the_python_tracer(case, host, pkg):
    fn, py_model, params, data, opts = case
    jp = _js_program(host, fn, params, data, opts)
    pp = _py_program(pkg, py_model(pkg), params, data, **({"faithful": True} if opts.get("faithful") else {}))
    _same_program(jp, pp)


def test_rewriter_output_and_free_identifiers(host):
    rw = host.load("amwg_rewrite")
    host.it.set_global("amwg_rewrite", rw)
    host.run("var rw_out = amwg_rewrite.rewrite(binomial_post.toString());")
    out = to_py(host.get("rw_out"))
    assert out["free"] == ["ld", "logit"]                     # Math is only used inside logit; `logit` is found in the global scope
    assert "__r.call(ld, \"norm\", [logit(__r.get(p, i))" in out["source"]
    host.run("var rw2 = amwg_rewrite.rewrite('function (s) { var t = 0; for (var i = 0; i < 3; i++) { t += s.x * i; } return s.y > 1 ? t : -t; }');")
    src = to_py(host.get("rw2"))["source"]
    assert "(t = __r.add(t, __r.mul(__r.get(s, \"x\"), i)))" in src and "(i = __r.add(i, 1))" in src
    assert "(__r.t(__r.gt(__r.get(s, \"y\"), 1)) ? t : __r.neg(t))" in src


def test_untraceable_closures_throw_strings(host):
    host.it.set_global("the_data", to_js(host.it, [1.0, 2.0]))
    with pytest.raises(JSThrow, match="use mcmc.where"):
        host.run("""new mcmc.AmwgSampler({theta: {type: "real"}}, function (s, d) { if (s.theta > 0.5) { return ld.bern(1, 0.5); } return ld.bern(1, s.theta); }, the_data, {});""")
    with pytest.raises(JSThrow, match="returned undefined"):
        host.run("""new mcmc.AmwgSampler({theta: {type: "real"}}, function (s, d) { var t = ld.norm(s.theta, 0, 1); }, the_data, {});""")
    with pytest.raises(JSThrow, match="not visible to the sampler"):
        host.run("""new mcmc.AmwgSampler({theta: {type: "real"}}, function (s, d) { return nowhere_defined(s.theta); }, the_data, {});""")


def test_complete_params_and_descriptor_match_the_reference_fixtures(host, pkg):
    """complete_params on the reference's golden fixtures (tests/test_data.js:9-74), and the descriptor js/mcmc.js hands to the
    native layer against the one the Python host builds (same options quirks: `a || b` merge, per-component arrays)."""
    import json
    import os
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_js.json")))
    gp = G["complete_params"]
    import golden_util as gu
    for which in ("params1", "params2"):
        params = gu.unhex(gp[which]) if hasattr(gu, "unhex") else gp[which]
        host.it.set_global("the_params", to_js(host.it, {k: {kk: vv for kk, vv in v.items() if vv != "<function>"} for k, v in params.items()}))
        host.run("var completed = mcmc.complete_params(the_params);")
        got = to_py(host.get("completed"))
        want = pkg.mcmc.complete_params({k: {kk: vv for kk, vv in v.items() if vv != "<function>"} for k, v in params.items()})
        for name in want:
            for key in ("type", "dim", "lower", "upper", "init"):
                assert got[name][key] == want[name][key] or np.allclose(np.asarray(got[name][key], dtype=float), np.asarray(want[name][key], dtype=float)), (name, key)
    # option merge: global + per-parameter override, falsy values fall through (mcmc.js:871-878)
    host.recorder.created.clear()
    host.it.set_global("the_data", to_js(host.it, NB12))
    host.run("""
      var s = new mcmc.AmwgSampler({p1: {type: "real", lower: 0, upper: 1}, n1: {type: "int", lower: 1, init: 1}, m: {type: "binary"}}, complex_model_post, the_data,
                                   {max_adaptation: 0.5, batch_size: 20, is_adapting: false, params: {p1: {max_adaptation: 0.1, prop_log_scale: 0}}, chains: 8, seed: 3});
    """)
    desc, args = host.recorder.created[-1]
    d = to_py(desc)
    assert args == [8.0, 0.0, 3.0, 0.0]
    opts = {"max_adaptation": 0.5, "batch_size": 20, "is_adapting": False, "params": {"p1": {"max_adaptation": 0.1, "prop_log_scale": 0}}}
    res = pkg.mcmc.resolve_stepper_options(pkg.mcmc.complete_params(models.PARAMS_COMPLEX), opts)
    co = d["comp_options"]
    assert co[0]["max_adaptation"] == res["p1"]["max_adaptation"][0] == 0.1 and co[1]["max_adaptation"] == 0.5
    assert co[0]["batch_size"] == 20 and co[0]["is_adapting"] == (1 if res["p1"]["is_adapting"][0] else 0) == 0
    assert co[0]["prop_log_scale"] == res["p1"]["prop_log_scale"][0] == 0
    assert [p["type"] for p in d["params"]] == [0, 1, 2] and d["init"] == [0.5, 1.0, 1.0]


def test_napi_addon_is_well_formed_and_binds_every_export(pkg):
    """js/amwg_napi.cc cannot be built here (no Node headers): it is syntax-checked against a stub of the Node-API declarations it
    uses, and must bind every AMWG_API export of include/amwg.h."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I" + os.path.join(root, "tests", "napi_stub"), os.path.join(root, "js", "amwg_napi.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(root, "js", "amwg_napi.cc")).read()
    hdr = open(os.path.join(root, "include", "amwg.h")).read()
    exports = re.findall(r"AMWG_API\s+[\w\s\*]+?\b(amwg_\w+)\s*\(", hdr)
    assert sorted(set(exports)) == sorted(pkg._ffi.EXPORTS)
    for name in exports:
        assert name + "(" in src, name + " has no binding in js/amwg_napi.cc"
