"""GPU vs the golden vectors made by executing the unmodified reference JS (oracle/minijs): the CUDA sampler, driven through
the `mcmc` mirror and the C ABI, reproduces what mcmc.js itself computes -- every draw, the adaptation state and the derived
quantities -- when Math.random() is the matched Philox stream. `faithful: True` keeps the likelihood loops term-by-term (no
factorised plate), so the comparison is bit for bit for every scenario, including the Normal models."""
import copy

import numpy as np
import pytest

import golden_util as gu

pytestmark = pytest.mark.gpu
G = gu.load()


@pytest.mark.parametrize("case", G["samplers"], ids=lambda c: f"{c['name']}-chain{c['chain']}")
def test_gpu_reproduces_the_js_sampler_draw_for_draw(case, gpu_pkg):
    _c, py_model, params, data, _dc = gu.resolve_case(case, gpu_pkg)
    opts = copy.deepcopy(case["options"]) or {}
    opts.update({"seed": case["seed"], "first_chain": case["chain"], "chains": 1, "faithful": True})
    s = gpu_pkg.mcmc.AmwgSampler(copy.deepcopy(params), py_model, data, opts)
    results = iter(case["results"])
    for step in case["script"]:
        op = step[0]
        if op == "burn": s.burn(step[1])
        elif op == "thin": s.thin(step[1])
        elif op == "monitor": s.monitor(step[1])
        elif op == "stop_adaptation": s.stop_adaptation()
        elif op == "start_adaptation": s.start_adaptation()
        elif op == "sample":
            want = gu.unhex(next(results)["draws"])
            got = s.sample(step[1])
            assert list(got.keys()) == list(want.keys())
            for k in want:
                assert gu.same(got[k], want[k]), (case["name"], k)
    info = s.info()["steppers"][0]
    for name, p in s.params.items():
        want = gu.flat_info(case["final_info"].get(name, {}))
        if not want:
            assert info[name] == {}
            continue
        g = info[name]
        assert gu.same(np.asarray(g["prop_log_scale"]).reshape(-1), [w["prop_log_scale"] for w in want])
        assert gu.same(np.asarray(g["acceptance_count"]).reshape(-1), [w["acceptance_count"] for w in want])
        assert gu.same(np.asarray(g["iterations_since_adaption"]).reshape(-1), [w["iterations_since_adaption"] for w in want])
        assert gu.same(np.asarray(g["batch_count"]).reshape(-1), [w["batch_count"] for w in want])
        assert np.asarray(g["is_adapting"]).reshape(-1).tolist() == [w["is_adapting"] for w in want]
    st = s.state
    for name, want in gu.unhex(case["final_state"]).items():
        assert gu.same(np.asarray(st[name]).reshape(-1), np.asarray(want, dtype=np.float64).reshape(-1)), name


def test_gpu_ld_matches_the_js_bit_for_bit(gpu_pkg):
    ld = gpu_pkg.ld
    for fname, rows in G["ld"].items():
        if fname in ("bivarnorm", "dirichlet", "cat"):
            continue                                   # array arguments: test_gpu_ld_array_functions_match_the_js
        args = np.array([[float.fromhex(a) for a in r[0]] for r in rows])
        want = np.array([float.fromhex(r[1]) for r in rows])
        got = getattr(ld, fname)(*[args[:, k] for k in range(args.shape[1])])
        if fname in ("t", "weibull"):
            # these two go through Math.pow with a non-integer exponent. pow is the one Math function that is NOT pinned across the
            # three sides (the goldens used the host's libm pow, the device uses CUDA's, V8 its own fdlibm port; all faithfully
            # rounded): equal to a few ulp of the result instead of bit for bit
            ok = (got == want) | (np.isnan(got) & np.isnan(want)) | (np.abs(got - want) <= 8 * np.spacing(np.abs(want)))
            assert ok.all(), (fname, args[~ok][:3], got[~ok][:3], want[~ok][:3])
        else:
            assert gu.same(got, want), fname


def _js_args(v):
    if isinstance(v, list):
        return [_js_args(x) for x in v]
    return float.fromhex(v) if isinstance(v, str) else float(v)


def test_gpu_ld_array_functions_match_the_js(gpu_pkg):
    """ld.bivarnorm / ld.dirichlet / ld.cat take array arguments (distributions.js:123-134, 203-215, 232-238): the golden rows
    keep them nested."""
    ld = gpu_pkg.ld
    for fname in ("bivarnorm", "dirichlet", "cat"):
        for r in G["ld"].get(fname, []):
            args = _js_args(r[0])
            want = float.fromhex(r[1])
            got = getattr(ld, fname)(*args)
            got = float(np.asarray(got).reshape(-1)[0])
            assert got == want or (np.isnan(got) and np.isnan(want)) or abs(got - want) <= 8 * np.spacing(abs(want)), (fname, args, got, want)


def test_exported_helpers_match_the_js(gpu_pkg):
    """mcmc.rnorm / runif / runif_discrete (mcmc.js:31-54) of the PRODUCT on the golden stream: same values and the same number of
    Math.random() calls as the reference's own helpers."""
    mcmc = gpu_pkg.mcmc
    h = G["helpers"]
    mcmc.set_random_stream(h["seed"], h["chain"])
    got = [mcmc.rnorm(10, 5) for _ in h["rnorm_10_5"]]
    assert gu.same(got, [float.fromhex(v) for v in h["rnorm_10_5"]])
    got = [mcmc.runif(2, 5) for _ in h["runif_2_5"]]
    assert gu.same(got, [float.fromhex(v) for v in h["runif_2_5"]])
    got = [mcmc.runif_discrete(1, 6) for _ in h["runif_discrete_1_6"]]
    assert gu.same(np.asarray(got, dtype=np.float64), [float.fromhex(v) if isinstance(v, str) else float(v) for v in h["runif_discrete_1_6"]])
    assert mcmc._HostStream.n == h["uniforms_consumed"]


@pytest.mark.parametrize("case", [c for c in G["samplers"] if c["log_post"] in ("spike_bern", "complex_model_post")],
                         ids=lambda c: f"{c['name']}-chain{c['chain']}")
def test_literal_if_on_a_binary_parameter_matches_the_js(case, gpu_pkg):
    """The reference's `if (m === 0) ... else ...` (tests/test_data.js:163-168) written as a plain Python `if`: log_post is
    recorded once per value of m and the device picks the program that matches the evaluated state. Same draws as mcmc.js."""
    import models
    _c, _py, params, data, _dc = gu.resolve_case(case, gpu_pkg)
    literal = models.spike_bern_literal(gpu_pkg.ld) if case["log_post"] == "spike_bern" else models.complex_model_post_literal(gpu_pkg.ld)
    opts = copy.deepcopy(case["options"]) or {}
    opts.update({"seed": case["seed"], "first_chain": case["chain"], "chains": 1})
    s = gpu_pkg.mcmc.AmwgSampler(copy.deepcopy(params), literal, data, opts)
    assert len(s._program.variant_logpost) == 2
    results = iter(case["results"])
    for step in case["script"]:
        if step[0] == "burn": s.burn(step[1])
        elif step[0] == "thin": s.thin(step[1])
        elif step[0] == "sample":
            want = gu.unhex(next(results)["draws"])
            got = s.sample(step[1])
            for k in want:
                assert gu.same(got[k], want[k]), (case["name"], k)


@pytest.mark.parametrize("case", G["steppers"], ids=lambda c: f"{c['class']}-chain{c['chain']}")
def test_gpu_standalone_steppers_match_the_js(case, gpu_pkg):
    """mcmc.RealMetropolisStepper & co used directly, as tests/test_mcmc_js.R:55-142 do: `stepper.step()` returns what the
    reference's stepper returns, the caller's `state` object is updated in place, info() matches."""
    import copy as _copy
    mcmc = gpu_pkg.mcmc
    su = gu.stepper_setup(case, gpu_pkg)
    state = gpu_pkg.tracer.State(_copy.deepcopy(su["state"]))
    if case["class"] == "AmwgStepper":
        import models
        x = gu.NB12
        model = models.complex_model_post_literal(gpu_pkg.ld)
        posterior = lambda: model(state, x)                                   # noqa: E731
    else:
        posterior = lambda: su["model"](state)                                # noqa: E731
    opts = dict(_copy.deepcopy(su["options"]) or {})
    opts.update({"seed": case["seed"], "first_chain": case["chain"]})
    stepper = getattr(mcmc, case["class"])(_copy.deepcopy(su["params"]), state, posterior, opts)
    results = iter(case["results"])
    for step in case["script"]:
        if step[0] == "stop_adaptation": stepper.stop_adaptation()
        elif step[0] == "start_adaptation": stepper.start_adaptation()
        else:
            want = gu.unhex(next(results))
            for k in range(step[1]):
                r = stepper.step()
                if case["class"] == "AmwgStepper":
                    assert {n: float(r[n]) for n in want[k]} == want[k], k
                else:
                    assert gu.same(np.asarray(r, dtype=np.float64), np.asarray(want[k], dtype=np.float64)), k
                    assert r is state["x"] or np.isscalar(r) or isinstance(r, float)
    want_state = gu.unhex(case["final_state"])
    for n, v in want_state.items():
        assert gu.same(np.asarray(state[n], dtype=np.float64), np.asarray(v, dtype=np.float64)), n
    info = stepper.info()
    want_info = case["final_info"]
    if case["class"] == "AmwgStepper":
        for n in want_info:
            for w, g in zip(gu.flat_info(want_info[n]), [info[n]] if info[n] else []):
                assert g["prop_log_scale"] == w["prop_log_scale"] and g["batch_count"] == w["batch_count"] and g["acceptance_count"] == w["acceptance_count"]
    else:
        flat_g = []

        def walk(o):
            if isinstance(o, list):
                [walk(v) for v in o]
            elif o:
                flat_g.append(o)
        walk(info)
        for w, g in zip(gu.flat_info(want_info), flat_g):
            assert g["prop_log_scale"] == w["prop_log_scale"] and g["batch_count"] == w["batch_count"]
            assert g["acceptance_count"] == w["acceptance_count"] and g["iterations_since_adaption"] == w["iterations_since_adaption"]
            assert bool(g["is_adapting"]) == w["is_adapting"]
