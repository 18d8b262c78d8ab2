"""Golden vectors produced by executing the UNMODIFIED reference JS (oracle/minijs) -- no GPU.
They pin (i) the C oracle draw for draw, (ii) the product's host logic (complete_params, option merge, thrown strings)."""
import copy
import ctypes
import math

import numpy as np
import pytest

import golden_util as gu

G = gu.load()


def test_complete_params_matches_the_js(pkg):
    g = G["complete_params"]
    assert g["params1"] == g["params1_expected"] or gu.unhex(g["params1"]) == gu.unhex(g["params1_expected"])
    import test_host_logic as hl
    got1 = pkg.mcmc.complete_params(hl.PARAMS1)
    got2 = pkg.mcmc.complete_params(hl.PARAMS2)
    assert got1 == gu.unhex(g["params1"]) and got2 == gu.unhex(g["params2"])
    # same key sets as the reference's expected fixtures; values equal (JS property order is not part of the contract)
    for name in got2:
        assert set(got2[name]) == set(gu.unhex(g["params2_expected"])[name])


def test_param_init_fixed_matches_the_js(pkg, orc):
    for t, lo, hi, want in G["param_init_fixed"]["values"]:
        lo, hi, want = float.fromhex(lo), float.fromhex(hi), float.fromhex(want)
        assert pkg.mcmc.param_init_fixed(t, lo, hi) == want
        assert orc.param_init_fixed(t, lo, hi) == want
    for t, lo, hi, msg in G["param_init_fixed"]["throws"]:
        with pytest.raises(pkg.JsThrow) as e:
            pkg.mcmc.param_init_fixed(t, float.fromhex(lo), float.fromhex(hi))
        assert e.value.message == msg


def test_thrown_strings_match_the_js(pkg):
    import models
    m, ld = pkg.mcmc, pkg.ld
    t = G["throws"]
    with pytest.raises(pkg.JsThrow) as e:
        m.AmwgSampler({"x": {"type": "real", "dim": [2, 2]}}, models.multivar_norm_dens(ld), None, {"prop_log_scale": [1, 2, 3]})
    assert e.value.message == t["option_dim"]
    with pytest.raises(pkg.JsThrow) as e:
        m.AmwgSampler({"q": {"type": "complex"}}, models.norm_dens(ld))
    assert e.value.message == t["bad_type"]
    with pytest.raises(pkg.JsThrow) as e:
        m.AmwgSampler({"q": {"type": "real", "lower": 2, "upper": 1}}, models.norm_dens(ld))
    assert e.value.message == t["bad_bounds"]


def test_oracle_ld_matches_the_js_bit_for_bit(orc):
    L = orc.lib()
    loose = {"t", "weibull"}            # Math.pow with a non-integer exponent: libm pow in both, may differ from V8 by an ulp
    n = 0
    for fname, rows in G["ld"].items():
        if fname in ("bivarnorm", "dirichlet", "cat"):
            args, want = rows[0][0], float.fromhex(rows[0][1])
            if fname == "bivarnorm":
                a = [np.array(v, dtype=np.float64) for v in args[:3]]
                got = L.orc_ld_bivarnorm(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, args[3])
            elif fname == "dirichlet":
                a = [np.array(v, dtype=np.float64) for v in args]
                got = L.orc_ld_dirichlet(a[0].ctypes.data, a[1].ctypes.data, len(args[0]))
            else:
                a = np.array(args[1], dtype=np.float64)
                got = L.orc_ld_cat(args[0], a.ctypes.data, len(args[1]))
            assert got == want, fname
            continue
        f = getattr(L, "orc_ld_" + fname)
        for hexargs, hexwant in rows:
            args, want = [float.fromhex(a) for a in hexargs], float.fromhex(hexwant)
            got = f(*args)
            if fname in loose:
                assert got == want or abs(got - want) <= 4 * abs(np.spacing(want)), (fname, args)
            else:
                assert got == want or (math.isnan(got) and math.isnan(want)), (fname, args, got, want)
            n += 1
    assert n > 500


def test_oracle_helpers_match_the_js(orc):
    L, h = orc.lib(), G["helpers"]
    pos = ctypes.c_uint64(0)
    assert [L.orc_rnorm(h["seed"], h["chain"], ctypes.byref(pos), 10.0, 5.0) for _ in range(300)] == gu.unhex(h["rnorm_10_5"])
    assert [L.orc_runif(h["seed"], h["chain"], ctypes.byref(pos), 2.0, 5.0) for _ in range(20)] == gu.unhex(h["runif_2_5"])
    assert [L.orc_runif_discrete(h["seed"], h["chain"], ctypes.byref(pos), 1.0, 6.0) for _ in range(20)] == gu.unhex(h["runif_discrete_1_6"])
    assert pos.value == h["uniforms_consumed"]


@pytest.mark.parametrize("case", G["samplers"], ids=lambda c: f"{c['name']}-chain{c['chain']}")
def test_oracle_reproduces_the_js_sampler_draw_for_draw(case, pkg, orc):
    """Every sampler scenario run by the real mcmc.js: same draws, same adaptation state, same number of Math.random() calls.
    The per-component options handed to the oracle come from the PRODUCT's option merge, so that is pinned here as well."""
    c_model, _py, params, _data, data_c = gu.resolve_case(case, pkg)
    m = pkg.mcmc
    cp = m.complete_params(params)
    assert {k: {kk: vv for kk, vv in v.items()} for k, v in cp.items()} == gu.unhex(case["completed_params"])
    resolved = m.resolve_stepper_options(cp, copy.deepcopy(case["options"]))
    comp_options = {name: r for name, r in resolved.items() if r}
    thin0 = (case["options"] or {}).get("thin", 1)
    s = orc.OracleSampler(c_model, data_c, params, seed=case["seed"], chain=case["chain"], comp_options=comp_options, thin=thin0)
    monitor = None
    results = iter(case["results"])
    for step in case["script"]:
        op = step[0]
        if op == "burn": s.burn(step[1])
        elif op == "thin": s.set_thin(step[1])
        elif op == "monitor": monitor = step[1]
        elif op == "stop_adaptation": s.set_adapting(False)
        elif op == "start_adaptation": s.set_adapting(True)
        elif op == "sample":
            want = gu.unhex(next(results)["draws"])
            got = s.sample(step[1], monitor)
            assert list(got.keys()) == list(want.keys())
            for k in want:
                assert gu.same(got[k], want[k]), (case["name"], k)
    assert s.rng_position() == case["uniforms_consumed"]
    info = s.info()
    for name in cp:
        want = gu.flat_info(case["final_info"].get(name, {}))
        if not want:
            continue
        for c, w in zip(s.entries(name), want):
            assert info[c, 0] == w["prop_log_scale"] and info[c, 2] == w["acceptance_count"], (name, c)
            assert info[c, 3] == w["iterations_since_adaption"] and info[c, 4] == w["batch_count"] and bool(info[c, 1]) == w["is_adapting"]
    st = s.state()
    for name, want in gu.unhex(case["final_state"]).items():
        assert gu.same(st[s.entries(name)], np.asarray(want, dtype=np.float64).reshape(-1)), name


@pytest.mark.parametrize("case", G["steppers"], ids=lambda c: f"{c['class']}-chain{c['chain']}")
def test_oracle_reproduces_the_js_standalone_steppers(case, pkg, orc):
    """tests/test_mcmc_js.R:55-142 drive the steppers directly (`stepper.step()` in a loop). A stand-alone stepper over one
    parameter consumes Math.random exactly like an AmwgSampler over that parameter (the substepper shuffle of a one-element
    array draws nothing), so the oracle sampler reproduces every returned value."""
    import copy as _copy
    su = gu.stepper_setup(case, pkg)
    m = pkg.mcmc
    params = _copy.deepcopy(su["params"])
    for name in params:
        params[name]["init"] = _copy.deepcopy(su["state"][name])
        if su["type"]:
            params[name]["type"] = su["type"]
    cp = m.complete_params(params)
    resolved = (m.resolve_stepper_options if case["class"] == "AmwgStepper" else m._resolve_direct)(cp, _copy.deepcopy(su["options"]))
    s = orc.OracleSampler(su["c_model"], su["data_c"], params, seed=case["seed"], chain=case["chain"],
                          comp_options={n: r for n, r in resolved.items() if r})
    results = iter(case["results"])
    for step in case["script"]:
        if step[0] == "stop_adaptation": s.set_adapting(False)
        elif step[0] == "start_adaptation": s.set_adapting(True)
        else:
            want = gu.unhex(next(results))
            for k in range(step[1]):
                s.step()
                st = s.state()
                if case["class"] == "AmwgStepper":
                    got = {n: st[s.entries(n)][0] for n in cp}
                    assert got == want[k], k
                else:
                    assert gu.same(st[: s.D], np.asarray(want[k], dtype=np.float64).reshape(-1)), k
    assert s.rng_position() == case["uniforms_consumed"]
