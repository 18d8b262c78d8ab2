"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

from conftest import NORM_DATA, PRESIDENTS, config2_data, config3_data
import models

pytestmark = pytest.mark.gpu


def _same_bits(a, b):
    """bit-for-bit equality; JS has a single NaN, so any NaN equals any NaN."""
    return bool(np.all((a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))))


def test_primitives_bit_exact(gpu_pkg, orc):
    """Math.log / Math.exp / Math.round / the Philox stream / rnorm: device == oracle, bit for bit."""
    L, O = gpu_pkg._ffi.lib(), orc.lib()
    rng = np.random.default_rng(0)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(1e-3, 10, 200000),
                        [0.0, -1.0, np.inf, 1.0, 5e-324, 2.2250738585072014e-308, np.nan]])
    out = np.empty_like(x)
    gpu_pkg._ffi.check(L.amwg_primitive_eval(0, x.ctypes.data, x.size, 0, 0, out.ctypes.data, 0))
    ref = np.array([O.orc_log(v) for v in x])
    assert _same_bits(out, ref)
    x = np.concatenate([rng.uniform(-745, 710, 200000), rng.uniform(-5, 5, 200000), [0.0, -np.inf, np.inf, 709.9, -745.2, 1e-10, np.nan]])
    gpu_pkg._ffi.check(L.amwg_primitive_eval(1, x.ctypes.data, x.size, 0, 0, out.ctypes.data, 0))
    ref = np.array([O.orc_exp(v) for v in x])
    assert _same_bits(out, ref)
    x = np.concatenate([rng.uniform(-10, 10, 10000), [-2.5, 2.5, 0.5, -0.5, 0.49999999999999994, -0.0, 1e300]])
    out = np.empty_like(x)
    gpu_pkg._ffi.check(L.amwg_primitive_eval(4, x.ctypes.data, x.size, 0, 0, out.ctypes.data, 0))
    assert np.array_equal(out, np.array([O.orc_js_round(v) for v in x]))
    n = 4097
    out = np.empty(n)
    dummy = np.zeros(n)
    gpu_pkg._ffi.check(L.amwg_primitive_eval(2, dummy.ctypes.data, n, 12345, 77, out.ctypes.data, 0))
    assert np.array_equal(out, np.array([O.orc_stream_uniform(12345, 77, i) for i in range(n)]))
    out = np.empty(5000)
    mean_sd = np.array([10.0, 5.0] + [0.0] * 4998)            # keep alive across the call
    gpu_pkg._ffi.check(L.amwg_primitive_eval(3, mean_sd.ctypes.data, 5000, 9, 3, out.ctypes.data, 0))
    import ctypes
    pos = ctypes.c_uint64(0)
    ref = np.array([O.orc_rnorm(9, 3, ctypes.byref(pos), 10.0, 5.0) for _ in range(5000)])
    assert np.array_equal(out, ref)


def test_ld_bit_exact(gpu_pkg, orc):
    """Every scalar ld.* on the device equals the oracle restatement of distributions.js, bit for bit (pow-free ones)."""
    ld, O = gpu_pkg.ld, orc.lib()
    rng = np.random.default_rng(1)
    n = 20000
    cases = {
        "norm": (rng.normal(0, 50, n), rng.normal(0, 50, n), rng.uniform(0.01, 100, n)),
        "unif": (rng.uniform(-1, 2, n), np.zeros(n), np.ones(n)),
        "beta": (rng.uniform(-0.1, 1.1, n), rng.uniform(0.5, 5, n), rng.uniform(0.5, 5, n)),
        "bern": (rng.integers(0, 3, n).astype(float), rng.uniform(0, 1, n)),
        "pois": (rng.integers(-1, 50, n).astype(float), rng.uniform(0.01, 40, n)),
        "lgamma": (rng.uniform(0.01, 200, n),),
        "lfactorial": (rng.integers(-1, 100, n).astype(float),),
        "lchoose": (rng.integers(1, 60, n).astype(float), rng.integers(0, 30, n).astype(float)),
        "lbeta": (rng.uniform(0.1, 30, n), rng.uniform(0.1, 30, n)),
        "cauchy": (rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.1, 5, n)),
        "laplace": (rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.1, 5, n)),
        "gamma": (rng.uniform(-0.5, 20, n), rng.uniform(0.2, 9, n), rng.uniform(0.2, 9, n)),
        "invgamma": (rng.uniform(-0.5, 20, n), rng.uniform(0.2, 9, n), rng.uniform(0.2, 9, n)),
        "lnorm": (rng.uniform(-0.5, 20, n), rng.normal(0, 2, n), rng.uniform(0.2, 3, n)),
        "pareto": (rng.uniform(0.1, 20, n), rng.uniform(0.2, 9, n), rng.uniform(0.2, 9, n)),
        "logis": (rng.normal(0, 5, n), rng.normal(0, 5, n), rng.uniform(0.1, 5, n)),
        "exp": (rng.uniform(-0.5, 20, n), rng.uniform(0.1, 5, n)),
        "binom": (rng.integers(-1, 30, n).astype(float), rng.integers(1, 30, n).astype(float), rng.uniform(0, 1, n)),
        "nbinom": (rng.integers(-1, 30, n).astype(float), rng.integers(1, 30, n).astype(float), rng.uniform(0.01, 0.99, n)),
        "hyper": (rng.integers(0, 10, n).astype(float), rng.integers(10, 30, n).astype(float), rng.integers(10, 30, n).astype(float), rng.integers(5, 10, n).astype(float)),
    }
    for name, args in cases.items():
        got = getattr(ld, name)(*args)
        f = getattr(O, "orc_ld_" + name)
        ref = np.array([f(*[float(a[i]) for a in args]) for i in range(n)])
        same = (got.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(got) & np.isnan(ref))
        assert same.all(), (name, int((~same).sum()), got[~same][:3], ref[~same][:3])
    # known answers of SURVEY 8(c) (closed forms; glibc-vs-fdlibm log differs in the last digits)
    assert abs(ld.norm(183, 180, 5) - (-2.708376445638773)) < 1e-14
    assert abs(ld.unif(1, 0, 100) - (-4.605170185988091)) < 1e-14
    assert abs(ld.pois(3, 10) - (-4.884004190245917)) < 1e-14


def _oracle_and_gpu(gpu_pkg, orc, model_py, model_c, params, data_py, data_c, chains, seed, burn, sample, thin=1, options=None):
    mcmc = gpu_pkg.mcmc
    opts = {"chains": chains, "seed": seed, "thin": thin}
    opts.update(options or {})
    s = mcmc.AmwgSampler(params, model_py, data_py, opts)
    s.burn(burn)
    got = s.sample(sample)
    ref = orc.run_model(model_c, data_c, params, chains=chains, seed=seed, burn=burn, sample=sample, thin=thin)
    return s, got, ref


def test_binary_and_real_steppers_bit_exact(gpu_pkg, orc):
    """BASELINE config 3 (N=256 Bernoulli data, theta real in [0,1] + binary indicator m): every draw of every chain,
    including the binary stepper's choices and the Metropolis accept decisions, equals the oracle under the same Philox stream."""
    y = config3_data()
    s, got, ref = _oracle_and_gpu(gpu_pkg, orc, models.spike_bern(gpu_pkg.ld, gpu_pkg.mcmc), "spike_bern", models.PARAMS_SPIKE,
                                  {"x": y.tolist()}, {"x": y}, chains=512, seed=11, burn=150, sample=100)
    assert s.program_summary()[-1].startswith("plate BERN_IID")
    assert got["m"].shape == (100, 512)
    assert np.array_equal(got["m"], ref["m"])
    assert np.array_equal(got["theta"], ref["theta"])


def test_beta_bernoulli_readme_bit_exact(gpu_pkg, orc):
    y = [1, 0, 1, 1, 0, 1, 1, 1]                                     # README.md:200
    s, got, ref = _oracle_and_gpu(gpu_pkg, orc, models.beta_bern(gpu_pkg.ld), "beta_bern", models.PARAMS_THETA,
                                  {"x": y}, {"x": np.array(y, float)}, chains=256, seed=5, burn=0, sample=400, thin=2)
    assert got["theta"].shape == (200, 256)
    assert np.array_equal(got["theta"], ref["theta"])


def test_normal_model_matches_oracle_statistically(gpu_pkg, orc):
    """README model on the presidents data: the factorised Normal plate is not bit-faithful, so compare distributions:
    pooled GPU draws vs pooled oracle draws (two-sample KS) and vs the exact posterior moments."""
    from scipy import stats
    chains = 8192
    s, got, ref = _oracle_and_gpu(gpu_pkg, orc, models.norm_post_readme(gpu_pkg.ld), "norm_readme", models.PARAMS_NORM,
                                  PRESIDENTS, PRESIDENTS, chains=chains, seed=3, burn=1500, sample=1)
    for name, mean, sd in (("mu", 184.4525, 1.6048), ("sigma", 4.8686, 1.4321)):
        g, r = got[name].reshape(-1), ref[name].reshape(-1)
        ks = stats.ks_2samp(g, r).statistic
        assert ks < 0.03, (name, ks)                       # two independent-noise samples of 8192: KS noise ~0.015
        assert abs(g.mean() - mean) < 4 * sd / np.sqrt(chains) + 0.01
        assert abs(g.std() - sd) < 0.08
    # the first steps, before rounding differences can matter, are identical to the oracle draw for draw
    s2 = gpu_pkg.mcmc.AmwgSampler(models.PARAMS_NORM, models.norm_post_readme(gpu_pkg.ld), PRESIDENTS, {"chains": 64, "seed": 3})
    g2 = s2.sample(30)
    r2 = orc.run_model("norm_readme", PRESIDENTS, models.PARAMS_NORM, chains=64, seed=3, burn=0, sample=30)
    agree = np.mean(g2["mu"] == r2["mu"])
    assert agree > 0.99, agree


def test_derived_quantities_thin_monitor(gpu_pkg, orc):
    """tests/test_mcmc_js.R:224-236 shape facts: var == sigma^2 is monitored; thin(k) gives ceil(n/k) rows; monitor() filters."""
    mcmc = gpu_pkg.mcmc
    s = mcmc.AmwgSampler(models.PARAMS1, models.norm_post_test(gpu_pkg.ld), NORM_DATA, {"seed": 2})
    s.burn(200)
    d = s.sample(100)
    assert list(d.keys()) == ["mu", "sigma", "var"]
    assert d["mu"].shape == (100,)
    assert np.array_equal(d["var"], d["sigma"] * d["sigma"])
    s.thin(10)
    assert s.sample(95)["mu"].shape == (10,)
    s.monitor(["sigma"])
    assert list(s.sample(5).keys()) == ["sigma"]
    st = s.step()
    assert set(st.keys()) == {"mu", "sigma", "var"} and st["var"] == st["sigma"] ** 2


def test_adaptation_and_info_match_oracle(gpu_pkg, orc):
    """prop_log_scale / batch_count after burn-in, and stop_adaptation()/start_adaptation(), follow mcmc.js:536-561."""
    y = config3_data()
    mcmc = gpu_pkg.mcmc
    s = mcmc.AmwgSampler(models.PARAMS_THETA, models.beta_bern(gpu_pkg.ld), {"x": y.tolist()}, {"chains": 8, "seed": 21})
    o = [orc.OracleSampler("beta_bern", {"x": y}, models.PARAMS_THETA, seed=21, chain=c) for c in range(8)]
    s.burn(175); [q.burn(175) for q in o]
    s.stop_adaptation(); [q.set_adapting(False) for q in o]
    s.burn(60); [q.burn(60) for q in o]
    s.start_adaptation(); [q.set_adapting(True) for q in o]
    s.burn(40); [q.burn(40) for q in o]
    info = s.info()["steppers"][0]["theta"]
    ref = np.array([q.info()[0] for q in o])          # prop_log_scale, is_adapting, acceptance_count, iterations_since, batch_count
    assert np.array_equal(info["prop_log_scale"], ref[:, 0])
    assert np.array_equal(info["acceptance_count"], ref[:, 2].astype(np.int32))
    assert info["iterations_since_adaption"] == ref[0, 3] == 15
    assert info["batch_count"] == ref[0, 4] == 4
    assert np.array_equal(s.state["theta"], np.array([q.state()[0] for q in o]))


def test_chain_sharding_invariance(gpu_pkg):
    """Chain g's draws depend on (seed, g) only: a handle over chains [64, 128) reproduces that block of a handle over [0, 128)
    -- what makes results independent of how chains are sharded over GPUs (parallel.shard_bounds picks first_chain)."""
    from bayes_js_b200.parallel import shard_bounds
    y = config3_data()
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    full = mcmc.AmwgSampler(models.PARAMS_SPIKE, models.spike_bern(ld, mcmc), {"x": y.tolist()}, {"chains": 128, "seed": 4})
    a = full.sample(60)
    first, count = shard_bounds(128, 1, 2)
    assert (first, count) == (64, 64)
    half = mcmc.AmwgSampler(models.PARAMS_SPIKE, models.spike_bern(ld, mcmc), {"x": y.tolist()}, {"chains": count, "seed": 4, "first_chain": first})
    b = half.sample(60)
    assert np.array_equal(a["theta"][:, 64:], b["theta"])
    assert np.array_equal(a["m"][:, 64:], b["m"])


def test_log_post_method_and_the_rest_of_the_ld_surface(gpu_pkg, orc):
    """`sampler.log_post()` (the closure the Sampler ctor stores, mcmc.js:958-960) and every remaining `ld.*` function, including the
    array-valued ones (bivarnorm, dirichlet, cat), evaluated through a traced log_post against the oracle's restatement."""
    import ctypes
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    O = orc.lib()
    # README model at its deterministic init (mu = sigma = 0.5): SURVEY 8(c) known answer
    s = mcmc.AmwgSampler(models.PARAMS_NORM, models.norm_post_readme(ld), PRESIDENTS, {"seed": 1, "faithful": True})
    ref = orc.OracleSampler("norm_readme", PRESIDENTS, models.PARAMS_NORM)
    fn = O.orc_model_norm_readme
    fn.restype = ctypes.c_double
    fn.argtypes = [ctypes.c_void_p] * 3
    st = ref.state()
    want = fn(st.ctypes.data, ctypes.cast(ctypes.pointer(ref._keep[-1]), ctypes.c_void_p), None)
    assert s.log_post() == want and abs(want - (-677441.3872049316)) < 1e-6
    fast = mcmc.AmwgSampler(models.PARAMS_NORM, models.norm_post_readme(ld), PRESIDENTS, {"seed": 1})
    assert abs(fast.log_post() - want) <= 1e-12 * abs(want)
    fast.burn(25)
    lp = fast.log_post()
    stt = fast.state
    arr = np.array([stt["mu"], stt["sigma"]])
    assert abs(lp - fn(arr.ctypes.data, ctypes.cast(ctypes.pointer(ref._keep[-1]), ctypes.c_void_p), None)) <= 1e-12 * abs(lp)

    def everything(state, data):
        x = state.x
        lp = 0
        lp += ld.bivarnorm([x, 1.5], [0.5, 1.0], [1.2, 0.7], 0.3)
        lp += ld.dirichlet([0.2, 0.3, 0.5], [1.5, x + 1, 3.0])
        lp += ld.cat(2, [0.2, 0.5, 0.3])
        lp += ld.t(x, 0.3, 1.2, 5)
        lp += ld.weibull(x + 1, 1.5, 2.0)
        lp += ld.gamma(x + 1, 2.0, 3.0) + ld.invgamma(x + 1, 2.0, 3.0) + ld.lnorm(x + 1, 0.2, 0.7) + ld.pareto(x + 3, 2.0, 3.0)
        lp += ld.logis(x, 0.1, 0.9) + ld.exp(x, 1.3) + ld.binom(3, 10, x / 2) + ld.nbinom(4, 3, x / 2) + ld.hyper(2, 12, 9, 6)
        lp += ld.cauchy(x, 0.2, 1.1) + ld.laplace(x, 0.2, 1.1) + ld.dexp(x, 0.2, 1.1)
        lp += ld.lgamma(x + 2) + ld.lfactorial(4) + ld.lchoose(7, 3) + ld.lbeta(x + 1, 2.5)
        return lp
    s = mcmc.AmwgSampler({"x": {"type": "real", "lower": 0, "upper": 1}}, everything, None, {"seed": 3})
    x = 0.5
    a = lambda v: np.array(v, dtype=np.float64)          # noqa: E731
    xv, mv, sv = a([x, 1.5]), a([0.5, 1.0]), a([1.2, 0.7])
    dx, da = a([0.2, 0.3, 0.5]), a([1.5, x + 1, 3.0])
    pr = a([0.2, 0.5, 0.3])
    want = 0.0
    want += O.orc_ld_bivarnorm(xv.ctypes.data, mv.ctypes.data, sv.ctypes.data, 0.3)
    want += O.orc_ld_dirichlet(dx.ctypes.data, da.ctypes.data, 3)
    want += O.orc_ld_cat(2.0, pr.ctypes.data, 3)
    want += O.orc_ld_t(x, 0.3, 1.2, 5)
    want += O.orc_ld_weibull(x + 1, 1.5, 2.0)
    want += O.orc_ld_gamma(x + 1, 2.0, 3.0) + O.orc_ld_invgamma(x + 1, 2.0, 3.0) + O.orc_ld_lnorm(x + 1, 0.2, 0.7) + O.orc_ld_pareto(x + 3, 2.0, 3.0)
    want += O.orc_ld_logis(x, 0.1, 0.9) + O.orc_ld_exp(x, 1.3) + O.orc_ld_binom(3, 10, x / 2) + O.orc_ld_nbinom(4, 3, x / 2) + O.orc_ld_hyper(2, 12, 9, 6)
    want += O.orc_ld_cauchy(x, 0.2, 1.1) + O.orc_ld_laplace(x, 0.2, 1.1) + O.orc_ld_laplace(x, 0.2, 1.1)
    want += O.orc_ld_lgamma(x + 2) + O.orc_ld_lfactorial(4) + O.orc_ld_lchoose(7, 3) + O.orc_ld_lbeta(x + 1, 2.5)
    got = s.log_post()
    assert abs(got - want) <= 1e-13 * abs(want), (got, want)
    s.burn(50)
    assert np.isfinite(s.log_post()) and 0 <= s.state["x"] <= 1


def test_more_than_16_named_parameters_and_dim0_beyond_256(gpu_pkg, orc):
    """The reference has no limit on the number of named parameters or on dim[0] (mcmc.js:844-881). Up to 16 / 256 the two
    permutations of a sweep live in a register / a local array; beyond, in per-chain global arrays (amwg_tma.cuh perm_get / ord_get).
    Same draws as the oracle, bit for bit, including the in-place substepper shuffle that persists across sweeps."""
    ld, mcmc = gpu_pkg.ld, gpu_pkg.mcmc
    O = orc.lib()
    P = 23
    params = {"t%d" % k: ({"type": "real"} if k % 3 else {"type": "int", "lower": -50, "upper": 50}) for k in range(P)}

    def lp_many(state, d=None):
        lp = 0
        for k in range(P):
            lp += ld.norm(state["t%d" % k], 0.5 * k, 1 + 0.1 * k)
        return lp
    s = mcmc.AmwgSampler(params, lp_many, None, {"chains": 3, "seed": 41, "first_chain": 7})
    got = s.sample(60)
    for c in range(3):
        o = orc.OracleSampler(lambda st: sum(O.orc_ld_norm(st[k], 0.5 * k, 1 + 0.1 * k) for k in range(P)), None, params, seed=41, chain=7 + c)
        ref = o.sample(60)
        for k in range(P):
            assert np.array_equal(got["t%d" % k][:, c], ref["t%d" % k]), (c, k)
    J = 300
    params = {"x": {"type": "real", "dim": [J]}, "s": {"type": "real", "lower": 0}}

    def lp_wide(state, d=None):
        lp = ld.gamma(state.s, 2, 1)
        for j in range(J):
            lp += ld.norm(state.x[j], 0.01 * j, state.s)
        return lp
    s = mcmc.AmwgSampler(params, lp_wide, None, {"chains": 2, "seed": 43})
    got = s.sample(25)

    def f(st):
        v = O.orc_ld_gamma(st[J], 2, 1)
        for j in range(J):
            v += O.orc_ld_norm(st[j], 0.01 * j, st[J])
        return v
    for c in range(2):
        o = orc.OracleSampler(f, None, params, seed=43, chain=c)
        ref = o.sample(25)
        assert np.array_equal(got["x"][:, c], ref["x"]) and np.array_equal(got["s"][:, c], ref["s"]), c
