"""The oracle (CPU restatement of mcmc.js / distributions.js) against known answers and exact posteriors -- no GPU.
These are the tests that pin the checker itself."""
import ctypes
import math

import numpy as np
import pytest
from scipy import stats

from conftest import NORM_DATA, PRESIDENTS, config3_data

PARAMS_NORM = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}


def test_philox_known_answer_vectors(orc):
    """Random123 kat_vectors for philox4x32-10."""
    L = orc.lib()

    def ph(ctr, key):
        c, k, o = np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32), np.zeros(4, dtype=np.uint32)
        L.orc_philox4x32_10(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        return [int(v) for v in o]
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = np.array([L.orc_stream_uniform(5, 9, i) for i in range(20000)])
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert stats.kstest(u, "uniform").pvalue > 0.001


def test_math_log_exp_are_fdlibm_accurate(orc):
    """orc_log / orc_exp (fdlibm algorithms, what V8 ports) stay within 1 ulp of glibc and hit the special cases."""
    L = orc.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([np.exp(rng.uniform(-700, 700, 20000)), rng.uniform(1e-3, 10, 20000)])
    for x in xs:
        a, b = L.orc_log(x), math.log(x)
        assert abs(a - b) <= abs(np.spacing(b))
    for x in np.concatenate([rng.uniform(-745, 709, 20000), rng.uniform(-5, 5, 20000)]):
        a, b = L.orc_exp(x), math.exp(x)
        assert abs(a - b) <= abs(np.spacing(b)) or b < 1e-300
    assert L.orc_log(0.0) == -math.inf and math.isnan(L.orc_log(-1.0)) and L.orc_log(1.0) == 0.0 and L.orc_log(math.inf) == math.inf
    assert L.orc_exp(-math.inf) == 0.0 and L.orc_exp(710.0) == math.inf and L.orc_exp(0.0) == 1.0 and L.orc_exp(-746.0) == 0.0
    assert [L.orc_js_round(v) for v in (-2.5, 2.5, 0.49999999999999994, 7.0)] == [-2.0, 3.0, 0.0, 7.0]


def test_ld_known_answers(orc):
    """SURVEY.md 8(c) table (closed forms computed independently with glibc; fdlibm log differs in the last digits)."""
    L = orc.lib()
    close = lambda a, b: abs(a - b) <= 2e-15 * max(1.0, abs(b))       # noqa: E731
    assert close(L.orc_ld_norm(183, 180, 5), -2.708376445638773)
    assert close(L.orc_ld_norm(0.5, 0, 100), -5.524121219192764)
    assert close(L.orc_ld_unif(1, 0, 100), -4.605170185988091)
    assert close(L.orc_ld_beta(0.3, 2, 2), 0.23111172096338706)
    assert close(L.orc_ld_pois(3, 10), -4.884004190245917)
    assert close(L.orc_ld_bern(1, 0.85), -0.16251892949777494) and close(L.orc_ld_bern(0, 0.85), -1.897119984885881)
    assert close(L.orc_ld_lgamma(0.5), 0.572364942924743) and close(L.orc_ld_lgamma(10), 12.801827480081961)
    assert abs(L.orc_ld_lgamma(100.5) - 361.4355404678821) < 1e-9
    # support checks (distributions.js:105, 222, 229, 283)
    assert L.orc_ld_beta(1.5, 2, 2) == -math.inf and L.orc_ld_unif(101, 0, 100) == -math.inf
    assert L.orc_ld_bern(0.5, 0.3) == -math.inf and L.orc_ld_pois(-1, 3) == -math.inf and L.orc_ld_beta(0.3, 1, 1) == 0
    # the whole surface against scipy's log densities (Lanczos lgamma: ~1e-10 absolute)
    rng = np.random.default_rng(3)
    for _ in range(200):
        x, a, b = rng.uniform(0.05, 5), rng.uniform(0.5, 4), rng.uniform(0.5, 4)
        assert abs(L.orc_ld_gamma(x, a, b) - stats.gamma.logpdf(x, a, scale=1 / b)) < 1e-8
        assert abs(L.orc_ld_t(x, 0.3, b, a + 1) - stats.t.logpdf(x, a + 1, loc=0.3, scale=b)) < 1e-8
        assert abs(L.orc_ld_cauchy(x, a, b) - stats.cauchy.logpdf(x, a, b)) < 1e-10
        assert abs(L.orc_ld_laplace(x, a, b) - stats.laplace.logpdf(x, a, b)) < 1e-10
        assert abs(L.orc_ld_lnorm(x, a, b) - stats.lognorm.logpdf(x, b, scale=math.exp(a))) < 1e-10
        assert abs(L.orc_ld_logis(x, a, b) - stats.logistic.logpdf(x, a, b)) < 1e-10
        assert abs(L.orc_ld_weibull(x, a, b) - stats.weibull_min.logpdf(x, a, scale=b)) < 1e-10
        assert abs(L.orc_ld_invgamma(x, a, b) - stats.invgamma.logpdf(x, a, scale=b)) < 1e-8
        assert abs(L.orc_ld_exp(x, a) - stats.expon.logpdf(x, scale=1 / a)) < 1e-12
        k, n, p = float(rng.integers(0, 15)), float(rng.integers(15, 30)), rng.uniform(0.05, 0.95)
        assert abs(L.orc_ld_binom(k, n, p) - stats.binom.logpmf(k, n, p)) < 1e-8
        assert abs(L.orc_ld_nbinom(k, n, p) - stats.nbinom.logpmf(k, n, p)) < 1e-8
        assert abs(L.orc_ld_pois(k, a * 3) - stats.poisson.logpmf(k, a * 3)) < 1e-8
        assert abs(L.orc_ld_hyper(min(k, 5.0), 12, 9, 6) - stats.hypergeom.logpmf(min(k, 5.0), 21, 12, 6)) < 1e-8


def test_log_post_at_init_known_answers(orc):
    """README model at init (mu=.5, sigma=.5) on the presidents data; test model at init (mu=.5, sigma=1) on norm_data."""
    s = orc.OracleSampler("norm_readme", PRESIDENTS, PARAMS_NORM)
    st = s.state()
    assert list(st) == [0.5, 0.5]
    lp = orc.lib().orc_model_norm_readme
    lp.restype = ctypes.c_double
    lp.argtypes = [ctypes.c_void_p] * 3
    v = lp(st.ctypes.data, ctypes.cast(ctypes.pointer(s._keep[-1]), ctypes.c_void_p), None)
    assert abs(v - (-677441.3872049316)) < 1e-6
    s2 = orc.OracleSampler("norm_test", NORM_DATA, {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0, "init": 1}})
    st2 = s2.state()
    assert list(st2) == [0.5, 1.0, 1.0]                 # derived var = sigma^2 filled by the ctor's log_post call (mcmc.js:963)
    lp2 = orc.lib().orc_model_norm_test
    lp2.restype = ctypes.c_double
    lp2.argtypes = [ctypes.c_void_p] * 3
    assert abs(lp2(st2.ctypes.data, ctypes.cast(ctypes.pointer(s2._keep[-1]), ctypes.c_void_p), None) - (-55428.56867673723)) < 1e-7


def test_rnorm_and_shuffle_helpers(orc):
    """tests/test_mcmc_js.R:48-53: 4500 draws of rnorm(10, 5) look normal; shuffle_array is a uniform permutation."""
    L = orc.lib()
    pos = ctypes.c_uint64(0)
    d = np.array([L.orc_rnorm(1, 0, ctypes.byref(pos), 10.0, 5.0) for _ in range(4500)])
    assert stats.shapiro(d[:4500]).pvalue > 0.001 and stats.ttest_1samp(d, 10).pvalue > 0.001
    assert abs(d.std() - 5) < 0.25
    assert 2 * 4500 <= pos.value < 2 * 4500 * 1.6          # two uniforms per trial, ~1.37 trials per draw
    counts = {}
    for _ in range(6000):
        a = np.arange(3, dtype=np.int32)
        L.orc_shuffle(1, 0, ctypes.byref(pos), a.ctypes.data, 3)
        counts[tuple(a)] = counts.get(tuple(a), 0) + 1
    assert len(counts) == 6 and stats.chisquare(list(counts.values())).pvalue > 0.001


def test_sampler_semantics(orc):
    """Row 0 is the pre-step state; thin; derived quantities; 2 log_post calls per in-bounds step (+1 with derived); the
    substepper order persists across sweeps; stop_adaptation freezes prop_log_scale."""
    s = orc.OracleSampler("norm_test", NORM_DATA, {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0, "init": 1}}, seed=4, thin=3)
    d = s.sample(10)
    assert d["mu"].shape == (4,) and d["mu"][0] == 0.5 and d["sigma"][0] == 1.0 and d["var"][0] == 1.0
    assert np.array_equal(d["var"], d["sigma"] ** 2)
    calls = s.logpost_calls()
    assert 1 + 10 * (2 * 2 + 1) - 2 * 10 * 2 <= calls <= 1 + 10 * (2 * 2 + 1)      # out-of-bounds proposals skip both evals
    s.burn(500)
    info = s.info()
    assert info[0, 4] == 10 and info[0, 3] == 10         # batch_count after 510 sweeps, iterations_since_adaption
    s.set_adapting(False)
    before = s.info()[:, 0].copy()
    s.burn(200)
    assert np.array_equal(s.info()[:, 0], before) and s.info()[0, 3] == 10
    assert sorted(s.substepper_order()) == [0, 1]


def test_exact_posteriors(orc):
    """The reference's own test design (tests/test_mcmc_js.R): draws match an independent ground truth. Ground truth here is
    exact: grid posterior moments of the README model, the conjugate Beta posterior, P(x=1)=0.85 for the binary stepper,
    Pois(10) for the int stepper."""
    r = orc.run_model("norm_readme", PRESIDENTS, PARAMS_NORM, chains=3000, seed=1, burn=1200, sample=1)
    for name, mean, sd in (("mu", 184.4525, 1.6048), ("sigma", 4.8686, 1.4321)):
        x = r[name].reshape(-1)
        assert abs(x.mean() - mean) < 5 * sd / math.sqrt(3000) and abs(x.std() - sd) < 0.1
    y = config3_data()
    k = int(y.sum())
    r = orc.run_model("beta_bern", {"x": y}, {"theta": {"type": "real", "lower": 0, "upper": 1}}, chains=3000, seed=2, burn=600, sample=1)
    assert stats.kstest(r["theta"].reshape(-1), stats.beta(2 + k, 2 + 256 - k).cdf).statistic < 0.03
    r = orc.run_model("bern_dens", None, {"x": {"type": "binary"}}, chains=4000, seed=3, burn=3, sample=1)
    assert abs(r["x"].mean() - 0.85) < 0.02                                     # tests/test_mcmc_js.R:123-130
    r = orc.run_model("poisson_dens", None, {"x": {"type": "int", "lower": 0}}, chains=3000, seed=4, burn=400, sample=1)
    x = r["x"].reshape(-1)
    assert np.all(x == np.round(x)) and abs(x.mean() - 10) < 0.25 and abs(x.var() - 10) < 1.2      # tests/test_mcmc_js.R:67-81
    r = orc.run_model("multi_bern_dens", None, {"x": {"type": "binary", "dim": [2, 2]}}, chains=4000, seed=5, burn=6, sample=1)
    assert abs(r["x"][0, :, 0, 0].mean() - 1.0 / 1.3) < 0.03 and abs(r["x"][0, :, 1, 1].mean() - 1.0 / 1.5) < 0.03   # R:132-142
