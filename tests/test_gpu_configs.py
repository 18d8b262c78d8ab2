"""GPU tests of the BASELINE configs: config 2 at full size against the exact posterior (KS < 0.01), configs 4 and 5 at
oracle-sized inputs (bit-exact in faithful mode, statistical with the factorised plates), properties at larger sizes."""
import numpy as np
import pytest

import models
from conftest import config2_data

pytestmark = pytest.mark.gpu


def _grid_posterior_cdfs(x):
    """Exact marginal CDFs of (mu, sigma) for the README model (N(0,100) x U(0,100) priors), by grid integration."""
    n, xbar = x.size, x.mean()
    S = float(np.sum((x - xbar) ** 2))
    s = np.sqrt(S / (n - 1))
    mu = np.linspace(xbar - 9 * s / np.sqrt(n), xbar + 9 * s / np.sqrt(n), 1201)
    sg = np.linspace(s * 0.72, s * 1.38, 1201)
    M, SG = np.meshgrid(mu, sg, indexing="ij")
    lp = -0.5 * (M / 100.0) ** 2 - n * np.log(SG) - (S + n * (xbar - M) ** 2) / (2 * SG ** 2)
    p = np.exp(lp - lp.max())
    pm, ps = p.sum(axis=1), p.sum(axis=0)
    return (mu, np.cumsum(pm) / pm.sum()), (sg, np.cumsum(ps) / ps.sum())


def _ks_against_cdf(samples, grid, cdf):
    xs = np.sort(samples)
    F = np.interp(xs, grid, cdf, left=0.0, right=1.0)
    n = xs.size
    return max(np.max(np.arange(1, n + 1) / n - F), np.max(F - np.arange(0, n) / n))


def test_config2_full_size_ks_against_exact_posterior(gpu_pkg):
    """BASELINE config 2: N=1024, 2^20 chains. One draw per chain after burn-in, pooled: KS < 0.01 on both marginals
    (north_star's tolerance) against the exact posterior; the reference's own tests compare against JAGS at p > 0.01."""
    x = config2_data()
    s = gpu_pkg.mcmc.AmwgSampler(models.PARAMS_NORM, models.norm_post_readme(gpu_pkg.ld), x.tolist(), {"chains": 1 << 20, "seed": 0})
    assert "plate NORM_IID n=1024" in s.program_summary()
    assert s.jit_status()[0], s.jit_status()[1]            # 2^20 chains: the run-time specialised sweep (the production path)
    s.burn(3000)
    d = s.sample(1)
    (gm, cm), (gs, cs) = _grid_posterior_cdfs(x)
    ks_mu = _ks_against_cdf(d["mu"].reshape(-1), gm, cm)
    ks_sg = _ks_against_cdf(d["sigma"].reshape(-1), gs, cs)
    assert ks_mu < 0.01 and ks_sg < 0.01, (ks_mu, ks_sg)
    # every chain is in the support and has adapted its proposal scale
    assert np.all(d["sigma"] > 0) and np.all(np.isfinite(d["mu"]))
    info = s.info()["steppers"][0]
    assert info["mu"]["batch_count"] == 60 and np.all(np.abs(info["mu"]["prop_log_scale"] - np.log(0.14 * 2.4)) < 1.5)


def _hier(J, per, seed):
    rng = np.random.default_rng(seed)
    g = np.repeat(np.arange(J), per)
    y = rng.normal(100, 20, J)[g] + rng.normal(0, 5, J * per)
    params = {"mu": {"type": "real", "dim": [J]}, "sigma": {"type": "real", "lower": 0}}
    return y, g, params


def hier_post(ld, J):
    def f(state, d):                               # BASELINE config 4
        lp = 0
        for j in range(J):
            lp += ld.norm(state.mu[j], 0, 100)
        lp += ld.unif(state.sigma, 0, 100)
        for i in range(len(d.y)):
            lp += ld.norm(d.y[i], state.mu[d.g[i]], state.sigma)
        return lp
    return f


def test_config4_hierarchical_normal(gpu_pkg, orc):
    J, per = 8, 32
    y, g, params = _hier(J, per, 64)
    data = {"y": y.tolist(), "g": g.tolist()}
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    # faithful: every draw equals the oracle (random top-level order over mu[0..J), D = J+1 components)
    s = mcmc.AmwgSampler(params, hier_post(ld, J), data, {"chains": 64, "seed": 5, "faithful": True})
    s.burn(60)
    got = s.sample(40)
    ref = orc.run_model("hier_norm", {"y": y, "g": g}, params, chains=64, seed=5, burn=60, sample=40)
    assert got["mu"].shape == (40, 64, J)
    assert np.array_equal(got["mu"], ref["mu"]) and np.array_equal(got["sigma"], ref["sigma"])
    # fast path (one factorised plate per group): same posterior
    s = mcmc.AmwgSampler(params, hier_post(ld, J), data, {"chains": 4096, "seed": 6})
    assert [x for x in s.program_summary() if x.startswith("plate")] == [f"plate NORM_IID n={per}"] * J
    assert any(x.startswith("pre-evaluated statistics") for x in s.program_summary())
    s.burn(1500)
    fast = s.sample(1)
    ref = orc.run_model("hier_norm", {"y": y, "g": g}, params, chains=1024, seed=6, burn=1500, sample=1)
    from scipy import stats
    for j in (0, J - 1):
        assert stats.ks_2samp(fast["mu"][0, :, j], ref["mu"][0, :, j]).statistic < 0.06
        assert abs(fast["mu"][0, :, j].mean() - y[g == j].mean()) < 0.2
    assert stats.ks_2samp(fast["sigma"].reshape(-1), ref["sigma"].reshape(-1)).statistic < 0.06


MODES = (("stat", {}),                                                     # pre-evaluated statistics: one data pass per sweep
         ("full", {"AMWG_STAT_SWEEP": "0"}),                               # same programs, every step evaluates the full one
         ("block", {"AMWG_STAT_LOWERING": "0"}),                           # block step for mu (one evaluation), term cache
         ("cache", {"AMWG_STAT_LOWERING": "0", "AMWG_BLOCK_STEPS": "0"}),   # per-component programs with the term cache
         ("plain", {"AMWG_STAT_LOWERING": "0", "AMWG_TERM_CACHE": "0"}),   # the plain full program with single-op plates
         ("block_l2", {"AMWG_STAT_LOWERING": "0", "AMWG_PHASE_SYNC": "0"}))


def _run_modes(gpu_pkg, modes, make, burn, sample):
    import os
    out = {}
    for name, env in modes:
        os.environ.update(env)
        try:
            s = make()
            out[name + "_summary"] = s.program_summary()
            s.burn(burn)
            out[name] = s.sample(sample)
            out[name + "_info"] = s.info()["steppers"][0]
            out[name + "_state"] = s.state
        finally:
            for k in env:
                del os.environ[k]
    return out


def test_evaluation_modes_are_bit_identical(gpu_pkg):
    """config-4 shape. However the work is organised -- one data pass per sweep with pre-evaluated statistics, one evaluation per
    block, per-component programs over the term cache, or the full program at every step -- the terms are the same values added in
    the same order and the uniforms are consumed in the same order: every draw is identical."""
    J, per = 8, 32
    y, g, params = _hier(J, per, 65)
    data = {"y": y.tolist(), "g": g.tolist()}
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    out = _run_modes(gpu_pkg, MODES, lambda: mcmc.AmwgSampler(params, hier_post(ld, J), data, {"chains": 500, "seed": 12}), 120, 60)
    assert any(x.startswith("pre-evaluated statistics: 8 plate(s), 256 points") for x in out["stat_summary"])
    assert any(x.startswith("block steps") for x in out["block_summary"])
    for other in [m for m, _ in MODES[1:]]:
        assert np.array_equal(out["stat"]["mu"], out[other]["mu"]) and np.array_equal(out["stat"]["sigma"], out[other]["sigma"]), other
        assert np.array_equal(out["stat_info"]["mu"]["prop_log_scale"], out[other + "_info"]["mu"]["prop_log_scale"]), other
        assert np.array_equal(out["stat_info"]["sigma"]["acceptance_count"], out[other + "_info"]["sigma"]["acceptance_count"]), other
    assert out["stat"]["mu"].std() > 0


def test_statistics_sweep_on_the_headline_shape_with_int_bounds_thin_and_derived(gpu_pkg):
    """config-2 shape (two scalar parameters, N=1024, sigma bounded below: out-of-bounds proposals draw no uniform), an int mean
    (Math.round proposals), thinning, a derived quantity and a chain count that leaves the last CTA ragged."""
    import models
    from conftest import config2_data
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    data = config2_data().tolist()
    two = (("stat", {}), ("full", {"AMWG_STAT_SWEEP": "0"}), ("plain", {"AMWG_STAT_LOWERING": "0"}))
    pars = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}
    out = _run_modes(gpu_pkg, two, lambda: mcmc.AmwgSampler(pars, models.norm_post_test(ld), data, {"chains": 1000, "seed": 3, "thin": 3}), 230, 100)
    assert out["stat"]["mu"].shape == (34, 1000) and set(out["stat"]) == {"mu", "sigma", "var"}
    assert any(x.startswith("pre-evaluated statistics") for x in out["stat_summary"]) and len(out["plain_summary"]) == 3
    for other in ("full", "plain"):
        for k in ("mu", "sigma", "var"):
            assert np.array_equal(out["stat"][k], out[other][k]), (other, k)
        assert np.array_equal(out["stat_state"]["var"], out[other + "_state"]["var"])
        assert np.array_equal(out["stat_info"]["sigma"]["prop_log_scale"], out[other + "_info"]["sigma"]["prop_log_scale"])
    assert out["stat_state"]["mu"].std() > 0
    pars = {"mu": {"type": "int", "lower": 150, "upper": 200, "init": 170}, "sigma": {"type": "real", "lower": 0, "upper": 50}}
    out = _run_modes(gpu_pkg, two, lambda: mcmc.AmwgSampler(pars, models.norm_post_readme(ld), data, {"chains": 333, "seed": 4}), 150, 80)
    for other in ("full", "plain"):
        assert np.array_equal(out["stat"]["mu"], out[other]["mu"]) and np.array_equal(out["stat"]["sigma"], out[other]["sigma"]), other
    assert np.all(out["stat"]["mu"] == np.round(out["stat"]["mu"])) and out["stat"]["mu"].std() > 0


def test_data_larger_than_shared_memory_streams_through_the_tile_ring(gpu_pkg):
    """config-4 shape with data that does not fit in shared memory (8 groups x 4096 points = 256 KB): the sweep's single data pass
    (pre-evaluated statistics) and the block step's single evaluation are CTA-uniform, so the column goes through the TMA tile ring;
    one tile per group here, so the sums are the same as on the L2 path -- identical draws."""
    J, per = 8, 4096
    y, g, params = _hier(J, per, 66)
    data = {"y": y.tolist(), "g": g.tolist()}
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    modes = (("stat_ring", {}), ("full_l2", {"AMWG_STAT_SWEEP": "0", "AMWG_PHASE_SYNC": "0"}),
             ("block_ring", {"AMWG_STAT_LOWERING": "0"}), ("block_l2", {"AMWG_STAT_LOWERING": "0", "AMWG_PHASE_SYNC": "0"}))
    out = _run_modes(gpu_pkg, modes, lambda: mcmc.AmwgSampler(params, hier_post(ld, J), data, {"chains": 256, "seed": 13}), 60, 20)
    for other in ("full_l2", "block_ring", "block_l2"):
        assert np.array_equal(out["stat_ring"]["mu"], out[other]["mu"]) and np.array_equal(out["stat_ring"]["sigma"], out[other]["sigma"]), other
    assert np.isfinite(out["stat_ring"]["mu"]).all() and out["stat_ring"]["mu"].std() > 0 and out["stat_ring"]["sigma"].std() > 0


def poisreg_post(ld, mcmc, K):
    def f(state, d):                               # BASELINE config 5
        lp = 0
        for k in range(K):
            lp += ld.norm(state.beta[k], 0, 10)
        for i in mcmc.points(len(d.y)):
            eta = 0
            for k in range(K):
                eta += d.X[i][k] * state.beta[k]
            lp += ld.pois(d.y[i], mcmc.Math.exp(eta))
        return lp
    return f


def test_config5_poisson_regression(gpu_pkg, orc):
    K, n = 4, 300
    rng = np.random.default_rng(8)
    X = np.column_stack([np.ones(n), rng.normal(0, 0.5, (n, K - 1))])
    beta_true = rng.normal(0, 0.3, K)
    yy = rng.poisson(np.exp(X @ beta_true)).astype(float)
    params = {"beta": {"type": "real", "dim": [K]}}
    data = {"y": yy.tolist(), "X": X.tolist()}
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    s = mcmc.AmwgSampler(params, poisreg_post(ld, mcmc, K), data, {"chains": 32, "seed": 9, "faithful": True})
    assert s.program_summary()[-1] == f"plate GENERIC n={n} body=LD_POIS"
    s.burn(40)
    got = s.sample(30)
    ref = orc.run_model("pois_reg", {"y": yy, "X": X}, params, chains=32, seed=9, burn=40, sample=30)
    assert np.array_equal(got["beta"], ref["beta"])
    s = mcmc.AmwgSampler(params, poisreg_post(ld, mcmc, K), data, {"chains": 4096, "seed": 10})
    assert s.program_summary()[-1] == f"plate POIS_LOGLIN n={n} K={K}"
    s.burn(1200)
    fast = s.sample(1)["beta"][0]
    ref = orc.run_model("pois_reg", {"y": yy, "X": X}, params, chains=512, seed=10, burn=1200, sample=1)["beta"][0]
    from scipy import stats
    for k in range(K):
        assert stats.ks_2samp(fast[:, k], ref[:, k]).statistic < 0.08, k
    assert np.all(np.abs(fast.mean(axis=0) - beta_true) < 0.25)


def test_data_larger_than_shared_memory_is_served_from_l2(gpu_pkg):
    """N = 40000 fp64 (320 KB > the 200 KB staging budget): the column stays in global memory; same posterior as a run on
    the sufficient statistics would give (exact Normal-model posterior mean of mu = xbar to prior shrinkage)."""
    rng = np.random.default_rng(11)
    x = rng.normal(50.0, 3.0, 40000)
    s = gpu_pkg.mcmc.AmwgSampler(models.PARAMS_NORM, models.norm_post_readme(gpu_pkg.ld), x.tolist(), {"chains": 2048, "seed": 1})
    s.burn(1500)
    d = s.sample(1)
    assert abs(d["mu"].mean() - x.mean()) < 0.005 and abs(d["sigma"].mean() - x.std(ddof=1)) < 0.01
    assert abs(d["mu"].std() - x.std() / np.sqrt(x.size)) < 0.004


def test_poisson_plate_exponential_is_accurate(gpu_pkg):
    """The POIS_LOGLIN plate sums exp(eta_i) with a table-driven exponential (2^(j/256) x degree-4 polynomial, csrc exp_acc): within
    2 ulp of the correctly rounded value over the whole range a linear predictor can take, exact at 0, library exp() beyond +-690."""
    import ctypes as C
    L = gpu_pkg._ffi.lib()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-30, 30, 200000), rng.uniform(-689, 689, 50000), rng.normal(0, 1e-6, 1000),
                        [0.0, -0.0, 1.0, -1.0, 689.9, -689.9, 700.0, -745.0, 710.0, np.inf, -np.inf, np.nan]])
    out = np.empty_like(x)
    assert L.amwg_primitive_eval(5, x.ctypes.data, x.size, 0, 0, out.ctypes.data, 0) == 0
    want = np.exp(x)
    fin = np.isfinite(want) & (want > 0)
    ulp = np.abs(out[fin] - want[fin]) / np.spacing(want[fin])
    assert ulp.max() <= 2.0, ulp.max()
    assert out[x == 0.0].tolist() == [1.0, 1.0]
    assert np.isnan(out[-1]) and out[-2] == 0.0 and out[-3] == np.inf


def test_pooled_oracle_ks_on_the_production_path(gpu_pkg, orc):
    """KS of pooled GPU draws against >= 1e5 (nearly independent) pooled draws of the CPU restatement of mcmc.js on the
    same model and data -- the reference sampler itself as the yardstick, beside the exact-posterior KS above. The GPU side is the
    production lowering (statistics sweep, run-time specialised kernel); the oracle evaluates log_post twice per step, term by term."""
    x = np.random.default_rng(64).normal(184.5, 4.5, 64)
    P = models.PARAMS_NORM
    s = gpu_pkg.mcmc.AmwgSampler(P, models.norm_post_readme(gpu_pkg.ld), x.tolist(), {"chains": 1 << 17, "seed": 5})
    assert s.jit_status()[0], s.jit_status()[1]
    s.burn(2000)
    g = s.sample(1)
    ref = orc.run_model("norm_readme", x, P, chains=512, seed=99, burn=2000, sample=4000, thin=20, threads=8)      # 512 x 200 kept draws
    for name in ("mu", "sigma"):
        a, b = np.sort(g[name].reshape(-1)), np.sort(ref[name].reshape(-1))
        allv = np.concatenate([a, b])
        d = np.max(np.abs(np.searchsorted(a, allv, side="right") / a.size - np.searchsorted(b, allv, side="right") / b.size))
        assert b.size >= 100000 and d < 0.01, (name, d)


def test_poisson_plate_on_the_tensor_core_matches_numpy(gpu_pkg):
    """K = 8 with X streamed through the tile ring: the dot products run as DMMA.8x8x4 (csrc PoisMma). The carried log_post of every
    chain (the plate's value at the last accepted proposal) against a float64 numpy evaluation at the chain's state; n = 20003 rows:
    partial last tile, partial last group of 8 rows; 200 chains: a partial CTA and a partial warp."""
    K, n = 8, 20003
    rng = np.random.default_rng(21)
    X = np.column_stack([np.ones(n), rng.normal(0, 0.4, (n, K - 1))])
    beta_true = np.concatenate([[0.3], rng.normal(0, 0.25, K - 1)])
    yy = rng.poisson(np.exp(X @ beta_true)).astype(float)
    params = {"beta": {"type": "real", "dim": [K]}}
    mcmc, ld = gpu_pkg.mcmc, gpu_pkg.ld
    s = mcmc.AmwgSampler(params, poisreg_post(ld, mcmc, K), {"y": yy.tolist(), "X": X.tolist()}, {"chains": 200, "seed": 3})
    assert s.program_summary()[-1] == f"plate POIS_LOGLIN n={n} K={K}"
    from scipy.special import gammaln
    lfact = gammaln(yy + 1.0).sum()

    def numpy_lp(beta):                                            # [chains, K]
        eta = beta @ X.T
        prior = (-0.5 * np.log(2 * np.pi) - np.log(10.0) - beta * beta / 200.0).sum(axis=1)
        return prior + (eta * yy).sum(axis=1) - np.exp(eta).sum(axis=1) - lfact
    for burn in (0, 25, 60):
        if burn:
            s.burn(burn)
        beta = np.asarray(s.state["beta"], np.float64).reshape(200, K)
        got, want = np.asarray(s.log_post(), np.float64), numpy_lp(beta)
        assert np.allclose(got, want, rtol=2e-12, atol=0), (burn, np.max(np.abs(got - want) / np.abs(want)))
    assert np.unique(beta[:, 1]).size > 150                        # the chains have moved, each on its own
    s.burn(600)
    b = np.asarray(s.state["beta"], np.float64).reshape(200, K)
    assert np.all(np.abs(b.mean(axis=0) - beta_true) < 0.05)
