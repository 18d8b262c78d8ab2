"""Randomly composed models through the code generator, executed on the host (tests/test_jit_codegen_semantics.py's harness).

Each seed builds a log_post out of the operations a user can write -- arithmetic, Math.log/exp/sqrt/abs/pow/min/max, comparisons and
mcmc.where, every scalar ld.*, loops over data columns with data-indexed parameters, a Bernoulli column, derived quantities, real / int /
binary / vector parameters -- lowers it with the product's tracer, specialises it (csrc/amwg_jit.cuh), compiles the generated code for
the host and compares jit_logpost() / jit_derived() with the bytecode run on the oracle's arithmetic at random states, bit for bit.
Half of the models are lowered as with AMWG_TERM_CACHE=0 (every step runs the full program); of the others, those the generator
turns down (they step with a term cache) are counted, not failed."""
import numpy as np
import pytest

from test_jit_codegen_semantics import HostProgram


def _build(pkg, seed):
    ld, mcmc = pkg.ld, pkg.mcmc
    M = mcmc.Math
    rng = np.random.default_rng(seed)
    params = {"theta": {"type": "real", "lower": 0, "upper": 1}, "mu": {"type": "real"}, "s": {"type": "real", "lower": 0, "init": 2},
              "n": {"type": "int", "lower": 1, "init": 3}, "m": {"type": "binary"}, "v": {"type": "real", "dim": [3]}}
    n_pts = int(rng.integers(5, 40))
    data = {"y": rng.normal(1, 2, n_pts).tolist(), "g": rng.integers(0, 3, n_pts).astype(float).tolist(),
            "x": (rng.random(n_pts) < 0.6).astype(float).tolist(), "k": rng.poisson(3, n_pts).astype(float).tolist()}
    plan = {"terms": [int(v) for v in rng.integers(0, 1 << 30, int(rng.integers(2, 7)))], "loops": [int(v) for v in rng.integers(0, 4, int(rng.integers(0, 3)))],
            "derived": int(rng.integers(0, 3)), "seed": seed}

    def log_post(state, d):
        r = np.random.default_rng(plan["seed"] + 1)               # the same choices every time the closure is traced

        def leaf():
            k = int(r.integers(0, 9))
            if k == 0: return state.theta
            if k == 1: return state.mu
            if k == 2: return state.s
            if k == 3: return state.n
            if k == 4: return state.m
            if k == 5: return state.v[int(r.integers(0, 3))]
            return float(np.round(r.normal(0, 2), 3))

        def expr(depth):
            if depth <= 0 or r.random() < 0.25:
                return leaf()
            k = int(r.integers(0, 14))
            a = expr(depth - 1)
            if k == 0: return a + expr(depth - 1)
            if k == 1: return a - expr(depth - 1)
            if k == 2: return a * expr(depth - 1)
            if k == 3: return a / (M.abs(expr(depth - 1)) + 0.5)
            if k == 4: return M.log(M.abs(a) + 0.1)
            if k == 5: return M.exp(a * 0.1)
            if k == 6: return M.sqrt(M.abs(a))
            if k == 7: return M.pow(M.abs(a) + 0.1, float(r.choice([2.0, 0.5, 3.0, 1.7])))
            if k == 8: return M.max(a, expr(depth - 1))
            if k == 9: return M.min(a, expr(depth - 1))
            if k == 10: return mcmc.where(a > expr(depth - 1), expr(depth - 1), a)
            if k == 11: return mcmc.where(state.m == 0, a, expr(depth - 1))
            if k == 12: return mcmc.where((a <= 0.3) | (state.n >= 4), 1.5, a) if hasattr(a, "__or__") and not isinstance(a, float) else a
            return -a

        pos = lambda e: M.abs(e) + 0.2
        prob = lambda e: 1 / (1 + M.exp(-e))
        lp = 0
        for t in plan["terms"]:
            k = t % 22
            e1, e2 = expr(2), expr(2)
            if k == 0: lp += ld.norm(e1, e2, pos(expr(1)))
            elif k == 1: lp += ld.beta(state.theta, pos(e1), pos(e2))
            elif k == 2: lp += ld.pois(state.n, pos(e1))
            elif k == 3: lp += ld.bern(state.m, prob(e1))
            elif k == 4: lp += ld.binom(state.n, state.n + 4, prob(e1))
            elif k == 5: lp += ld.gamma(state.s, pos(e1), pos(e2))
            elif k == 6: lp += ld.unif(state.theta, -0.5, pos(e1))
            elif k == 7: lp += ld.exp(state.s, pos(e1))
            elif k == 8: lp += ld.cauchy(e1, e2, pos(expr(1)))
            elif k == 9: lp += ld.laplace(e1, e2, pos(expr(1)))
            elif k == 10: lp += ld.logis(e1, e2, pos(expr(1)))
            elif k == 11: lp += ld.lnorm(state.s, e1, pos(e2))
            elif k == 12: lp += ld.nbinom(state.n, pos(e1), prob(e2))
            elif k == 13: lp += ld.t(e1, e2, pos(expr(1)), pos(expr(1)))
            elif k == 14: lp += ld.weibull(state.s, pos(e1), pos(e2))
            elif k == 15: lp += ld.pareto(state.s + 1, pos(e1) * 0.1, pos(e2))
            elif k == 16: lp += ld.invgamma(state.s, pos(e1), pos(e2))
            elif k == 17: lp += ld.hyper(state.n, state.n + 6, 9, 5)
            elif k == 18: lp += ld.lgamma(pos(e1)) - ld.lfactorial(state.n) + ld.lchoose(state.n + 3, 2) - ld.lbeta(pos(e2), 1.5)
            elif k == 19: lp += e1 * 0.01
            elif k == 20: lp += ld.norm(state.v[0], state.v[1] * state.m, pos(state.v[2]))
            else: lp += ld.beta(prob(e1), 2, 3)
        for kind in plan["loops"]:
            for i in range(len(d.y)):
                if kind == 0: lp += ld.norm(d.y[i], state.mu + state.v[d.g[i]], state.s)
                elif kind == 1: lp += ld.bern(d.x[i], mcmc.where(state.m == 0, 0.5, state.theta))
                elif kind == 2: lp += ld.pois(d.k[i], M.exp(state.mu * 0.1 + state.v[d.g[i]] * 0.05) + 0.1)
                else: lp += ld.laplace(d.y[i], state.v[d.g[i]] * d.x[i], state.s + 0.5)
        if plan["derived"] >= 1: state.z1 = expr(2)
        if plan["derived"] >= 2: state.z2 = state.s * state.s + state.n
        return lp
    return params, log_post, data


@pytest.mark.parametrize("block", range(4))
def test_random_models_generated_code_equals_the_program(pkg, orc, tmp_path, block):
    compiled = declined = 0
    for seed in range(100 * block, 100 * block + 8):
        params, log_post, data = _build(pkg, seed)
        sub = tmp_path / f"m{seed}"
        sub.mkdir()
        try:
            hp = HostProgram(pkg, orc, sub, params, log_post, data, faithful=bool(seed & 1), _force_full=bool(seed & 2))
        except AssertionError as e:
            if "term cache" in str(e):
                declined += 1
                continue
            raise
        hp.check(np.random.default_rng(seed), trials=120)
        compiled += 1
    assert compiled >= 4, (compiled, declined)
