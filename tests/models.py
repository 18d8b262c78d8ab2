"""The reference's models (README.md, tests/test_data.js) written against the `mcmc` / `ld` mirror, and the names of
their C twins in the oracle."""


def norm_post_readme(ld):
    def log_post(state, data):                      # README.md:26-36
        log_post = 0
        log_post += ld.norm(state.mu, 0, 100)
        log_post += ld.unif(state.sigma, 0, 100)
        for i in range(len(data)):
            log_post += ld.norm(data[i], state.mu, state.sigma)
        return log_post
    return log_post


def norm_post_test(ld):
    def norm_post(par, data):                       # tests/test_data.js:80-91
        mu = par.mu
        sigma = par.sigma
        log_post = 0
        log_post += ld.norm(mu, 0, 100)
        log_post += ld.unif(sigma, 0, 100)
        for i in range(len(data)):
            log_post += ld.norm(data[i], mu, sigma)
        par.var = sigma * sigma
        return log_post
    return norm_post


def beta_bern(ld):
    def log_post(state, data):                      # README.md:149-164
        log_post = 0
        log_post += ld.beta(state.theta, 2, 2)
        n = len(data.x)
        for i in range(n):
            log_post += ld.bern(data.x[i], state.theta)
        return log_post
    return log_post


def spike_bern(ld, mcmc):
    def log_post(state, data):                      # BASELINE config 3 (+ binary indicator, pattern of test_data.js:154-171)
        theta, m = state.theta, state.m
        log_post = 0
        log_post += ld.beta(theta, 2, 2)
        log_post += ld.bern(m, 0.5)
        for i in range(len(data.x)):
            log_post += ld.bern(data.x[i], mcmc.where(m == 0, 0.5, theta))
        return log_post
    return log_post


PARAMS_NORM = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}                      # README.md:22-24
PARAMS1 = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0, "init": 1}}             # tests/test_data.js:9-18
PARAMS_THETA = {"theta": {"type": "real", "lower": 0, "upper": 1}}                               # README.md:201
PARAMS_SPIKE = {"theta": {"type": "real", "lower": 0, "upper": 1}, "m": {"type": "binary"}}


# ---- the remaining fixtures of tests/test_data.js -------------------------------------------------------------------
def norm_dens(ld):
    return lambda par, data=None: ld.norm(par.x, 10, 5)                      # tests/test_data.js:93-95


def poisson_dens(ld):
    return lambda par, data=None: ld.pois(par.x, 10)                         # :109-111


def bern_dens(ld):
    return lambda par, data=None: ld.bern(par.x, 0.85)                       # :125-127


def multivar_norm_dens(ld):
    def f(par, data=None):                                                   # :97-107
        x1, x2, x3, x4 = par.x[0][0], par.x[0][1], par.x[1][0], par.x[1][1]
        return ld.norm(x1, 1000, 50) + ld.norm(x2, 10, 5) + ld.norm(x3, 0.1, 0.5) + ld.norm(x4, 0.001, 0.05)
    return f


def multivar_poisson_dens(ld):
    def f(par, data=None):                                                   # :113-123
        x1, x2, x3, x4 = par.x[0][0], par.x[0][1], par.x[1][0], par.x[1][1]
        return ld.pois(x1, 0.1) + ld.pois(x2, 10) + ld.pois(x3, 1000) + ld.pois(x4, 100000)
    return f


def multi_bern_dens(mcmc):
    def f(par, data=None):                                                   # :129-136
        x1, x2, x3, x4 = par.x[0][0], par.x[0][1], par.x[1][0], par.x[1][1]
        return mcmc.Math.log(x1 * x2 * 0.85 + (1 - x1 * x2) * 0.15) + mcmc.Math.log(x3 * x4 * 0.75 + (1 - x3 * x4) * 0.25)
    return f


def complex_model_post(ld, mcmc):
    def f(par, x):                                                           # :154-171 (`if (m === 0)` written as where)
        p1, n1, m = par.p1, par.n1, par.m
        log_post = 0
        log_post += ld.bern(m, 0.4)
        log_post += ld.beta(p1, 2, 2)
        log_post += ld.nbinom(n1, 2, 0.1)
        for i in range(len(x)):
            log_post += mcmc.where(m == 0, ld.nbinom(x[i], 21, 0.5), ld.nbinom(x[i], n1, p1))
        return log_post
    return f


def hierarchical_binomial_post(ld, mcmc):
    def logit(p):
        return mcmc.Math.log(p / (1 - p))                                    # :195-197

    def f(par, d):                                                           # :199-211
        p = par.p[0]
        mu_logit_p, sigma_logit_p = par.mu_logit_p, par.sigma_logit_p
        log_post = 0
        log_post += ld.norm(mu_logit_p, 0, 10)
        log_post += ld.norm(sigma_logit_p, 0, 10)
        for i in range(len(d.x)):
            log_post += ld.norm(logit(p[i]), mu_logit_p, sigma_logit_p)
            log_post += ld.binom(d.x[i], d.n[i], p[i])
        return log_post
    return f


PARAMS_COMPLEX = {"p1": {"type": "real", "lower": 0, "upper": 1}, "n1": {"type": "int", "lower": 1, "init": 1}, "m": {"type": "binary"}}   # :138-152
BINOM_DATA = {"x": [5, 6, 9, 14, 13, 20], "n": [10, 10, 20, 20, 30, 30]}                                                                # :174
PARAMS_HIER_BINOM = {"p": {"type": "real", "init": 0.5, "lower": 0, "upper": 1, "dim": [1, 6]},
                     "mu_logit_p": {"type": "real", "init": 0}, "sigma_logit_p": {"type": "real", "lower": 0, "init": 1}}                # :176-193


# ---- the same models with the reference's literal control flow on the binary parameter (recorded once per configuration) ----
def spike_bern_literal(ld):
    def log_post(state, data):
        theta, m = state.theta, state.m
        log_post = 0
        log_post += ld.beta(theta, 2, 2)
        log_post += ld.bern(m, 0.5)
        for i in range(len(data.x)):
            if m == 0:
                log_post += ld.bern(data.x[i], 0.5)
            else:
                log_post += ld.bern(data.x[i], theta)
        return log_post
    return log_post


def complex_model_post_literal(ld):
    def f(par, x):                                                           # tests/test_data.js:154-171, as written
        p1, n1, m = par.p1, par.n1, par.m
        log_post = 0
        log_post += ld.bern(m, 0.4)
        log_post += ld.beta(p1, 2, 2)
        log_post += ld.nbinom(n1, 2, 0.1)
        for i in range(len(x)):
            if m == 0:
                log_post += ld.nbinom(x[i], 21, 0.5)
            else:
                log_post += ld.nbinom(x[i], n1, p1)
        return log_post
    return f


def hier_norm_post(ld):
    def f(state, d):                                # BASELINE config 4, any number of groups
        lp = 0
        for j in range(len(state.mu)):
            lp += ld.norm(state.mu[j], 0, 100)
        lp += ld.unif(state.sigma, 0, 100)
        for i in range(len(d.y)):
            lp += ld.norm(d.y[i], state.mu[d.g[i]], state.sigma)
        return lp
    return f


def pois_reg_post(ld, mcmc):
    def f(state, d):                                # BASELINE config 5, any number of coefficients
        lp = 0
        for k in range(len(state.beta)):
            lp += ld.norm(state.beta[k], 0, 10)
        for i in range(len(d.y)):
            eta = 0
            for k in range(len(state.beta)):
                eta += d.X[i][k] * state.beta[k]
            lp += ld.pois(d.y[i], mcmc.Math.exp(eta))
        return lp
    return f
