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
