#!/usr/bin/env python
"""Throughput of the five BASELINE.json configs (BASELINE.md section 3 "table to fill"). Configs 2 is bench.py's line; the
others are parity-test cases whose rates are recorded here for the report (profiles/r01_configs.json). One JSON line each.
Sizes for configs 4 and 5 are per GPU (the BASELINE sizes are 2^18 chains over 4 GPUs and 2^22 over 8)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
import bench  # noqa: E402  (the cpu_baseline legs go through bench.py, the one place outside tests/ that executes oracle/)
mcmc, ld = pkg.mcmc, pkg.ld
PRESIDENTS = [183, 192, 182, 183, 177, 185, 188, 188, 182, 185]


def norm_post(state, data):
    lp = 0
    lp += ld.norm(state.mu, 0, 100)
    lp += ld.unif(state.sigma, 0, 100)
    for i in range(len(data)):
        lp += ld.norm(data[i], state.mu, state.sigma)
    return lp


def emit(**kw):
    print(json.dumps(kw), flush=True)


def gpu_rate(sampler, burn, iters, reps=3):
    sampler.burn(burn)
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        sampler.burn(iters)
        dt = time.perf_counter() - t0
        best = max(best, sampler.n_chains * iters / dt)
    return best, sampler.last_sweep_kernel_ms()


P_NORM = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}

WANT = set(os.environ.get("CONFIGS", "1,3,4,5").split(","))   # CONFIGS=4 runs one section


# config 1: the reference's own CPU-runnable case (README.md:39-42): 1 chain, burn 1000, sample 5000
def config1():
    t = bench.cpu_baseline_time("norm_readme", PRESIDENTS, P_NORM, 20, 1000, 5000)
    emit(config=1, what="CPU restatement of mcmc.js, presidents data N=10, 1 chain x (1000 burn + 5000 draws), single thread",
         draws_per_s=20 * 6000 / t, note="Node unavailable; README.md:252 reports ~4e4 draws/s at N=1000 on the author's machine")
    s = mcmc.AmwgSampler(P_NORM, norm_post, PRESIDENTS, {"chains": 1 << 20, "seed": 0})
    r, ms = gpu_rate(s, 200, 200)
    emit(config=1, what="GPU, presidents data N=10, 2^20 chains", draws_per_s=r)
    del s


if "1" in WANT:
    config1()

# config 3: Beta-Bernoulli N=256 + binary indicator (exercises the binary stepper), 2^20 chains
def config3():
    y = (np.random.default_rng(256).random(256) < 0.7).astype(float)


    def spike(state, d):
        lp = 0
        lp += ld.beta(state.theta, 2, 2)
        lp += ld.bern(state.m, 0.5)
        for i in range(len(d.x)):
            lp += ld.bern(d.x[i], mcmc.where(state.m == 0, 0.5, state.theta))
        return lp


    P3 = {"theta": {"type": "real", "lower": 0, "upper": 1}, "m": {"type": "binary"}}
    s = mcmc.AmwgSampler(P3, spike, {"x": y.tolist()}, {"chains": 1 << 20, "seed": 0})
    r, ms = gpu_rate(s, 200, 100)
    emit(config=3, what="GPU, Beta-Bernoulli N=256 + binary indicator, 2^20 chains (sequential bit-faithful Bernoulli plate)", draws_per_s=r,
         program=s.program_summary())
    del s
    t = bench.cpu_baseline_time("spike_bern", {"x": y}, P3, 1, 0, 20000)
    emit(config=3, what="CPU restatement, 1 chain x 20000 draws, single thread", draws_per_s=20000 / t)


if "3" in WANT:
    config3()

# config 4: hierarchical Normal, 64 groups x 1024, D = 65; per GPU 2^16 chains in BASELINE -> 2^14 here (rate is per chain-iteration)
def config4():
    J, per = 64, 1024
    g = np.repeat(np.arange(J), per)
    mu_true = np.random.default_rng(64).normal(100, 20, J)
    yy = mu_true[g] + np.random.default_rng(65).normal(0, 5, J * per)


    def hier(state, d):
        lp = 0
        for j in range(J):
            lp += ld.norm(state.mu[j], 0, 100)
        lp += ld.unif(state.sigma, 0, 100)
        for i in range(len(d.y)):
            lp += ld.norm(d.y[i], state.mu[d.g[i]], state.sigma)
        return lp


    P4 = {"mu": {"type": "real", "dim": [J]}, "sigma": {"type": "real", "lower": 0}}
    t0 = time.perf_counter()
    s = mcmc.AmwgSampler(P4, hier, {"y": yy.tolist(), "g": g.tolist()}, {"chains": 1 << 16, "seed": 0})
    t_trace = time.perf_counter() - t0
    r, ms = gpu_rate(s, 2, 4, reps=2)
    emit(config=4, what="GPU, hierarchical Normal N=65536, D=65, 2^16 chains on one GPU (the BASELINE per-GPU share); pre-evaluated statistics: "
         "all 65 proposals drawn first, ONE pass over y (512 KB through the TMA tile ring) gives every group's sum of squares at its proposal, "
         "then 65 O(1) steps from cached terms: 65536 point-terms per sweep instead of 65 x 65536", draws_per_s=r,
         trace_seconds=t_trace, n_plates=len(s._program.plates), program=s.program_summary()[-1])
    del s
    t = bench.cpu_baseline_time("hier_norm", {"y": yy, "g": g}, P4, 1, 0, 3)
    emit(config=4, what="CPU restatement, 1 chain x 3 draws, single thread", draws_per_s=3 / t)


if "4" in WANT:
    config4()

# config 5: Poisson regression, 8 coefficients, N = 1e6; 2^12 chains on one GPU
def config5():
    K, n = 8, 1000000
    X = np.column_stack([np.ones(n), np.random.default_rng(8).normal(0, 0.5, (n, K - 1))])
    beta_true = np.random.default_rng(9).normal(0, 0.3, K)
    yc = np.random.default_rng(10).poisson(np.exp(X @ beta_true)).astype(float)


    def poisreg(state, d):
        lp = 0
        for k in range(K):
            lp += ld.norm(state.beta[k], 0, 10)
        for i in mcmc.points(len(d.y)):
            eta = 0
            for k in range(K):
                eta += d.X[i][k] * state.beta[k]
            lp += ld.pois(d.y[i], mcmc.Math.exp(eta))
        return lp


    P5 = {"beta": {"type": "real", "dim": [K]}}
    s = mcmc.AmwgSampler(P5, poisreg, {"y": yc, "X": X}, {"chains": 1 << 12, "seed": 0})
    r, ms = gpu_rate(s, 1, 1, reps=2)
    emit(config=5, what="GPU, Poisson regression N=1e6, K=8, 2^12 chains on one GPU (X 64 MB streamed from L2)", draws_per_s=r, program=s.program_summary()[-1])
    del s
    t = bench.cpu_baseline_time("pois_reg", {"y": yc, "X": X}, P5, 1, 0, 1)
    emit(config=5, what="CPU restatement, 1 chain x 1 draw, single thread", draws_per_s=1 / t)


if "5" in WANT:
    config5()
