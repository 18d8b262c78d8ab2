#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE config 2.

    metric   posterior draws/sec (chains x iters), Normal(mu,sigma), N=1024 synthetic data
    step     one `sample(ITERS)` call for 2^20 chains per GPU (weak scaling), adaptation running, after a
             burn-in done in setup; `value` = kernel path with samples left in HBM (amwg_sample_device),
             `e2e` = the public API call mcmc.AmwgSampler.sample() returning host arrays (D2H inside).
    --impl reference   the CPU restatement of mcmc.js (oracle/, Node is absent) on all host cores.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

N_DATA = 1024
CHAINS_PER_GPU = 1 << 20
ITERS = 100                      # sweeps per step
BURN = 1000                      # setup (untimed): BASELINE throughput run burns 1000 first
FLOP_PER_DRAW = 6144.0           # 2 components x 1024 points x 3 flop
HBM_BYTES_PER_DRAW = 17.4        # 16 B sample write + ~1.4 B amortised state/log-SD/counters
FP64_NOMINAL_TFLOPS = 37.0       # B200 non-tensor fp64 (HGX B200 datasheet: 296 TF / 8 GPUs); no measured figure in MEASURED_PEAKS.json
NCU_DRAM_BYTES_PER_LAUNCH = 1.66e9   # ncu --set full (profiles/r01g), one amwg_sweep_kernel launch (50 sweeps, 2^20 chains): 1.589 GB written
                                     # (0.84 GB of samples + local-memory spill traffic at the 72-register cap) + 0.071 GB read
METRIC = "posterior draws/sec (chains x iters) Normal(mu,sigma) N=1024 at 1/2/4/8 B200"
PARAMS = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0}}


def config2_data():
    return np.random.default_rng(1024).normal(184.5, 4.5, N_DATA)


def make_log_post(ld):
    def log_post(state, data):                    # README.md:26-36
        log_post = 0
        log_post += ld.norm(state.mu, 0, 100)
        log_post += ld.unif(state.sigma, 0, 100)
        for i in range(len(data)):
            log_post += ld.norm(data[i], state.mu, state.sigma)
        return log_post
    return log_post


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(self.rows)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
def cpu_port_draws_per_sec(orc, seconds_target: float = 12.0):
    """Oracle (CPU restatement of mcmc.js, 2 evals per step like the reference) on ONE core, config-2 shape."""
    data = config2_data()
    t = orc.time_model("norm_readme", data, PARAMS, chains=1, burn=0, sample=2000)
    n = int(max(2000, min(200000, 2000 * seconds_target / max(t, 1e-6))))
    t = orc.time_model("norm_readme", data, PARAMS, chains=1, burn=0, sample=n)
    return n / t, f"1 chain x {n} draws, N={N_DATA}, single thread"


def cpu_baseline_time(model, data, params, chains, burn, sample):
    """cpu_baseline leg for the other BASELINE configs (scripts/bench_configs.py calls this; like every use of oracle/ outside tests/
    and smoke(), it lives in bench.py): seconds the CPU restatement takes for `chains` x (burn + sample) draws, one thread."""
    return graft.load_oracle().time_model(model, data, params, chains=chains, burn=burn, sample=sample)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path = the oracle port (Node is absent), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    orc = graft.load_oracle()
    data = config2_data()
    cores = os.cpu_count() or 1
    per_thread = 1500                                # draws per thread per step (bounded sample)

    def work(k, out):
        out[k] = orc.time_model("norm_readme", data, PARAMS, chains=1, burn=0, sample=per_thread, seed=k)

    def step():
        out = [0.0] * cores
        th = [threading.Thread(target=work, args=(k, out)) for k in range(cores)]
        [t.start() for t in th]
        [t.join() for t in th]

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = cores * per_thread * args.steps / dt
    sample = f"{cores} threads x 1 chain x {per_thread} draws per step, N={N_DATA}"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "draws/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: Normal(mu,sigma), N=1024 synthetic", "note": "CPU restatement of mcmc.js (Node unavailable)"},
            "cpu_baseline": {"value": value, "unit": "draws/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sampler has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = graft.load_package()
    mcmc, ld, ffi = pkg.mcmc, pkg.ld, pkg._ffi
    L = ffi.lib()

    chains_total = args.chains * world
    iters = args.iters
    sampler = mcmc.AmwgSampler(PARAMS, make_log_post(ld), config2_data().tolist(),
                               {"chains": chains_total, "seed": 0, "device": local_rank, "distributed": world > 1, "gather": "root"})
    local = sampler.local_chains
    sampler.burn(args.burn)
    mon = np.array([0, 1], dtype=np.int32)
    monp = mon.ctypes.data_as(C.POINTER(C.c_int32))
    dev_out = torch.empty((iters, 2, local), dtype=torch.float64, device=f"cuda:{local_rank}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        ffi.check(L.amwg_sample_device(sampler._handle, iters, 1, monp, 2, dev_out.data_ptr()))
        return sampler.last_sweep_kernel_ms()

    # ---- value: device-resident --------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    launches0 = sampler.kernel_launches()
    barrier()
    t0 = time.perf_counter()
    kernel_ms = 0.0
    for _ in range(args.steps):
        kernel_ms += step_device()
    barrier()
    dt = time.perf_counter() - t0
    launches = sampler.kernel_launches() - launches0
    clk = clocks.stop() if clocks else None

    # ---- e2e: public API, host arrays ---------------------------------------------------------------------------
    # N > 1: every process receives the draws of its own chains in its own host memory (options.gather = "none": one PCIe link
    # per GPU, the way a one-process-per-GPU job consumes them); the variant where rank 0 alone ends up with all chains
    # (gather = "root": NCCL gather first, then ONE host copy) is timed beside it as e2e_gather_root.
    e2e_steps = max(1, min(args.steps, 5 if world == 1 else 3))

    def e2e_leg(mode, steps):
        sampler.gather = mode
        w1 = sampler.sample(iters)             # warm-up: two live results = the two pinned buffers the loop alternates between
        w2 = sampler.sample(iters)
        del w1, w2
        barrier()
        t = time.perf_counter()
        for _ in range(steps):
            draws = sampler.sample(iters)
        barrier()
        dt_leg = time.perf_counter() - t
        want = chains_total if (mode == "root" and rank == 0) else local
        assert draws["mu"].shape == (iters, want)
        return dt_leg

    dt_e2e = e2e_leg("none", e2e_steps)
    root_steps = 2
    dt_root = e2e_leg("root", root_steps) if world > 1 else 0.0

    # ---- e2e with the summary formed on the device: same sweeps, only mean/sd/quantiles/R-hat leave the GPUs
    summ = sampler.sample_summary(iters)       # warm-up
    barrier()
    t2 = time.perf_counter()
    sum_ms = []
    for _ in range(e2e_steps):
        t3 = time.perf_counter()
        summ = sampler.sample_summary(iters)
        sum_ms.append(round(1e3 * (time.perf_counter() - t3), 2))
    barrier()
    dt_sum = time.perf_counter() - t2

    times = torch.tensor([dt, dt_e2e, kernel_ms, dt_sum, dt_root], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dt, dt_e2e, kernel_ms, dt_sum, dt_root = [float(v) for v in times.tolist()]

    if rank == 0:
        draws_per_step = chains_total * iters
        value = draws_per_step * args.steps / dt
        e2e = draws_per_step * e2e_steps / dt_e2e
        # roofline of the dominant kernel (amwg_sweep_kernel), per GPU, from the library's CUDA-event time on its launch stream
        kern_draws_per_s = (local * iters * args.steps) / (kernel_ms * 1e-3)
        hbm_peak, how = measured_peaks()
        ach_gbs = kern_draws_per_s * HBM_BYTES_PER_DRAW / 1e9
        ach_tf = kern_draws_per_s * FLOP_PER_DRAW / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": "draws/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config 2: Normal(mu,sigma), N=1024 synthetic, 2^20 chains per GPU", "chains_per_gpu": local,
                       "chains_total": chains_total, "iters_per_step": iters, "burn_in_setup": args.burn, "adapting": True,
                       "parallelism": f"chains sharded x{world}, no data-path collective",
                       "l2": "every step writes its samples (iters*2*chains*8 B = %.2f GB > 126 MB L2), which flushes L2" % (iters * 2 * local * 8 / 1e9)},
            "roofline": {"bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": NCU_DRAM_BYTES_PER_LAUNCH if (iters % 50 == 0 and local == CHAINS_PER_GPU) else None,
                         "traffic_source": "profiles/r01g: dram__bytes_read.sum + dram__bytes_write.sum of one 50-sweep launch over 2^20 chains",
                         "algorithmic_bytes_per_launch": HBM_BYTES_PER_DRAW * 50 * local,
                         "peak_source": how, "algorithmic_bytes_per_draw": HBM_BYTES_PER_DRAW,
                         "note": "north_star names HBM; with the N-point sum fused in-kernel the binding roof is the fp64 pipe, see roofline_fp64"},
            "roofline_fp64": {"bound": "fp64-issue", "achieved": ach_tf, "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                              "frac": ach_tf / FP64_NOMINAL_TFLOPS, "peak_source": "nominal (no measured fp64 peak)",
                              "algorithmic_flop_per_draw": FLOP_PER_DRAW, "kernel": "amwg_sweep_kernel",
                              "kernel_ms_per_step": kernel_ms / args.steps},
            "e2e": {"value": e2e, "unit": "draws/s", "h2d_bytes_per_step": int(mon.nbytes),
                    "d2h_bytes_per_step": int(iters * 2 * chains_total * 8), "steps": e2e_steps,
                    "note": "mcmc.AmwgSampler.sample(): pinned host buffer, D2H overlapped with the sweeps"
                            + ("; every rank copies the draws of its own chains to its own host memory (gather=\"none\")" if world > 1 else "")},
            "e2e_gather_root": ({"value": draws_per_step * root_steps / dt_root, "unit": "draws/s", "steps": root_steps,
                                 "note": "gather=\"root\": NCCL gather of the shards to rank 0, one host copy of all draws"} if world > 1 else None),
            "e2e_summary": {"value": draws_per_step * e2e_steps / dt_sum, "unit": "draws/s", "steps": e2e_steps,
                            "d2h_bytes_per_step": 2 * 4 * 8 + 8 * 2 * 10 * 256 * 8, "ms_per_call_rank0": sum_ms,
                            "median_call_value": draws_per_step / (1e-3 * sorted(sum_ms)[len(sum_ms) // 2]),
                            "mu": {"mean": summ["mu"]["mean"], "sd": summ["mu"]["sd"], "rhat": summ["mu"]["rhat"],
                                   "q2.5_50_97.5": [float(summ["mu"]["quantiles"][i]) for i in (0, 2, 4)]},
                            "note": "mcmc.AmwgSampler.sample_summary(): the draws stay in HBM; pooled mean/sd, exact quantiles (8-pass radix "
                                    "select) and R-hat over all chains x iterations come back"
                                    + ("; shards combined by an all-gather of moment records and an all-reduce of digit counts (NCCL)" if world > 1 else "")},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if world == 1 and not args.no_cpu:
            orc = graft.load_oracle()
            v, sample = cpu_port_draws_per_sec(orc)
            line["cpu_baseline"] = {"value": v, "unit": "draws/s", "cores": 1, "kind": "port", "sample": sample,
                                    "published_reference": "README.md:252: ~4e4 draws/s at N=1000 (author's machine, 2015)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS_PER_GPU, help="chains per GPU")
    ap.add_argument("--iters", type=int, default=ITERS, help="sweeps per step")
    ap.add_argument("--burn", type=int, default=BURN, help="burn-in sweeps done in setup (untimed)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
