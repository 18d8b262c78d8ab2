"""Host logic that mirrors mcmc.js, against the reference's own golden fixtures (tests/test_data.js) -- no GPU."""
import copy
import os
import ctypes
import re

import numpy as np
import pytest

INF = float("inf")

# tests/test_data.js:9-35 and :37-74 -- the only deterministic fixtures the reference's tests hold (tests/test_mcmc_js.R:39-46)
PARAMS1 = {"mu": {"type": "real"}, "sigma": {"type": "real", "lower": 0, "init": 1}}
PARAMS1_COMPLETED = {"mu": {"type": "real", "dim": [1], "upper": INF, "lower": -INF, "init": 0.5},
                     "sigma": {"type": "real", "dim": [1], "upper": INF, "lower": 0, "init": 1}}
PARAMS2 = {"theta": {"init": lambda: 1.5}, "state": {"type": "binary", "init": 1},
           "mat": {"type": "int", "dim": [3, 3], "init": lambda: 2}}
PARAMS2_COMPLETED = {"theta": {"type": "real", "dim": [1], "upper": INF, "lower": -INF, "init": 1.5},
                     "state": {"type": "binary", "init": 1, "dim": [1], "upper": 1, "lower": 0},
                     "mat": {"type": "int", "dim": [3, 3], "upper": INF, "lower": -INF, "init": [[2, 2, 2], [2, 2, 2], [2, 2, 2]]}}


def test_complete_params_golden_fixtures(pkg):
    mcmc = pkg.mcmc
    src1, src2 = copy.deepcopy(PARAMS1), dict(PARAMS2)
    assert mcmc.complete_params(PARAMS1, mcmc.param_init_fixed) == PARAMS1_COMPLETED
    assert mcmc.complete_params(PARAMS2, mcmc.param_init_fixed) == PARAMS2_COMPLETED
    assert PARAMS1 == src1 and set(PARAMS2) == set(src2)            # the input is not modified (mcmc.js:358)


def test_param_init_fixed_table(pkg, orc):
    """mcmc.js:313-341, every branch; the oracle's independent restatement agrees."""
    f = pkg.mcmc.param_init_fixed
    cases = [("real", -INF, INF, 0.5), ("real", -INF, 3, 2.5), ("real", 2, INF, 2.5), ("real", 0, 1, 0.5),
             ("int", -INF, INF, 1), ("int", -INF, 3, 2), ("int", 2, INF, 3), ("int", 0, 5, 3), ("int", -3, -2, -2),
             ("binary", 0, 1, 1)]
    for t, lo, hi, want in cases:
        assert f(t, lo, hi) == want
        assert orc.param_init_fixed(t, lo, hi) == want
    with pytest.raises(pkg.JsThrow, match="Can not initialize parameter where lower bound > upper bound"):
        f("real", 2, 1)
    with pytest.raises(pkg.JsThrow, match=re.escape("Could not initialize parameter of type complex[-Infinity, Infinity]")):
        f("complex", -INF, INF)


def test_array_helpers(pkg):
    m = pkg.mcmc
    assert m.create_array([2, 3], 1) == [[1, 1, 1], [1, 1, 1]]                      # mcmc.js:143-145
    assert m.array_dim(m.create_array([4, 2, 1], 0)) == [4, 2, 1]                    # mcmc.js:174-176
    assert m.array_equal([1, [2, 3]], [1, [2, 3]]) and not m.array_equal([1, 2], [1, 3])
    with pytest.raises(pkg.JsThrow, match="create_array can't create a dimensionless array"):
        m.create_array([], 0)
    assert m.get_option("pi", {"pi": 3.14159}, 3.14) == 3.14159 and m.get_option("pi", {"pi": None}, 3.14) == 3.14
    assert m.get_option("x", {"x": 0}, 5) == 0 and m.get_option("x", None, 5) == 5       # 0 is kept (mcmc.js:282-284)
    assert m.get_multidim_option("b", {"b": 10}, [2, 2], 50) == [[10, 10], [10, 10]]
    with pytest.raises(pkg.JsThrow, match=re.escape("The option b is of dimension [3] but should be [2,2].")):
        m.get_multidim_option("b", {"b": [1, 2, 3]}, [2, 2], 50)
    assert [m.js_round(v) for v in (-2.5, 2.5, 0.49999999999999994, -0.5, 1.2)] == [-2.0, 3.0, 0.0, 0.0, 1.0]


def test_option_merge_reproduces_the_or_quirk(pkg):
    """AmwgStepper ctor, mcmc.js:871-878: `a || b` lets falsy values fall through and mutates options.params[name]."""
    m = pkg.mcmc
    params = m.complete_params({"mu": {"type": "real"}, "p": {"type": "real", "dim": [2]}, "z": {"type": "binary"}})
    opts = {"max_adaptation": 0.5, "batch_size": 10, "is_adapting": False, "prop_log_scale": 0,
            "params": {"mu": {"max_adaptation": 0.1, "prop_log_scale": 0}, "p": {"target_accept_rate": [0.3, 0.2]}}}
    r = m.resolve_stepper_options(params, opts)
    assert r["mu"]["max_adaptation"] == [0.1] and r["p"]["max_adaptation"] == [0.5, 0.5]     # README.md:192 example
    assert r["mu"]["batch_size"] == [10] and r["p"]["batch_size"] == [10, 10]
    assert r["p"]["target_accept_rate"] == [0.3, 0.2] and r["mu"]["target_accept_rate"] == [0.44]
    assert r["mu"]["is_adapting"] == [False]         # undefined || false === false, and get_option keeps false (mcmc.js:282-284)
    r2 = m.resolve_stepper_options(params, {"is_adapting": True, "prop_log_scale": 2, "params": {"mu": {"is_adapting": False, "prop_log_scale": 0}}})
    assert r2["mu"]["is_adapting"] == [True] and r2["mu"]["prop_log_scale"] == [2]      # falsy per-parameter values fall through to the global ones
    assert r["mu"]["prop_log_scale"] == [0]
    assert r["z"] == {}                               # BinaryStepper takes no options
    assert opts["params"]["mu"]["batch_size"] == 10   # mutated in place, like the reference
    with pytest.raises(pkg.JsThrow, match="AmwgStepper can't handle parameter q with type complex"):
        m.resolve_stepper_options({"q": {"type": "complex", "dim": [1]}}, None)


def test_abi_header_and_binding_agree(pkg):
    """Opcode numbering, plate kinds and every exported symbol of include/amwg.h match the ctypes binding and the built library."""
    import os
    ffi = pkg._ffi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "amwg.h")).read()
    ops = re.findall(r"AMWG_OP_([A-Z0-9_]+)\b", hdr.split("enum {\n  AMWG_OP_END = 0,")[1].split("AMWG_OP__COUNT")[0])
    names = ["END"] + [o for o in ops]
    seen = []
    for n in names:
        if n not in seen:
            seen.append(n)
    assert seen == ffi._OPS
    kinds = re.findall(r"AMWG_PLATE_([A-Z_]+)", hdr.split("AMWG_PLATE_GENERIC = 0")[1].split("};")[0])
    assert ["GENERIC"] + kinds[:4] == ["GENERIC", "NORM_IID", "BERN_IID", "NORM_GROUPED", "POIS_LOGLIN"]
    declared = set(re.findall(r"AMWG_API [a-z0-9_ \*]+?(amwg_[a-z0-9_]+)\(", hdr))
    assert declared == set(ffi.EXPORTS)
    lib = ctypes.CDLL(ffi.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.amwg_abi_version() == int(re.search(r"#define AMWG_ABI_VERSION (\d+)", hdr).group(1)) == ffi.ABI_VERSION
    assert ctypes.sizeof(ffi.AmwgParam) == 32 and ctypes.sizeof(ffi.AmwgCompOptions) == 48 and ctypes.sizeof(ffi.AmwgPlate) == 40


def test_no_gpu_means_a_loud_failure_not_a_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import models
    with pytest.raises(pkg.JsThrow):
        pkg.mcmc.AmwgSampler(models.PARAMS_NORM, models.norm_post_readme(pkg.ld), [1.0, 2.0, 3.0])
    with pytest.raises(pkg._ffi.AmwgError):
        pkg.ld.norm(183, 180, 5)


def test_product_never_touches_the_oracle():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "bayes.js_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src and "oracle/" not in src.replace("oracle/`` holds", "").replace("``oracle/``", ""), f
    # the measurement scripts reach the CPU restatement only through bench.py's cpu_baseline helpers
    for f in os.listdir(os.path.join(root, "scripts")):
        if f.endswith(".py"):
            src = open(os.path.join(root, "scripts", f)).read()
            assert "load_oracle" not in src and "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_bench_reference_arm_runs_on_cpu_and_the_product_arm_refuses_without_a_gpu():
    """bench.py --impl reference times the CPU restatement (the one place outside tests/ and smoke() that may execute oracle/);
    the product arm has no CPU fallback: without a CUDA device it exits with an error instead of printing a number."""
    import json
    import subprocess
    import sys
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "draws/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["gpu_launches"] == 0
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"],
                           capture_output=True, text=True, timeout=600, cwd=root)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
