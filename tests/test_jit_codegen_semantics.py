"""What the code generator prints, executed -- without a GPU.

tests/test_jit_codegen.py checks that the specialised full-program sweep of a model COMPILES for sm_100a; the GPU tests check that it
draws what the interpreter kernels draw. This file closes the gap on the CPU tier: the generated `jit_logpost()` / `jit_derived()`
(csrc/amwg_jit.cuh) is compiled for the host together with the product's own arithmetic headers and the plate helpers of the kernel
skeleton (csrc/amwg_jit_full_kernel.cuh, taken from the file's text) and evaluated at random states, moved components and proposal
values against tests/prog_eval.py -- the bytecode run with the oracle's arithmetic, which tests/test_tracer.py in turn holds against
the oracle's C models. Bit for bit: the generated code performs the program's operations in the program's order.

Constants that the library folds on the device at create (log(2*pi), lbeta(2,2) ...) are folded here by prog_eval with the oracle's
arithmetic and written into the model descriptor before the generator sees it, exactly what try_jit() reads back from the device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import models
import prog_eval

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bayes.js_b200", "csrc")

HOST_PRELUDE = r'''
#include "cuda_runtime.h"      // tests/host_shim
#include "math_constants.h"
#define __constant__
static unsigned char* g_smem_base = nullptr;                    // "shared memory": a host buffer, addressed by offset
namespace amwg {
static inline unsigned smem_u32(const void* p) { return (unsigned)((const unsigned char*)p - g_smem_base); }
static inline double lds_f64_sa(unsigned a) { double v; std::memcpy(&v, g_smem_base + a, 8); return v; }
}
#include "amwg_math.cuh"
#include "amwg_ld.cuh"
'''

HOST_EXPORTS = r'''
extern "C" {
void hs_set_smem(unsigned char* p) { g_smem_base = p; }
int hs_n_res() { return JN_RES; }
unsigned hs_res_off(int k) { return amwg::JRES_OFF[k]; }
int hs_res_col(int k) { return amwg::JRES_COL[k]; }
int hs_n_bern() { return JN_BERN; }
unsigned hs_bern_data(int k) { return amwg::JBERN_DATA[k]; }
int hs_bern_n(int k) { return amwg::JBERN_N[k]; }
unsigned hs_bern_mask(int k) { return amwg::JBERN_MASK[k]; }
double hs_logpost(const double* state, int moved, double val) { return amwg::jit_logpost(g_smem_base, state, 1ull, moved, val); }
void hs_derived(const double* state, double* der) {
#if JN_DERIVED > 0
  amwg::jit_derived(g_smem_base, state, 1ull, der);
#else
  (void)state; (void)der;
#endif
}
}
'''


def _skeleton_helpers():
    """the macros and plate helpers of the kernel skeleton: its text from `#define ST(c)` to the end of the first namespace block"""
    text = open(os.path.join(CSRC, "amwg_jit_full_kernel.cuh")).read()
    a = text.index("#define ST(c)")
    b = text.index("}  // namespace amwg", a)
    return "namespace amwg {\n" + text[a:b] + "}  // namespace amwg\n"


class HostProgram:
    def __init__(self, pkg, orc, tmp_path, params, log_post, data, **opts):
        o = {"chains": 4096, "_model_only": True}
        o.update(opts)
        self.s = s = pkg.mcmc.AmwgSampler(params, log_post, data, o)
        self.prog, self.O = s._program, orc.lib()
        self.consts = prog_eval.fold_constants(self.prog, self.O)
        m = s._model_keepalive[-1]
        for i, v in enumerate(self.consts):                           # what try_jit() reads back from the device after folding
            m.consts[i] = float(v)
        rc, msg, src = s.jit_compile_check()
        assert rc == 0, msg
        assert "#define JFULL 1" in src
        cpp = tmp_path / "gen.cpp"
        # the generated text = the #define prelude (up to the first namespace block), then tables and programs
        cut = src.index("namespace amwg {")
        cpp.write_text(src[:cut] + HOST_PRELUDE + _skeleton_helpers() + src[cut:] + HOST_EXPORTS)
        so = tmp_path / "gen.so"
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "tests", "host_shim"),
                            "-I" + CSRC, str(cpp), "-o", str(so)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-4000:]
        self.lib = L = C.CDLL(str(so))
        L.hs_logpost.restype, L.hs_logpost.argtypes = C.c_double, [C.POINTER(C.c_double), C.c_int, C.c_double]
        L.hs_derived.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
        for f in ("hs_res_off", "hs_bern_data", "hs_bern_mask"):
            getattr(L, f).restype = C.c_uint
        # stage the data the way the kernel does: resident columns by bulk copy, Bernoulli columns also as bit masks + "bad point" word
        self.smem = np.zeros(1 << 18, dtype=np.uint8)
        for k in range(L.hs_n_res()):
            col = np.asarray(self.prog.columns[L.hs_res_col(k)], dtype=np.float64)
            off = L.hs_res_off(k)
            self.smem[off:off + 8 * col.size] = col.view(np.uint8)
        for k in range(L.hs_n_bern()):
            n, d0, m0 = L.hs_bern_n(k), L.hs_bern_data(k), L.hs_bern_mask(k)
            y = self.smem[d0:d0 + 8 * n].view(np.float64)
            words = np.zeros((n + 31) // 32 + 1, dtype=np.uint32)
            for i in range(n):
                if y[i] == 1.0:
                    words[i >> 5] |= np.uint32(1 << (i & 31))
            words[-1] = 1 if np.any(~((y == 1.0) | (y == 0.0))) else 0
            self.smem[m0:m0 + 4 * words.size] = words.view(np.uint8)
        L.hs_set_smem(self.smem.ctypes.data_as(C.POINTER(C.c_ubyte)))
        self.D = self.prog.n_comp if hasattr(self.prog, "n_comp") else int(m.n_comp)
        self.types = []                                             # per component: (type, lower, upper)
        for p in range(m.n_params):
            pa = m.params[p]
            self.types += [(pa.type, pa.lower, pa.upper)] * pa.n_comp

    def random_state(self, rng):
        st = np.empty(self.D)
        for c, (t, lo, hi) in enumerate(self.types):
            st[c] = self.random_value(rng, t, lo, hi)
        return st

    @staticmethod
    def random_value(rng, t, lo, hi):
        if t == 2:                                                  # binary
            return float(rng.integers(0, 2))
        a = lo if np.isfinite(lo) else (hi - 20 if np.isfinite(hi) else -10.0)
        b = hi if np.isfinite(hi) else a + 20
        v = rng.uniform(a, b)
        if rng.random() < 0.05:                                     # now and then a value outside the support: -Infinity / NaN paths
            v = a - 1.5 if rng.random() < 0.5 else b + 1.5
        return float(np.floor(v + 0.5)) if t == 1 else float(v)

    def check(self, rng, trials=300):
        n_der = len(self.s._derived_names)
        for _ in range(trials):
            st = self.random_state(rng)
            moved = int(rng.integers(-1, self.D))
            val = self.random_value(rng, *self.types[moved]) if moved >= 0 else 0.0
            got = self.lib.hs_logpost(st.ctypes.data_as(C.POINTER(C.c_double)), moved, val)
            want = prog_eval.logpost(self.prog, self.consts, st, self.O, moved=moved, val=val)
            assert np.float64(got).view(np.uint64) == np.float64(want).view(np.uint64) or (got != got and want != want), (st, moved, val, got, want)
            if n_der:
                der = np.full(n_der, np.nan)
                self.lib.hs_derived(st.ctypes.data_as(C.POINTER(C.c_double)), der.ctypes.data_as(C.POINTER(C.c_double)))
                ref = [None] * n_der
                pc = self.prog.derived_prog
                if self.prog.variant_comps:
                    v = sum((1 << k) for k, c in enumerate(self.prog.variant_comps) if st[c] != 0)
                    pc = self.prog.variant_derived[v]
                prog_eval.run(self.prog, self.consts, st, pc, self.O, der=ref)
                assert [np.float64(x).view(np.uint64) for x in der] == [np.float64(x).view(np.uint64) for x in ref]


def _cases(pkg):
    ld, mcmc = pkg.ld, pkg.mcmc
    rng = np.random.default_rng(5)
    y = (rng.random(77) < 0.7).astype(float).tolist()
    nb = [int(v) for v in rng.integers(5, 30, 12)]

    def norm_derived(state, data):
        lp = 0
        lp += ld.norm(state.mu, 0, 100)
        lp += ld.unif(state.sigma, 0, 100)
        for i in range(len(data)):
            lp += ld.norm(data[i], state.mu, state.sigma)
        state.cv = state.sigma / state.mu
        state.var = state.sigma * state.sigma
        return lp
    return {
        "spike_where": (models.PARAMS_SPIKE, models.spike_bern(ld, mcmc), {"x": y}, {}),
        "spike_literal": (models.PARAMS_SPIKE, models.spike_bern_literal(ld), {"x": y}, {}),
        "spike_bad_point": (models.PARAMS_SPIKE, models.spike_bern(ld, mcmc), {"x": y[:40] + [2.0] + y[41:]}, {}),
        "complex_literal": (models.PARAMS_COMPLEX, models.complex_model_post_literal(ld), nb, {}),
        "complex_where": (models.PARAMS_COMPLEX, models.complex_model_post(ld, mcmc), nb, {}),
        "norm_faithful_derived": ({"mu": {"type": "real", "init": 180}, "sigma": {"type": "real", "lower": 0, "init": 5}}, norm_derived,
                                  rng.normal(184.5, 4.5, 64).tolist(), {"faithful": True}),
        "multi_bern": ({"x": {"type": "binary", "dim": [2, 2]}}, models.multi_bern_dens(mcmc), None, {}),
        "beta_bern_faithful": (models.PARAMS_THETA, models.beta_bern(ld), {"x": y}, {"faithful": True}),
    }


@pytest.mark.parametrize("name", ["spike_where", "spike_literal", "spike_bad_point", "complex_literal", "complex_where", "norm_faithful_derived",
                                  "multi_bern", "beta_bern_faithful"])
def test_generated_log_post_equals_the_program(pkg, orc, tmp_path, name):
    params, f, data, opts = _cases(pkg)[name]
    hp = HostProgram(pkg, orc, tmp_path, params, f, data, **opts)
    hp.check(np.random.default_rng(11))
