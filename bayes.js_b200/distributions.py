"""``ld`` -- the log-density surface of /root/reference/distributions.js (``ld.norm``, ``ld.pois`` ...).

Inside ``log_post`` the arguments are symbolic, and each call records one device opcode whose implementation
(csrc/amwg_ld.cuh) follows the JS source operation by operation.  Called with plain numbers (or numpy arrays)
the same device function is evaluated on the GPU through ``amwg_ld_eval`` -- there is no CPU implementation
in the product; ``oracle/`` holds the CPU restatement used by the tests.
"""
from __future__ import annotations

import numpy as np

from . import _ffi
from . import tracer as _tracer
from .tracer import JsThrow, Math, Sym, is_sym, lift, where

__all__ = ["lgamma", "lfactorial", "lchoose", "lbeta", "beta", "cauchy", "norm", "bivarnorm", "laplace", "dexp", "gamma",
           "invgamma", "lnorm", "pareto", "t", "weibull", "logis", "dirichlet", "exp", "unif", "bern", "cat", "binom",
           "nbinom", "hyper", "pois"]


def _device_eval(op: str, args):
    import os
    arrs = np.broadcast_arrays(*[np.asarray(a, dtype=np.float64) for a in args])
    shape = arrs[0].shape
    flat = np.ascontiguousarray(np.stack([a.reshape(-1) for a in arrs], axis=1))
    out = np.empty(flat.shape[0])
    dev = int(os.environ.get("AMWG_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    _ffi.check(_ffi.lib().amwg_ld_eval(_ffi.OP[op], flat.ctypes.data, flat.shape[1], flat.shape[0], out.ctypes.data, dev))
    return float(out[0]) if shape == () else out.reshape(shape)


def _op(op: str, arity: int, cite: str):
    def f(*args):
        if len(args) != arity:
            raise JsThrow(f"ld.{op[3:].lower() if op.startswith('LD_') else op.lower()} takes {arity} arguments")
        if any(is_sym(a) for a in args) or _tracer._ACTIVE:      # inside log_post: record (constants are folded on the device)
            return Sym(op, tuple(lift(a) for a in args))
        return _device_eval(op, args)
    f.__doc__ = f"distributions.js:{cite}"
    return f


lgamma = _op("LGAMMA", 1, "63-77")
lfactorial = _op("LFACTORIAL", 1, "79-82")
lchoose = _op("LCHOOSE", 2, "84-87")
lbeta = _op("LBETA", 2, "89-92")
beta = _op("LD_BETA", 3, "104-113")
cauchy = _op("LD_CAUCHY", 3, "115-117")
norm = _op("LD_NORM", 3, "119-121")
laplace = _op("LD_LAPLACE", 3, "136-138")
dexp = laplace                                   # distributions.js:140
gamma = _op("LD_GAMMA", 3, "142-152")
invgamma = _op("LD_INVGAMMA", 3, "154-159")
lnorm = _op("LD_LNORM", 3, "161-167")
pareto = _op("LD_PARETO", 3, "169-174")
t = _op("LD_T", 4, "176-180")
weibull = _op("LD_WEIBULL", 3, "185-191")
logis = _op("LD_LOGIS", 3, "196-201")
exp = _op("LD_EXP", 2, "217-219")
unif = _op("LD_UNIF", 3, "221-223")
bern = _op("LD_BERN", 2, "228-230")
binom = _op("LD_BINOM", 3, "240-248")
nbinom = _op("LD_NBINOM", 3, "267-272")
hyper = _op("LD_HYPER", 4, "274-280")
pois = _op("LD_POIS", 2, "282-284")


def _sym_or_num(x):
    return x if is_sym(x) else lift(x)


def _concrete(tree):
    """A composed ld.* called with plain numbers outside log_post: evaluate its expression on the device, one primitive at a time
    (same opcodes, same arithmetic as inside a program), and hand back a float like the other ld.* do."""
    if _tracer._ACTIVE:
        return tree
    stack = [tree]
    while stack:
        n = stack.pop()
        if n.op in ("COMP", "DATA", "DATA_I", "COMP_I"):
            return tree                                  # symbolic somewhere: stays a recorded expression
        stack.extend(n.args)

    def ev(n):
        if n.op == "CONST":
            return float(n.val)
        return _device_eval(n.op, [ev(a) for a in n.args])
    return ev(tree)


def bivarnorm(x, mean, sd, corr):
    """distributions.js:125-133 -- composed from device primitives in the JS operation order."""
    x0, x1, m0, m1, s0, s1, r = map(_sym_or_num, (x[0], x[1], mean[0], mean[1], sd[0], sd[1], corr))
    z = Math.pow(x0 - m0, 2) / Math.pow(s0, 2) + Math.pow(x1 - m1, 2) / Math.pow(s1, 2) - \
        (2 * r * (x0 - m0) * (x1 - m1)) / (s0 * s1)
    nf = -(Math.log(2) + Math.log(Math.PI) + Math.log(s0) + Math.log(s1) + 0.5 * Math.log(1 - Math.pow(r, 2)))
    return _concrete(nf - z / (2 * (1 - Math.pow(r, 2))))


def dirichlet(x, alpha):
    """distributions.js:203-214"""
    sum_alpha = lift(0.0)
    sum_lgamma_alpha = lift(0.0)
    s = lift(0.0)
    for i in range(len(alpha)):
        a = _sym_or_num(alpha[i])
        sum_alpha = sum_alpha + a
        sum_lgamma_alpha = sum_lgamma_alpha + Sym("LGAMMA", (a,))
        s = s + (a - 1) * Math.log(x[i])
    return _concrete(Sym("LGAMMA", (sum_alpha,)) - sum_lgamma_alpha + s)


def cat(x, probs):
    """distributions.js:232-238 -- probs[x - 1] with a symbolic x becomes a SELECT chain."""
    xs = _sym_or_num(x)
    n = len(probs)
    picked = lift(float("nan"))
    for k in range(n, 0, -1):
        picked = where(xs == k, Math.log(probs[k - 1]), picked)
    return _concrete(where(Sym("OR", (xs < 1, xs > n)), -float("inf"), picked))
