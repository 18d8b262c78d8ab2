"""CPU evaluator of the lowered log_post program (include/amwg.h), for the `-m "not gpu"` tests of the tracer.
Test infrastructure: it executes the bytecode with the ORACLE's Math.log/exp and ld.* (oracle/liboracle.so), so a
program that lowers a model correctly reproduces the oracle's C model bit for bit (sequential plates) or to rounding
(factorised plates). It mirrors run_program() in csrc/amwg_kernels.cu."""
import math

import numpy as np

from bayes_js_b200._ffi import OP, PLATE_BERN_IID, PLATE_NORM_GROUPED, PLATE_NORM_IID, PLATE_POIS_LOGLIN

INV = {v: k for k, v in OP.items()}
JS_PI = 3.141592653589793


def _ld(O, name, *a):
    return getattr(O, "orc_ld_" + name)(*[float(v) for v in a])


def fold_constants(prog, O):
    consts = list(prog.consts)
    for pc, dst in zip(prog.fold_prog, prog.fold_dst):
        consts[dst] = run(prog, consts, None, pc, O, want_top=True)
    return consts


def run(prog, consts, state, pc, O, want_top=False, der=None, moved=-1, val=0.0, cache=None, cand=None, props=None):
    """`cache` / `cand`: the chain's term cache (list of n_terms values) and where STORE writes (cand, or the cache itself when
    cand is None) -- mirrors EvalState.tval / tcand / direct. `props`: every component at its proposal (stat_prog's evaluation)."""
    code, cols, plates = prog.code, prog.columns, prog.plates
    stk = []
    lp = 0.0
    li = ln = 0

    def comp(c):
        if props is not None:
            return float(props[c])
        return val if c == moved else float(state[c])

    def nxt():
        nonlocal pc
        v = code[pc]; pc += 1
        return v

    def opnd(mode):
        if mode == 0:
            return stk.pop()
        ix = nxt()
        return consts[ix] if mode == 1 else comp(ix)

    def div(x, y):
        if y != 0: return x / y
        if x == 0 or x != x: return math.nan
        return math.copysign(math.inf, x) * math.copysign(1.0, y)

    BIN = {"ADD": lambda x, y: x + y, "SUB": lambda x, y: x - y, "MUL": lambda x, y: x * y, "DIV": div,
           "POW": lambda x, y: x * x if y == 2.0 else math.pow(x, y),
           "LT": lambda x, y: float(x < y), "LE": lambda x, y: float(x <= y), "GT": lambda x, y: float(x > y),
           "GE": lambda x, y: float(x >= y), "EQ": lambda x, y: float(x == y), "NE": lambda x, y: float(x != y),
           "AND": lambda x, y: float(x != 0 and y != 0), "OR": lambda x, y: float(x != 0 or y != 0),
           "LCHOOSE": lambda x, y: _ld(O, "lchoose", x, y), "LBETA": lambda x, y: _ld(O, "lbeta", x, y)}
    UN = {"NEG": lambda x: -x, "LOG": lambda x: O.orc_log(x), "EXP": lambda x: O.orc_exp(x),
          "SQRT": lambda x: math.sqrt(x) if x >= 0 else math.nan, "ABS": abs, "NOT": lambda x: float(x == 0),
          "LGAMMA": lambda x: _ld(O, "lgamma", x), "LFACTORIAL": lambda x: _ld(O, "lfactorial", x)}
    while True:
        w = nxt() & 0xffffffff
        op = INV[w & 0xff]
        mA, mB, mC, mD = (w >> 8) & 3, (w >> 10) & 3, (w >> 12) & 3, (w >> 14) & 3
        acc, store, a = (w >> 16) & 1, (w >> 17) & 1, w >> 18
        r = None
        plate_v = None
        if op == "END":
            return stk[-1] if (want_top and stk) else lp
        if op == "CONST": r = consts[a]
        elif op == "COMP": r = comp(a)
        elif op == "DATA": r = float(cols[a][nxt()])
        elif op == "DATA_I":
            off = nxt(); stride = nxt(); r = float(cols[a][off + stride * li])
        elif op == "COMP_I":
            off = nxt(); stride = nxt(); base = nxt()
            r = comp(base + int(cols[a][off + stride * li]))
        elif op in BIN:
            y = opnd(mB); x = opnd(mA); r = BIN[op](x, y)
        elif op in UN:
            r = UN[op](opnd(mA))
        elif op == "SELECT":
            z = opnd(mC); y = opnd(mB); x = opnd(mA); r = y if x != 0 else z
        elif op in ("NORM_K", "UNIF_K", "BETA_K"):
            t = opnd(mD); z = opnd(mC); y = opnd(mB); x = opnd(mA)
            if op == "NORM_K":
                d = x - y; r = z - div(d * d, t)
            elif op == "UNIF_K":
                r = -math.inf if (x < y or x > z) else t
            else:
                r = -math.inf if (x > 1 or x < 0) else (y * O.orc_log(x) + z * O.orc_log(1 - x)) - t
        elif op.startswith("LD_"):
            n = {"LD_BERN": 2, "LD_POIS": 2, "LD_EXP": 2, "LD_T": 4, "LD_HYPER": 4}.get(op, 3)
            args = [opnd(m) for m in (mA, mB, mC, mD)[:n][::-1]][::-1]
            r = _ld(O, op[3:].lower(), *args)
        elif op == "PLATE_SS":
            mean = opnd(mA); slot = nxt()
            pl = plates[a]
            x = np.asarray(cols[pl["col"][0]][pl["iparam"][2]:pl["iparam"][2] + pl["n"]], dtype=np.float64)
            r = float(np.sum((x - mean) ** 2))
            if cache is not None:
                (cand if cand is not None else cache)[slot] = r
        elif op == "NORM_SS":
            sd = opnd(mB); S = opnd(mA)
            r = plates[a]["n"] * (-0.5 * O.orc_log(2 * JS_PI) - O.orc_log(sd)) - S / (2 * sd * sd)
        elif op == "CACHED": r = cache[a]
        elif op == "CAND": r = cand[a]
        elif op == "ACC": lp = lp + stk.pop()
        elif op == "ACC_RANGE":
            cnt = nxt()
            for k in range(cnt):
                lp = lp + cache[a + k]
        elif op == "STORE": der[a] = stk.pop()
        elif op == "LOOP_BEGIN":
            skip = nxt()
            li, ln = 0, plates[a]["n"]
            if ln <= 0: pc = skip
        elif op == "LOOP_END":
            body = nxt()
            lp = lp + stk.pop(); li += 1
            if li < ln: pc = body
            else: li = 0
        elif op == "PLATE":
            pl = plates[a]
            n, off = pl["n"], pl["iparam"][2]
            x = np.asarray(cols[pl["col"][0]][off:off + n], dtype=np.float64)
            c0 = -0.5 * O.orc_log(2 * JS_PI)
            if pl["kind"] == PLATE_NORM_IID:
                sd = opnd(mB); mean = opnd(mA)
                S = float(np.sum((x - mean) ** 2))
                plate_v = n * (c0 - O.orc_log(sd)) - S / (2 * sd * sd)
                lp = lp + plate_v
            elif pl["kind"] == PLATE_BERN_IID:
                p = opnd(mA)
                l1 = O.orc_log(1.0 * p + (1 - 1.0) * (1 - p)); l0 = O.orc_log(0.0 * p + (1 - 0.0) * (1 - p))
                for yi in x:
                    lp = lp + (l1 if yi == 1.0 else (l0 if yi == 0.0 else -math.inf))
            elif pl["kind"] == PLATE_NORM_GROUPED:
                sd = opnd(mA)
                start = cols[pl["col"][1]].astype(int); base, J = pl["iparam"][0], pl["iparam"][1]
                S = 0.0
                for j in range(J):
                    S += float(np.sum((x[start[j]:start[j + 1]] - comp(base + j)) ** 2))
                plate_v = n * (c0 - O.orc_log(sd)) - S / (2 * sd * sd)
                lp = lp + plate_v
            elif pl["kind"] == PLATE_POIS_LOGLIN:
                base, K = pl["iparam"][0], pl["iparam"][1]
                X = np.asarray(cols[pl["col"][1]]).reshape(-1, K)[:n]
                stats = np.asarray(cols[pl["col"][2]])           # [X^T y (K) | sum lfactorial(y)], amwg.h
                beta = np.array([comp(base + k) for k in range(K)])
                eta = X @ beta
                plate_v = float(stats[:K] @ beta - np.sum(np.exp(eta)) - stats[K])
                lp = lp + plate_v
            else:
                raise AssertionError("generic plates are LOOP_BEGIN/LOOP_END")
            if store:
                t_id = nxt()
                if cache is not None:
                    (cand if cand is not None else cache)[t_id] = plate_v if plate_v is not None else 0.0
        else:
            raise AssertionError(op)
        if r is not None:
            if acc:
                lp = lp + r
                if store:
                    t_id = nxt()
                    if cache is not None:
                        (cand if cand is not None else cache)[t_id] = r
            else: stk.append(r)


def logpost(prog, consts, state, O, moved=-1, val=0.0):
    pc = prog.logpost_prog
    if prog.variant_comps:                      # program selected by the configuration of the binary components
        v = 0
        for k, c in enumerate(prog.variant_comps):
            x = val if c == moved else float(state[c])
            v |= (1 << k) if x != 0 else 0
        pc = prog.variant_logpost[v]
    return run(prog, consts, state, pc, O, moved=moved, val=val)
