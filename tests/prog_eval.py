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


def run(prog, consts, state, pc, O, want_top=False, der=None, moved=-1, val=0.0):
    code, cols, plates = prog.code, prog.columns, prog.plates
    stk = []
    lp = 0.0
    li = ln = 0

    def comp(c):
        return val if c == moved else float(state[c])

    while True:
        w = code[pc]; pc += 1
        op, a = INV[w & 0xff], w >> 8
        if op == "END":
            return stk[-1] if (want_top and stk) else lp
        if op == "CONST": stk.append(consts[a])
        elif op == "COMP": stk.append(comp(a))
        elif op == "DATA": stk.append(float(cols[a][code[pc]])); pc += 1
        elif op == "DATA_I": stk.append(float(cols[a][code[pc] + code[pc + 1] * li])); pc += 2
        elif op == "COMP_I":
            off, stride, base = code[pc:pc + 3]; pc += 3
            stk.append(comp(base + int(cols[a][off + stride * li])))
        elif op in ("ADD", "SUB", "MUL", "DIV", "POW", "LT", "LE", "GT", "GE", "EQ", "NE", "AND", "OR"):
            y = stk.pop(); x = stk.pop()
            if op == "ADD": r = x + y
            elif op == "SUB": r = x - y
            elif op == "MUL": r = x * y
            elif op == "DIV":
                r = (x / y) if y != 0 else (math.nan if (x == 0 or x != x) else math.copysign(math.inf, x) * math.copysign(1.0, y))
            elif op == "POW": r = x * x if y == 2.0 else math.pow(x, y)
            elif op == "LT": r = float(x < y)
            elif op == "LE": r = float(x <= y)
            elif op == "GT": r = float(x > y)
            elif op == "GE": r = float(x >= y)
            elif op == "EQ": r = float(x == y)
            elif op == "NE": r = float(x != y)
            elif op == "AND": r = float(x != 0 and y != 0)
            else: r = float(x != 0 or y != 0)
            stk.append(r)
        elif op == "NEG": stk.append(-stk.pop())
        elif op == "LOG": stk.append(O.orc_log(stk.pop()))
        elif op == "EXP": stk.append(O.orc_exp(stk.pop()))
        elif op == "SQRT":
            x = stk.pop(); stk.append(math.sqrt(x) if x >= 0 else math.nan)
        elif op == "ABS": stk.append(abs(stk.pop()))
        elif op == "NOT": stk.append(float(stk.pop() == 0))
        elif op == "SELECT":
            b = stk.pop(); t = stk.pop(); c = stk.pop(); stk.append(t if c != 0 else b)
        elif op in ("LGAMMA", "LFACTORIAL"):
            stk.append(_ld(O, op.lower(), stk.pop()))
        elif op in ("LCHOOSE", "LBETA"):
            y = stk.pop(); x = stk.pop(); stk.append(_ld(O, op.lower(), x, y))
        elif op.startswith("LD_"):
            n = {"LD_BERN": 2, "LD_POIS": 2, "LD_EXP": 2, "LD_T": 4, "LD_HYPER": 4}.get(op, 3)
            args = stk[-n:]; del stk[-n:]
            stk.append(_ld(O, op[3:].lower(), *args))
        elif op == "ACC": lp = lp + stk.pop()
        elif op == "STORE": der[a] = stk.pop()
        elif op == "LOOP_BEGIN":
            skip = code[pc]; pc += 1
            li, ln = 0, plates[a]["n"]
            if ln <= 0: pc = skip
        elif op == "LOOP_END":
            lp = lp + stk.pop(); li += 1
            if li < ln: pc = a
            else: li = 0
        elif op == "PLATE":
            pl = plates[a]
            n, off = pl["n"], pl["iparam"][2]
            x = np.asarray(cols[pl["col"][0]][off:off + n], dtype=np.float64)
            c0 = -0.5 * O.orc_log(2 * JS_PI)
            if pl["kind"] == PLATE_NORM_IID:
                sd = stk.pop(); mean = stk.pop()
                S = float(np.sum((x - mean) ** 2))
                lp = lp + (n * (c0 - O.orc_log(sd)) - S / (2 * sd * sd))
            elif pl["kind"] == PLATE_BERN_IID:
                p = stk.pop()
                l1 = O.orc_log(1.0 * p + (1 - 1.0) * (1 - p)); l0 = O.orc_log(0.0 * p + (1 - 0.0) * (1 - p))
                for yi in x:
                    lp = lp + (l1 if yi == 1.0 else (l0 if yi == 0.0 else -math.inf))
            elif pl["kind"] == PLATE_NORM_GROUPED:
                sd = stk.pop()
                start = cols[pl["col"][1]].astype(int); base, J = pl["iparam"][0], pl["iparam"][1]
                S = 0.0
                for j in range(J):
                    S += float(np.sum((x[start[j]:start[j + 1]] - comp(base + j)) ** 2))
                lp = lp + (n * (c0 - O.orc_log(sd)) - S / (2 * sd * sd))
            elif pl["kind"] == PLATE_POIS_LOGLIN:
                base, K = pl["iparam"][0], pl["iparam"][1]
                X = np.asarray(cols[pl["col"][1]]).reshape(-1, K)[:n]
                lf = np.asarray(cols[pl["col"][2]])[:n]
                beta = np.array([comp(base + k) for k in range(K)])
                eta = X @ beta
                lp = lp + float(np.sum(x * eta - np.exp(eta) - lf))
            else:
                raise AssertionError("generic plates are LOOP_BEGIN/LOOP_END")
        else:
            raise AssertionError(op)


def logpost(prog, consts, state, O, moved=-1, val=0.0):
    return run(prog, consts, state, prog.logpost_prog, O, moved=moved, val=val)
