"""Run the JavaScript host side (js/*.js) without Node: the repo's ES5 interpreter (oracle/minijs, test infrastructure) loads the
modules as CommonJS and gets a native `amwg` binding implemented here with ctypes -- the same calls js/amwg_napi.cc makes from a
Node process -- so the JavaScript code drives the real libamwg_b200.so."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.minijs.minijs import Interpreter, JSArray, JSFunction, JSObject, JSThrow, to_js, to_py, undefined  # noqa: E402

JS_DIR = os.path.join(ROOT, "js")


class JsHost:
    def __init__(self, native=None, **interp_kw):
        self.it = Interpreter(**interp_kw)
        self.modules = {}
        it = self.it

        def require(this, a):
            name = it.to_str(a[0])
            key = os.path.basename(name).replace(".js", "")
            if key == "amwg_native":
                if native is None:
                    raise JSThrow("amwg native binding is not available in this host")
                return native(self)
            if key not in self.modules:
                self.load(key)
            return self.modules[key]
        it.set_global("require", it.make_native("require", require))

        def clone(this, a):                                         # snapshot of a live value (step() returns the caller's own state object)
            def cp(v):
                if isinstance(v, JSArray):
                    return it.new_array([cp(x) for x in v.items])
                if isinstance(v, JSObject) and not isinstance(v, JSFunction):
                    o = it.new_object()
                    for k2, x in v.props.items():
                        o.put(k2, cp(x))
                    return o
                return v
            return cp(a[0])
        it.set_global("JSON_clone", it.make_native("JSON_clone", clone))

    def load(self, key):
        it = self.it
        mod = it.new_object()
        mod.put("exports", it.new_object())
        scope = it.global_env.lookup("module")
        outer = scope.vars["module"] if scope is not None else undefined
        it.set_global("module", mod)
        it.run(open(os.path.join(JS_DIR, key + ".js")).read())
        self.modules[key] = mod.get("exports")
        it.set_global("module", outer)
        return self.modules[key]

    def run(self, src):
        return self.it.run(src)

    def get(self, name):
        return self.it.get_global(name)

    def set(self, name, value):
        self.it.set_global(name, to_js(self.it, value) if not isinstance(value, (JSObject, float, str)) else value)


def program_to_py(p):
    """JS Program object -> dict with the fields of tracer.Program"""
    d = to_py(p)
    return d


# ----------------------------------------------------------------------------------------------------------------------
# the `amwg_native` module: what js/amwg_napi.cc exports to Node, implemented over the same C ABI with ctypes
# ----------------------------------------------------------------------------------------------------------------------
def _num_list(v):
    return [float(x) for x in v.items] if isinstance(v, JSArray) else []


def make_model_struct(pkg, desc):
    """JS model descriptor (js/mcmc.js: DeviceModel) -> ctypes amwg_model + keep-alive list: the marshalling of amwg_napi.cc."""
    import ctypes as C
    import numpy as np
    ffi = pkg._ffi
    d = to_py(desc)
    keep = []
    P = len(d["params"])
    prm = (ffi.AmwgParam * P)()
    for k, p in enumerate(d["params"]):
        prm[k] = ffi.AmwgParam(int(p["type"]), int(p["n_comp"]), int(p["dim0"]), int(p["comp_offset"]), float(p["lower"]), float(p["upper"]))
    n_comp = len(d["init"])
    init = np.asarray(d["init"], dtype=np.float64)
    opts = (ffi.AmwgCompOptions * n_comp)()
    for c, o in enumerate(d["comp_options"]):
        opts[c] = ffi.AmwgCompOptions(float(o["prop_log_scale"]), float(o["batch_size"]), float(o["max_adaptation"]), float(o["initial_adaptation"]),
                                      float(o["target_accept_rate"]), int(o["is_adapting"]), 0)

    def i32(name, fallback=(0,)):
        a = np.asarray([int(x) for x in d[name]] if d[name] else list(fallback), dtype=np.int32)
        keep.append(a)
        return a
    code = i32("code")
    consts = np.asarray(d["consts"] if d["consts"] else [0.0], dtype=np.float64)
    cols_np = [np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1)) for c in d["columns"]]
    cols = (ffi.AmwgColumn * max(len(cols_np), 1))()
    for k, c in enumerate(cols_np):
        cols[k] = ffi.AmwgColumn(c.ctypes.data_as(C.POINTER(C.c_double)), c.size)
    plates = (ffi.AmwgPlate * max(len(d["plates"]), 1))()
    for k, pl in enumerate(d["plates"]):
        q = ffi.AmwgPlate()
        q.kind, q.n = int(pl["kind"]), int(pl["n"])
        for j in range(4):
            q.col[j] = int(pl["col"][j]); q.iparam[j] = int(pl["iparam"][j])
        plates[k] = q
    m = ffi.AmwgModel()
    m.abi_version = ffi.ABI_VERSION
    m.n_params, m.params = P, prm
    m.n_comp, m.init = n_comp, init.ctypes.data_as(C.POINTER(C.c_double))
    m.comp_options = opts
    m.n_code, m.code = code.size, code.ctypes.data_as(C.POINTER(C.c_int32))
    m.logpost_prog, m.derived_prog, m.n_derived = int(d["logpost_prog"]), int(d["derived_prog"]), int(d["n_derived"])
    m.n_consts, m.consts = consts.size, consts.ctypes.data_as(C.POINTER(C.c_double))
    m.n_columns, m.columns = len(cols_np), cols
    m.n_plates, m.plates = len(d["plates"]), plates
    pi = C.POINTER(C.c_int32)
    fp, fd = i32("fold_prog"), i32("fold_dst")
    m.n_fold, m.fold_prog, m.fold_dst = len(d["fold_prog"]), fp.ctypes.data_as(pi), fd.ctypes.data_as(pi)
    n_terms = int(d["n_terms"])
    cp, to, tt = i32("comp_prog"), i32("touch_off"), i32("touch_terms")
    m.n_terms = n_terms
    m.comp_prog = cp.ctypes.data_as(pi) if n_terms else None
    m.touch_off = to.ctypes.data_as(pi) if n_terms else None
    m.touch_terms = tt.ctypes.data_as(pi) if n_terms else None
    bp, tbc = i32("block_params"), i32("term_block_comp")
    m.n_block_params = len(d["block_params"])
    m.block_params = bp.ctypes.data_as(pi) if d["block_params"] else None
    m.term_block_comp = tbc.ctypes.data_as(pi) if d["block_params"] else None
    m.stat_prog, m.n_sum_terms = int(d["stat_prog"]), int(d["n_sum_terms"])
    vc, vl, vd = i32("variant_comps"), i32("variant_logpost"), i32("variant_derived", (-1,))
    m.n_variant_comps = len(d["variant_comps"])
    m.variant_comps, m.variant_logpost, m.variant_derived = vc.ctypes.data_as(pi), vl.ctypes.data_as(pi), vd.ctypes.data_as(pi)
    keep += [prm, init, opts, consts, cols_np, cols, plates, m]
    return m, keep, d


class RecordingNative:
    """`amwg_native` without a device: create() keeps the descriptor (tests compare it with the Python host's), nothing runs."""

    def __init__(self):
        self.created = []

    def __call__(self, host):
        it = host.it
        o = it.new_object()

        def create(this, a):
            self.created.append((a[0], [to_py(x) for x in a[1:]]))
            return float(len(self.created))
        o.put("create", it.make_native("create", create))
        o.put("destroy", it.make_native("destroy", lambda this, a: undefined))
        return o


class DeviceNative:
    """`amwg_native` over libamwg_b200.so: the calls of js/amwg_napi.cc, one for one."""

    def __init__(self, pkg):
        self.pkg = pkg
        self.handles = {}
        self.keep = {}
        self.meta = {}

    def _fail(self):
        raise JSThrow(self.pkg._ffi.lib().amwg_last_error().decode())          # the JS side sees a thrown STRING, like the reference's throws

    def __call__(self, host):
        import ctypes as C
        import numpy as np
        it, pkg = host.it, self.pkg
        L = pkg._ffi.lib()
        o = it.new_object()

        def nat(name):
            def deco(fn):
                o.put(name, it.make_native(name, fn))
                return fn
            return deco

        def arr(a):
            return it.new_array([float(x) for x in np.asarray(a).reshape(-1)])

        @nat("create")
        def create(this, a):
            m, keep, d = make_model_struct(pkg, a[0])
            h = C.c_void_p()
            rc = L.amwg_create(C.byref(m), int(a[1]), int(a[2]), int(a[3]), int(a[4]), C.byref(h))
            if rc != 0:
                self._fail()
            key = float(len(self.handles) + 1)
            self.handles[key], self.keep[key] = h, keep
            self.meta[key] = (len(d["init"]), int(d["n_derived"]), int(a[1]))
            return key

        @nat("destroy")
        def destroy(this, a):
            h = self.handles.pop(a[0], None)
            if h is not None:
                L.amwg_destroy(h)
            return undefined

        @nat("burn")
        def burn(this, a):
            if L.amwg_burn(self.handles[a[0]], int(a[1])) != 0:
                self._fail()
            return undefined

        @nat("sample")
        def sample(this, a):
            n, thin, mon = int(a[1]), int(a[2]), np.asarray(_num_list(a[3]), dtype=np.int32)
            C_ = self.meta[a[0]][2]
            rows = 0 if n <= 0 else (n + thin - 1) // thin
            out = np.empty((max(rows, 1), max(mon.size, 1), C_))
            if L.amwg_sample(self.handles[a[0]], n, thin, mon.ctypes.data_as(C.POINTER(C.c_int32)), mon.size, out.ctypes.data) != 0:
                self._fail()
            return arr(out[:rows, :mon.size])

        @nat("get_state")
        def get_state(this, a):
            D, nd, C_ = self.meta[a[0]]
            out = np.empty((D + nd, C_))
            if L.amwg_get_state(self.handles[a[0]], out.ctypes.data) != 0:
                self._fail()
            return arr(out)

        @nat("get_log_post")
        def get_log_post(this, a):
            out = np.empty(self.meta[a[0]][2])
            if L.amwg_get_log_post(self.handles[a[0]], out.ctypes.data) != 0:
                self._fail()
            return arr(out)

        @nat("set_adapting")
        def set_adapting(this, a):
            if L.amwg_set_adapting(self.handles[a[0]], int(a[1])) != 0:
                self._fail()
            return undefined

        @nat("info")
        def info(this, a):
            D, _nd, C_ = self.meta[a[0]]
            scal, pls, acc = np.empty(D * 3), np.empty((D, C_)), np.empty((D, C_), dtype=np.int32)
            if L.amwg_info(self.handles[a[0]], scal.ctypes.data, pls.ctypes.data, acc.ctypes.data) != 0:
                self._fail()
            r = it.new_object()
            r.put("scalars", arr(scal)); r.put("prop_log_scale", arr(pls)); r.put("acceptance_count", arr(acc))
            return r

        @nat("ld_eval")
        def ld_eval(this, a):
            rows = np.asarray([_num_list(r) for r in a[1].items], dtype=np.float64)
            out = np.empty(rows.shape[0])
            if L.amwg_ld_eval(int(a[0]), rows.ctypes.data, rows.shape[1], rows.shape[0], out.ctypes.data, 0) != 0:
                self._fail()
            return arr(out)

        @nat("stream_uniforms")
        def stream_uniforms(this, a):
            seed, chain, first, count = int(a[0]), int(a[1]), int(a[2]), int(a[3])
            want = first + count
            out = np.empty(want)
            if L.amwg_primitive_eval(2, np.zeros(want).ctypes.data, want, seed, chain, out.ctypes.data, 0) != 0:
                self._fail()
            return arr(out[first:])

        @nat("device_log")
        def device_log(this, a):
            x, out = np.array([float(a[0])]), np.empty(1)
            if L.amwg_primitive_eval(0, x.ctypes.data, 1, 0, 0, out.ctypes.data, 0) != 0:
                self._fail()
            return float(out[0])

        @nat("jit_status")
        def jit_status(this, a):
            buf = C.create_string_buffer(1024)
            on = L.amwg_jit_status(self.handles[a[0]], buf, len(buf))
            return ("specialised: " if on else "interpreter: ") + buf.value.decode()
        return o
