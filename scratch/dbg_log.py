import sys, numpy as np
sys.path.insert(0,'.')
import __graft_entry__ as g
pkg=g.load_package(); orc=g.load_oracle()
L,O=pkg._ffi.lib(),orc.lib()
rng=np.random.default_rng(0)
x = np.concatenate([np.exp(rng.uniform(-700, 700, 200000)), rng.uniform(1e-3, 10, 200000),[0.0, -1.0, np.inf, 1.0, 5e-324, 2.2250738585072014e-308, np.nan]])
out=np.empty_like(x)
pkg._ffi.check(L.amwg_primitive_eval(0, x.ctypes.data, x.size, 0, 0, out.ctypes.data, 0))
ref=np.array([O.orc_log(v) for v in x])
bad=np.where(out.view(np.uint64)!=ref.view(np.uint64))[0]
print("log mismatches", len(bad))
for i in bad[:10]: print(i, repr(x[i]), x[i].hex(), repr(out[i]), repr(ref[i]))
x = np.concatenate([rng.uniform(-745, 710, 200000), rng.uniform(-5, 5, 200000), [0.0, -np.inf, np.inf, 709.9, -745.2, 1e-10, np.nan]])
pkg._ffi.check(L.amwg_primitive_eval(1, x.ctypes.data, x.size, 0, 0, out.ctypes.data, 0))
ref=np.array([O.orc_exp(v) for v in x])
bad=np.where(out.view(np.uint64)!=ref.view(np.uint64))[0]
print("exp mismatches", len(bad))
for i in bad[:10]: print(i, repr(x[i]), x[i].hex(), repr(out[i]), repr(ref[i]))
