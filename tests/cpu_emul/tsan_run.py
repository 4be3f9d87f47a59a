import sys, ctypes as C
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import test_simt_emul as T
from plainmp_b200 import _capi, workloads
from oracle import bridge
lib = C.CDLL(sys.argv[1])
for name, (res, args) in _capi.SYMBOLS.items():
    fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
for mk in (workloads.make_panda, workloads.make_fetch):
    spec, cst = mk()
    t = T.EmuTables(lib, cst)
    Q = workloads.random_configs(spec, 160, 11)
    v = t.check(Q)
    o = bridge.OracleCst(cst)
    print("valid", v.mean(), "oracle", o.is_valid_batch(Q.astype(np.float64)).mean(), flush=True)
    d = t.evaluate(Q[:64])
spec, cst = workloads.make_panda(ceiled=True)
t = T.EmuTables(lib, cst)
o = bridge.OracleCst(cst)
q0, q1 = workloads.random_edges(spec, 64, 3, True, lambda q: o.is_valid_batch(q.astype(np.float64)))
print("edges", t.check_edges(0, np.full(7, 0.05), 0.0, q0, q1)[0].mean(), flush=True)
