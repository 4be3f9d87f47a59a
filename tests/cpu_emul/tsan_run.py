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
# cloud scene (index build, 8-lane traversal, nearest-point kernel) and the session's flat kernel
spec, cst = workloads.make_pr2_cloud(1500)
t = T.EmuTables(lib, cst)
Q = workloads.random_configs(spec, 64, 4)
print("pr2 cloud valid", t.check(Q).mean(), flush=True)
P = np.random.default_rng(0).uniform(0.2, 1.4, (64, 3)).astype(np.float32)
sq = np.zeros(64, dtype=np.float32)
lib.pmp_cloud_nearest(t.scene, P.ctypes.data, 64, sq.ctypes.data, None, None)
spec, cst = workloads.make_panda(ceiled=True)
t = T.EmuTables(lib, cst)
from plainmp_b200 import tables
v, keep = tables.make_validator(0, 7, np.full(7, 0.05), 0.0)
ses = C.c_void_p()
lib.pmp_session_create(t.robot, t.scene, C.byref(v), 64, C.byref(ses))
a, b = np.ascontiguousarray(q0[:32], dtype=np.float64), np.ascontiguousarray(q1[:32], dtype=np.float64)
out = np.zeros(32, dtype=np.uint8)
lib.pmp_session_check_motion(ses, a.ctypes.data, b.ctypes.data, 32, out.ctypes.data, None)
print("session edges valid", out.mean(), flush=True)
lib.pmp_session_destroy(ses)
# queue overflow with resume (stress scene), Jacobian rows, scene update path
sys.path.insert(0, '/root/repo/tests')
from fuzz_util import random_case, sample_q, stress_scene
from plainmp_b200 import psdf
for seed in (9004, 9007):
    cst, lb, ub, desc = random_case(seed)
    cst.set_sdf(stress_scene(np.random.default_rng(seed), seed % 3 == 0))
    t = T.EmuTables(lib, cst)
    print("stress", seed, t.check(sample_q(lb, ub, 64, seed)).mean(), flush=True)
cst, lb, ub, desc = random_case(1012)
t = T.EmuTables(lib, cst)
vals, jac = t.evaluate_grad(sample_q(lb, ub, 64, 1012))
print("jacobian rows", jac.shape, flush=True)
spec, cst = workloads.make_panda()
rng = np.random.default_rng(0)
cst.set_sdf(psdf.UnionSDF([psdf.BoxSDF([0.2, 0.3, 0.25], psdf.Pose([0.5, -0.4, 0.3])), psdf.CloudSDF(rng.uniform(-0.15, 0.15, (500, 3)) + [0.45, 0, 0.55], 0.01)]))
t = T.EmuTables(lib, cst)
for n in (900, 200):
    p32 = (rng.uniform(-0.15, 0.15, (n, 3)) + [0.5, 0, 0.55]).astype(np.float32)
    assert lib.pmp_scene_update_cloud(t.scene, p32.ctypes.data, n, 0.01, 0.02) == 0
    print("cloud update", n, t.check(workloads.random_configs(spec, 64, 2)).mean(), flush=True)
# the edge kernel (dynamic refill) on an overflow scene
cst, lb, ub, desc = random_case(9004)
cst.set_sdf(stress_scene(np.random.default_rng(9004), False))
t = T.EmuTables(lib, cst)
qa = sample_q(lb, ub, 48, 1)
qb = np.clip(qa + np.random.default_rng(2).uniform(-0.4, 0.4, qa.shape), lb, ub).astype(np.float32)
print("stress edges", t.check_edges(0, np.full(len(lb), 0.05), 0.0, qa, qb)[0].mean(), flush=True)
