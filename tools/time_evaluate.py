import sys, torch, numpy as np
sys.path.insert(0, '.')
from plainmp_b200 import tables, workloads
dev = torch.device('cuda', 0)
def timeit(fn, steps=5, warmup=2):
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / steps
for name, mk in (("panda", workloads.make_panda), ("fetch", workloads.make_fetch), ("pr2cloud100k", lambda: workloads.make_pr2_cloud(100000)), ("pr2cloud1M", lambda: workloads.make_pr2_cloud(1000000))):
    spec, cst = mk()
    robot, scene = cst._robot(), cst._scene()
    n = (1 << 18) if name.startswith('pr2') else (1 << 20)
    q = torch.from_numpy(workloads.random_configs(spec, n, 1)).to(dev)
    out = torch.empty((n, 2), dtype=torch.float32, device=dev)
    jac = torch.empty((n, 2, q.shape[1]), dtype=torch.float32, device=dev)
    lib = robot._lib
    s = torch.cuda.current_stream().cuda_stream
    t1 = timeit(lambda: lib.pmp_eval_configs(robot.handle, scene.handle, q.data_ptr(), n, out.data_ptr(), s))
    t2 = timeit(lambda: lib.pmp_eval_grad_configs(robot.handle, scene.handle, q.data_ptr(), n, out.data_ptr(), jac.data_ptr(), s))
    print(f"{name}: evaluate {n/t1:.3e}/s, evaluate+jacobian {n/t2:.3e}/s")
