"""One process per GPU over NCCL (the bench.py --gpus N path): needs >= 2 GPUs, skipped otherwise."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DTWA_ROOT"])
import __graft_entry__ as graft
pkg = graft.load_package()
from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg.set_default_context(pkg.Context([local], stream=torch.cuda.current_stream().cuda_stream))
pkg.set_distributed(rank, world, pkg.distributed.torch_allreduce(local))
system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
N, j = 200001, 100
ts = TWA.jrange(0, 0.1, 5)
jz = TWA.Weyl.Jz(j) / j
W = TWA.coherent_Wigner_HWxSU2([1.0, 0.31783724519578205, 1.0, 0.0], j=j, seed=5)
chain = TWA.mcs_chain(TWA.mcs_for_averaging(system, observable=jz, N=N, ts=ts, distribution=W),
                      TWA.mcs_for_distributions(system, x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1), N=N, ts=ts, distribution=W))
avg, hist = TWA.monte_carlo_integrate(system, chain, integrator_alg="RK4", dt=2.0 ** -6)
if rank == 0:
    print(json.dumps({"avg": avg.tolist(), "hist": np.rint(hist * N).astype(int).tolist(), "steps": TWA.last_info["steps_accepted"],
                      "n_valid": [int(x) for x in TWA.last_info["n_valid"]]}))
dist.destroy_process_group()
'''


def test_two_ranks_reproduce_the_single_rank_result(tmp_path, pkg):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DTWA_ROOT=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    two = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    N, j = 200001, 100
    ts = TWA.jrange(0, 0.1, 5)
    jz = TWA.Weyl.Jz(j) / j
    W = TWA.coherent_Wigner_HWxSU2([1.0, 0.31783724519578205, 1.0, 0.0], j=j, seed=5)
    chain = TWA.mcs_chain(TWA.mcs_for_averaging(system, observable=jz, N=N, ts=ts, distribution=W),
                          TWA.mcs_for_distributions(system, x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1), N=N, ts=ts, distribution=W))
    avg, hist = TWA.monte_carlo_integrate(system, chain, integrator_alg="RK4", dt=2.0 ** -6)
    assert np.array_equal(np.rint(hist * N).astype(int), np.array(two["hist"]))          # integer counts: bit-identical for any rank count
    assert np.abs(avg - np.array(two["avg"])).max() < 1e-12
    assert TWA.last_info["steps_accepted"] == two["steps"] and [int(x) for x in TWA.last_info["n_valid"]] == two["n_valid"]


def test_independent_batches_split_over_two_gpus(pkg):
    """Lyapunov maps, sections and dense integration have no exchange step: devices=[0, 1] gives each GPU a contiguous
    block of initial conditions, and the results equal the single-GPU ones bit for bit"""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS
    system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    rng = np.random.default_rng(9)
    X = np.column_stack([rng.uniform(0.7, 1.4, 301), rng.uniform(-0.5, 0.5, 301), rng.uniform(-0.4, 0.4, 301), rng.uniform(-0.5, 0.5, 301)])
    l1, s1 = CS.lyapunov_spectrum_batch(system, X, 20, tol=1e-8, maxiters=30)
    l2, s2 = CS.lyapunov_spectrum_batch(system, X, 20, tol=1e-8, maxiters=30, devices=[0, 1])
    assert np.array_equal(l1, l2) and np.array_equal(s1, s2)
    e1 = CS.poincare_section_batch(system, X, 30.0, coef=[0, 0, 0, 1], tol=1e-9, max_events=32)
    e2 = CS.poincare_section_batch(system, X, 30.0, coef=[0, 0, 0, 1], tol=1e-9, max_events=32, devices=[0, 1])
    assert np.array_equal(e1["count"], e2["count"]) and np.array_equal(e1["t"], e2["t"], equal_nan=True)
    p1 = CS.poincare_section_batch(system, X, 30.0, coef=[0, 0, 0, 1], tol=1e-9, max_events=32, packed=True)
    p2 = CS.poincare_section_batch(system, X, 30.0, coef=[0, 0, 0, 1], tol=1e-9, max_events=32, packed=True, devices=[0, 1])
    assert all(np.array_equal(p1[k], p2[k]) for k in ("offsets", "t", "u", "direction", "count", "u_end"))
    o1 = CS.integrate_batch(system, X, [0.0, 1.0, 2.0], tol=1e-9)
    o2 = CS.integrate_batch(system, X, [0.0, 1.0, 2.0], tol=1e-9, devices=[0, 1])
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(o1, o2))
    f1 = CS.integrate_fm_batch(system, X[:65], [0.0, 1.5], tol=1e-9)
    f2 = CS.integrate_fm_batch(system, X[:65], [0.0, 1.5], tol=1e-9, devices=[0, 1])
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(f1, f2))
    from dickemodel_jl_b200 import EnergyShellProjections as ESP
    m1 = ESP.matrix_QP_iint_dqdp_delta(system, f=lambda x: [1, x[1] ** 2], ϵ=-0.5, res=0.05)
    m2 = ESP.matrix_QP_iint_dqdp_delta(system, f=lambda x: [1, x[1] ** 2], ϵ=-0.5, res=0.05, devices=[0, 1])
    assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(m1, m2))


def test_batch_entry_points_split_inside_a_multi_device_context(pkg):
    """a context created on two devices (what the Julia shim holds) gives each device one contiguous block of the
    items of dtwa_integrate / _fm / _lyapunov / _integrate_events(_packed) / _shell_quadrature: same bits as one GPU"""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    L = pkg._lib
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS, EnergyShellProjections as ESP
    system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    rng = np.random.default_rng(5)
    X = np.column_stack([rng.uniform(0.7, 1.4, 257), rng.uniform(-0.5, 0.5, 257), rng.uniform(-0.4, 0.4, 257), rng.uniform(-0.5, 0.5, 257)])
    one, two = L.device_context(0), L.Context([0, 1])
    sys_c = system._c_system()
    sol = CS.solver_from_kargs(dict(tol=1e-9))
    try:
        for name, args in (("integrate", (sys_c, sol, X, [0.0, 1.0, 2.5])), ("integrate_fm", (sys_c, sol, X[:101], [0.0, 1.5])),
                           ("lyapunov", (sys_c, sol, X, 20.0, 1e-4, 1e-3, 30)),
                           ("integrate_events", (sys_c, sol, X, 30.0, [0, 0, 0, 1], 0.1, 0, 16)),
                           ("integrate_events_packed", (sys_c, sol, X, 30.0, [0, 0, 0, 1], 0.1, 0, 16)),
                           ("shell_quadrature", (sys_c, -0.5, np.column_stack([rng.uniform(-1, 1, 333), rng.uniform(-1, 1, 333)]), ["1", "q^2"], 0.02, 0))):
            a, b = getattr(one, name)(*args), getattr(two, name)(*args)
            for x, y in zip(a, b):
                if isinstance(x, np.ndarray):
                    assert np.array_equal(x, y, equal_nan=True), name
        # a batch smaller than 2 x devices stays on the first device; errors of a worker thread reach the caller
        a, b = one.integrate(sys_c, sol, X[:3], [0.0, 1.0]), two.integrate(sys_c, sol, X[:3], [0.0, 1.0])
        assert np.array_equal(a[0], b[0])
        with pytest.raises(L.DtwaError):
            two.shell_quadrature(sys_c, -0.5, np.zeros((64, 2)), ["1 +"], 0.02, 0)
    finally:
        two.close()
