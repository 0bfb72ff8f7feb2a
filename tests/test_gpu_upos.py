"""UPOs.py through libdicketwa.so on a B200: the same orbits as the oracle-backed run of the host logic
(tests/test_upos_host.py), and the batched monodromy method on a few hundred guesses."""
import math

import numpy as np
import pytest

from oracle_context import OracleContext

pytestmark = pytest.mark.gpu


@pytest.fixture()
def system(pkg):
    return pkg.ClassicalDicke.ClassicalDickeSystem(w=1.0, g=1.0, w0=1.0)


def test_family_orbits_match_the_oracle_backed_run(pkg, ctx, system):
    U, CD = pkg.UPOs, pkg.ClassicalDicke
    e0 = CD.minimum_energy(system)
    got = {}
    for name, c in (("gpu", ctx), ("oracle", OracleContext())):
        pkg.set_default_context(c)
        try:
            pa = U.family_A(system, verbose=False)(e0 + 0.05)
            pb = U.family_B(system, verbose=False)(e0 + 0.3)
            got[name] = (pa, pb, U.lyapunov(pa), U.action(pb), U.approximate_period(system, pb.u, verbose=False),
                         U.find_p_zero(system, pa.u + np.array([0, 0, 0, -1e-3])), U.jz_PO_average(pb))
        finally:
            pkg.set_default_context(ctx)
    (ga, gb, gl, gs, gt, gz, gj), (oa, ob, ol, os_, ot, oz, oj) = got["gpu"], got["oracle"]
    # same Newton iterations on trajectories that agree to ~1e-12: the orbits agree far below the method's tolerance
    assert np.abs(ga.u - oa.u).max() < 1e-7 and abs(ga.T - oa.T) < 1e-7
    assert np.abs(gb.u - ob.u).max() < 1e-7 and abs(gb.T - ob.T) < 1e-7
    assert abs(gl - ol) < 1e-6 and abs(gs - os_) < 1e-7 and abs(gt - ot) < 1e-7 and abs(gj - oj) < 1e-8
    assert np.abs(gz - oz).max() < 1e-7
    s = U.integrate(gb)
    assert np.linalg.norm(s.u[-1] - gb.u) < 1e-7
    assert abs(2 * math.pi / CD.normal_frequency(system, "-") - gb.T) < 0.2     # still close to the normal-mode period
    assert s.u[100] in gb and U.mirror_Pp(gb) == gb


def test_batched_monodromy_method(pkg, ctx, system):
    """300 perturbed guesses refined together: one launch of the variational kernel per Newton iteration"""
    U, CD = pkg.UPOs, pkg.ClassicalDicke
    po = U.family_B(system, verbose=False)(CD.minimum_energy(system) + 0.3)
    rng = np.random.default_rng(1)
    U0 = po.u + 1e-4 * rng.standard_normal((300, 4))
    Ub, Tb, conv, iters = U.monodromy_method_constant_energy_batch(system, U0, po.T, tol=1e-9)
    assert conv.all() and iters.max() <= 12
    x, _, status, _ = pkg.ClassicalSystems.integrate_fm_batch(system, Ub[:8], [0.0, float(Tb[0])], tol=1e-12)
    assert status[0] == 0 and np.linalg.norm(x[0, -1] - Ub[0]) < 1e-8
    lam = U.lyapunov_batch([U.PO(system, u, T) for u, T in zip(Ub[:16], Tb[:16])])
    assert np.all(np.abs(lam - U.lyapunov(po)) < 1e-4)
