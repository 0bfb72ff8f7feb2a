"""UPOs.py host logic (Newton steps of the monodromy method, continuation of families, period search) on the CPU:
the GPU entry points underneath are replaced by the oracle-backed stand-in of tests/oracle_context.py.
tests/test_gpu_upos.py runs the same functions through libdicketwa.so on a B200."""
import math

import numpy as np
import pytest

from oracle_context import OracleContext


@pytest.fixture()
def upos(pkg):
    old = pkg._lib._default_ctx
    pkg.set_default_context(OracleContext())
    yield pkg.UPOs
    pkg._lib._default_ctx = old


@pytest.fixture()
def system(pkg):
    return pkg.ClassicalDicke.ClassicalDickeSystem(w=1.0, g=1.0, w0=1.0)


def test_gradient_and_hessian_match_finite_differences_of_the_hamiltonian(pkg, upos, system):
    H = pkg.ClassicalDicke.hamiltonian(system)
    rng = np.random.default_rng(3)
    for _ in range(5):
        u = rng.uniform(-0.9, 0.9, 4)
        e = 1e-6
        g = np.array([(H(u + e * d) - H(u - e * d)) / (2 * e) for d in np.eye(4)])
        assert np.abs(g - upos.hamiltonian_gradient(system, u)).max() < 1e-8
        h = np.array([(upos.hamiltonian_gradient(system, u + e * d) - upos.hamiltonian_gradient(system, u - e * d)) / (2 * e) for d in np.eye(4)])
        assert np.abs(h - upos.hamiltonian_hessian(system, u)).max() < 1e-7
        F = upos.flow(system, u)    # Hamilton's equations: F = J grad H
        assert np.allclose(F, [g[2], g[3], -g[0], -g[1]], atol=1e-8)


def check_orbit(pkg, upos, po, tol=1e-7):
    s = upos.integrate(po)
    assert np.linalg.norm(s.u[-1] - po.u) < tol                      # it closes
    M = upos.monodromy(po)
    J = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [-1, 0, 0, 0], [0, -1, 0, 0]], dtype=float)
    assert np.abs(M.T @ J @ M - J).max() < 1e-6                      # symplectic
    ev = np.linalg.eigvals(M)
    assert np.sum(np.abs(ev - 1) < 1e-3) >= 2                        # flow direction and energy: two unit multipliers
    assert np.abs(M @ upos.flow(po.system, po.u) - upos.flow(po.system, po.u)).max() < 1e-6
    return s


def test_families_A_and_B_start_at_the_normal_mode_periods(pkg, upos, system):
    """the only closed-form pin the reference offers: the families are born at the ground-state fixed point with the
    periods 2 pi / normal_frequency (src/UPOs.jl:816-837, src/ClassicalDicke.jl:139-147)"""
    CD = pkg.ClassicalDicke
    e0 = CD.minimum_energy(system)
    for fam, sgn in ((upos.family_A, "+"), (upos.family_B, "-")):
        get = fam(system, verbose=False)
        T0 = 2 * math.pi / CD.normal_frequency(system, sgn)
        po1, po2 = get(e0 + 0.01), get(e0 + 0.02)
        for po, dE in ((po1, 0.01), (po2, 0.02)):
            assert abs(upos.energy(po) - (e0 + dE)) < 1e-3
            check_orbit(pkg, upos, po)
        # T(E) is smooth: linear extrapolation to the fixed point's energy gives the normal-mode period
        T_at_e0 = po1.T - (upos.energy(po1) - e0) * (po2.T - po1.T) / (upos.energy(po2) - upos.energy(po1))
        assert abs(T_at_e0 - T0) < 2e-3 * T0
        assert get(e0 + 0.01) is not None and abs(get(e0 + 0.01).T - po1.T) < 1e-12     # cached


def test_orbit_utilities(pkg, upos, system):
    CD = pkg.ClassicalDicke
    po = upos.family_A(system, verbose=False)(CD.minimum_energy(system) + 0.05)
    s = check_orbit(pkg, upos, po)
    assert abs(upos.approximate_period(system, po.u, verbose=False) - po.T) < 1e-6
    assert abs(upos.PO(system, po.u).T - po.T) < 1e-6
    assert s.u[300] in po and (s.u[300] + np.array([0.01, 0, 0, 0])) not in po
    assert upos.mirror_Pp(po) == po            # P,p -> -P,-p is time reversal: the same curve
    assert not (upos.mirror_Qq(po) == po)      # the other well
    assert upos.mirror_Qq(po).T == po.T and abs(upos.energy(upos.mirror_Qq(po)) - upos.energy(po)) < 1e-12
    assert 0 <= upos.lyapunov(po) < 1e-4       # family A is stable near its birth
    assert np.allclose(upos.lyapunov_batch([po, upos.mirror_Qq(po)]), upos.lyapunov(po), atol=1e-6)
    # averages: <dH/dp> = <dq/dt> = 0 over a period, and jz from the (Q,P) samples
    assert abs(upos.average_over_PO_qp(po, lambda q, p: p)) < 1e-9
    jz = upos.jz_PO_average(po)
    assert abs(jz - np.mean([pkg.PhaseSpaces.jz(x[0], x[2]) for x in s.u[:-1]])) < 1e-12
    assert abs(upos.average_over_PO(s, lambda x: 1.0) - 1.0) < 1e-12      # a solution object: the reference's sum
    # action: S = \oint P dQ + p dq; dS/dE = T along a family
    po2 = upos.family_A(system, verbose=False)(CD.minimum_energy(system) + 0.06)
    dS, dE = upos.action(po2) - upos.action(po), upos.energy(po2) - upos.energy(po)
    assert abs(dS / dE - 0.5 * (po.T + po2.T)) < 1e-3 * po.T
    assert len(upos.QP(po, npoints=64)) == 65 and len(upos.QPq(po, npoints=8)[0]) == 3
    z = upos.find_p_zero(system, po.u + np.array([0, 0, 0, -1e-3]))
    assert abs(z[3]) < 1e-9 and upos.flow(system, z)[3] > 0
    zn = upos.find_p_zero(system, po.u + np.array([0, 0, 0, 1e-3]), negative=True)
    assert abs(zn[3]) < 1e-9 and upos.flow(system, zn)[3] < 0
    with pytest.raises(NotImplementedError):
        upos.scarring_measure(po)


def test_constant_period_continuation_and_batched_newton(pkg, upos, system):
    CD = pkg.ClassicalDicke
    po = upos.family_A(system, verbose=False)(CD.minimum_energy(system) + 0.05)
    get = upos.follow_PO_family_from_period(po, verbose=False)
    p2 = get(po.T + 0.03)
    assert abs(p2.T - (po.T + 0.03)) < 1e-12 and upos.energy(p2) > upos.energy(po)
    check_orbit(pkg, upos, p2, tol=2e-5)
    rng = np.random.default_rng(0)
    U0 = po.u + 1e-4 * rng.standard_normal((12, 4))
    U, conv, iters = upos.monodromy_method_constant_period_batch(system, U0, po.T, tol=1e-9)
    assert conv.all() and iters.max() <= 10
    for u in U:       # every guess lands on a periodic orbit of that period (a point of the one-parameter family)
        s = pkg.ClassicalSystems.integrate(system, u, po.T, tol=1e-12)
        assert np.linalg.norm(s.u[-1] - u) < 2e-9
    one = upos.monodromy_method_constant_period(system, U0[0], po.T, tol=1e-9)
    assert np.allclose(one.u, U[0], atol=1e-12)
    U, T, conv, iters = upos.monodromy_method_constant_energy_batch(system, U0, po.T)
    assert conv.all()
    H = CD.hamiltonian(system)
    assert np.abs([H(u) - H(u0) for u, u0 in zip(U, U0)]).max() < 1e-9      # correct_energy: stays on each guess's shell
    with pytest.raises(RuntimeError):
        upos.monodromy_method_constant_energy(system, [0.2, 0.1, 0.3, 0.4], 0.37, maxiters=3)
