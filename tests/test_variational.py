"""fundamental matrix (get_fundamental_matrix=true, src/ClassicalSystems.jl:110-128) and
lyapunov_spectrum (:183-224).  CPU tests pin the oracle (mpmath/sympy truth, symplecticity, finite
differences); `gpu` tests compare the CUDA kernels (through the C ABI) with the oracle and the truth."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, relerr

DPAR = [1.0, 1.0, 1.0]
LPAR = [1.0, -1.0]
DOCS_PT = [1.0, 0.31783724519578205, 1.0, 0.0]
OMEGA4 = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [-1, 0, 0, 0], [0, -1, 0, 0]], float)   # order (Q, q, P, p)
OMEGA2 = np.array([[0, 1], [-1, 0]], float)


@pytest.fixture(scope="module")
def vgold():
    with open(os.path.join(GOLD, "variational.json")) as f:
        return json.load(f)["cases"]


def _truth(case):
    n = len(case["u0"])
    ts = [r["t"] for r in case["rows"]]
    x = np.array([r["x"] for r in case["rows"]])
    phi = np.array([np.array(r["phi_colmajor"]).reshape(n, n).T for r in case["rows"]])   # -> [r][c]
    return ts, x, phi


# ---------------------------------------------------------------- oracle (CPU)
def test_oracle_fundamental_matrix_matches_mpmath_sympy_truth(O, vgold):
    """hand-written Jacobian (oracle.c) == sympy Jacobian of the reference's equations, integrated to 30 digits"""
    for c in vgold:
        kind = O.SYS_DICKE if c["kind"] == "dicke" else O.SYS_LMG
        ts, x, phi = _truth(c)
        xo, fo, st, _ = O.integrate_fm(kind, c["par"], [c["u0"]], ts, alg=O.ALG_DOP853, tol=1e-13)
        assert st[0] == 0
        assert relerr(xo[0], x).max() < 1e-10
        assert relerr(fo[0], phi).max() < 1e-9, (c["kind"], relerr(fo[0], phi).max())


def test_oracle_fundamental_matrix_is_symplectic_and_matches_finite_differences(O):
    x, F, st, _ = O.integrate_fm(O.SYS_DICKE, DPAR, [DOCS_PT], [0.0, 1.0, 5.0], tol=1e-12)
    assert np.allclose(F[0, 0], np.eye(4))
    for k in range(3):
        M = F[0, k]
        assert np.abs(M.T @ OMEGA4 @ M - OMEGA4).max() < 1e-10
        assert abs(np.linalg.det(M) - 1) < 1e-10
    eps, cols = 1e-6, []
    for c in range(4):
        up, um = np.array(DOCS_PT), np.array(DOCS_PT)
        up[c] += eps
        um[c] -= eps
        a, _, _, _ = O.integrate(O.SYS_DICKE, DPAR, [up], [1.0], tol=1e-13)
        b, _, _, _ = O.integrate(O.SYS_DICKE, DPAR, [um], [1.0], tol=1e-13)
        cols.append((a[0, 0] - b[0, 0]) / (2 * eps))
    assert np.abs(np.array(cols).T - F[0, 1]).max() < 1e-8
    # initial matrix: Phi(t; U) = Phi(t; I) U   (src/ClassicalSystems.jl:113-115)
    U = np.array([[1, 2, 0, 0], [0, 1, 0, 3], [0, 0, 1, 0], [1, 0, 0, 1]], float)
    _, FU, _, _ = O.integrate_fm(O.SYS_DICKE, DPAR, [DOCS_PT], [1.0], fm0=[U], tol=1e-12)
    assert np.abs(FU[0, 0] - F[0, 1] @ U).max() < 1e-8


def test_oracle_lyapunov_spectrum(O):
    """chaotic docs point: lambda_max > 0 and the Hamiltonian pairing (l, ~0, ~0, -l) with a short interval;
    LMG (one degree of freedom, integrable): exponents -> 0 and sum = 0"""
    lam, xf, it, st, ns = O.lyapunov(O.SYS_DICKE, DPAR, [DOCS_PT], t=5.0, tol=1e-10, ltol=1e-4, maxiters=2000)
    assert st[0] == 0 and lam[0, 0] > 0.1
    assert abs(lam[0].sum()) < 1e-3 and abs(lam[0, 0] + lam[0, 3]) < 0.02 and abs(lam[0, 1]) < 0.1   # (the early return mixes rounds k and k-1)
    lam, xf, it, st, ns = O.lyapunov(O.SYS_LMG, LPAR, [[0.5, 0.3]], t=100.0, tol=1e-8)
    assert st[0] == 0 and abs(lam[0]).max() < 0.02 and abs(lam[0].sum()) < 1e-3
    # the reference's defaults (t = 100): only the largest exponent is meaningful in double precision
    lam, xf, it, st, ns = O.lyapunov(O.SYS_DICKE, DPAR, [DOCS_PT], t=100.0, tol=1e-8, ltol=5e-5)
    assert st[0] == 0 and 0.3 < lam[0].max() < 0.5
    # no convergence within maxiters -> status 5 ("Maximum iterations", src/ClassicalSystems.jl:223)
    lam, xf, it, st, ns = O.lyapunov(O.SYS_DICKE, DPAR, [DOCS_PT], t=1.0, tol=1e-8, ltol=1e-12, lreltol=0.0, maxiters=3)
    assert st[0] == 5 and it[0] == 3


# ---------------------------------------------------------------- GPU
def csys(pkg, kind, par):
    s = pkg._lib.System()
    s.kind = kind
    for i, v in enumerate(par):
        s.params[i] = v
    return s


def csol(pkg, alg, dt=0.0, tol=1e-12, backwards=False):
    s = pkg._lib.Solver()
    s.alg, s.backwards, s.dt, s.tol = alg, int(backwards), dt, tol
    return s


@pytest.mark.gpu
def test_gpu_fundamental_matrix_matches_truth_and_oracle(ctx, pkg, O, vgold):
    for c in vgold:
        kind = 0 if c["kind"] == "dicke" else 1
        ts, x, phi = _truth(c)
        for alg, kw, tol_x, tol_f in ((pkg._lib.ALG_DOP853, dict(tol=1e-13), 1e-10, 1e-9), (pkg._lib.ALG_DP5, dict(tol=1e-13), 1e-9, 1e-8),
                                      (pkg._lib.ALG_RK8, dict(dt=2.0 ** -5), 1e-10, 1e-9), (pkg._lib.ALG_RK4, dict(dt=2.0 ** -9), 1e-9, 1e-8)):
            xg, fg, st, ns = ctx.integrate_fm(csys(pkg, kind, c["par"]), csol(pkg, alg, **kw), [c["u0"]], ts)
            assert st[0] == 0 and ns[0] > 0
            assert relerr(xg[0], x).max() < tol_x, (c["kind"], alg, relerr(xg[0], x).max())
            assert relerr(fg[0], phi).max() < tol_f, (c["kind"], alg, relerr(fg[0], phi).max())


@pytest.mark.gpu
@pytest.mark.parametrize("kind,par,name", [(0, DPAR, "dicke"), (1, LPAR, "lmg")])
def test_gpu_fundamental_matrix_batch_matches_oracle(ctx, pkg, O, gold, kind, par, name):
    """matched step / tolerance against the oracle on the golden initial conditions, ragged batch size,
    custom initial matrices, backwards integration; the x part must equal dtwa_integrate's trajectories"""
    g = gold["truth_ld"][name]
    u0 = np.array(g["u0"])[:37]
    nv = u0.shape[1]
    ts = np.array([0.0, 0.5, 1.0, 2.0])
    rng = np.random.default_rng(5)
    fm0 = np.eye(nv)[None] + 0.3 * rng.standard_normal((len(u0), nv, nv))
    okind = O.SYS_DICKE if kind == 0 else O.SYS_LMG
    for alg, oalg, kw in ((pkg._lib.ALG_RK4, O.ALG_RK4, dict(dt=2.0 ** -7)), (pkg._lib.ALG_DOP853, O.ALG_DOP853, dict(tol=1e-11))):
        for backwards in (False, True):
            xg, fg, st, ns = ctx.integrate_fm(csys(pkg, kind, par), csol(pkg, alg, backwards=backwards, **kw), u0, ts, fm0)
            xo, fo, so, no = O.integrate_fm(okind, par, u0, ts, fm0=fm0, alg=oalg, backwards=backwards, **kw)
            assert np.array_equal(st, so)
            ok = st == 0
            assert ok.sum() >= len(u0) - 2
            assert relerr(xg[ok], xo[ok]).max() < 1e-10
            assert relerr(fg[ok], fo[ok], floor=1e-2).max() < 2e-9
            xs, st2, _ = ctx.integrate(csys(pkg, kind, par), csol(pkg, alg, backwards=backwards, **kw), u0, ts)
            if alg == pkg._lib.ALG_RK4:
                assert relerr(xg[ok], xs[ok]).max() < 1e-12     # same steps; the variational RHS is not w-folded
    om = OMEGA4 if kind == 0 else OMEGA2
    xg, fg, st, ns = ctx.integrate_fm(csys(pkg, kind, par), csol(pkg, pkg._lib.ALG_DOP853, tol=1e-12), u0, ts)
    for i in np.flatnonzero(st == 0):
        for k in range(len(ts)):
            assert np.abs(fg[i, k].T @ om @ fg[i, k] - om).max() < 1e-8


@pytest.mark.gpu
def test_gpu_lyapunov_matches_oracle(ctx, pkg, O):
    """a small Lyapunov map (docs/src/ClassicalDickeExamples.md:133-236 parameters): same exponents and the
    same number of renormalisation rounds as the oracle's restatement of the reference loop"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS
    system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    eps = -1.35
    pts = []
    for Q in np.linspace(CD.minimum_nonnegative_Q_for_ϵ(system, eps), CD.maximum_Q_for_ϵ(system, eps), 9)[1:-1]:
        for P in np.linspace(0, CD.maximum_P_for_ϵ(system, eps), 7)[:-1]:
            if CD.minimum_ϵ_for(system, P=P, p=0, Q=Q) <= eps:
                pts.append(CD.Point(system, Q=Q, P=P, p=0, ϵ=eps))
    pts = np.array(pts)
    assert len(pts) >= 15
    lam, status = CS.lyapunov_spectrum_batch(system, pts, 100, tol=1e-8, λtol=5e-5, integrator_alg="DP8")
    info = dict(CS.last_lyapunov_info)
    lo, xo, ito, sto, nso = O.lyapunov(O.SYS_DICKE, [1.0, 0.8, 0.8], pts, t=100, ltol=5e-5, lreltol=1e-3, maxiters=100, tol=1e-8)
    same = (info["iterations"] == ito) & (status == sto)
    assert same.mean() > 0.8          # chaotic orbits: rounding-level differences can shift the stopping round
    conv = same & (status == 0)
    assert conv.sum() >= 10
    assert np.abs(lam[conv].max(axis=1) - lo[conv].max(axis=1)).max() < 1e-2   # (chaotic orbits, single-precision step-size factors: the estimates agree to the convergence tolerance, not better)
    regular = conv & (lo.max(axis=1) < 0.01)
    if regular.any():
        assert np.abs(lam[regular] - lo[regular]).max() < 1e-5
    # drop-in single-point API (src/ClassicalSystems.jl:166-171)
    i = int(np.flatnonzero(conv)[0])
    le = CS.lyapunov_exponent(system, pts[i], tol=1e-8, λtol=5e-5, verbose=False)
    assert abs(le - lam[i].max()) < 1e-12
    with pytest.raises(RuntimeError, match="Maximum iterations"):
        CS.lyapunov_spectrum(system, pts[i], 1.0, tol=1e-8, λtol=1e-13, λreltol=0.0, maxiters=2)


@pytest.mark.gpu
def test_integrate_api_with_fundamental_matrix(pkg, O):
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    s = CS.integrate(system, DOCS_PT, 2.0, get_fundamental_matrix=True, tol=1e-12)
    x, Phi = s.x()
    xo, fo, st, _ = O.integrate_fm(O.SYS_DICKE, DPAR, [DOCS_PT], [0.0, 2.0], tol=1e-12)
    assert relerr(x, xo[0, 1]).max() < 1e-10 and relerr(Phi, fo[0, 1]).max() < 1e-9
    U = np.diag([1.0, 2.0, 3.0, 4.0])
    s2 = CS.integrate(system, DOCS_PT, 2.0, get_fundamental_matrix=True, fundamental_matrix_initial=U, tol=1e-12)
    assert relerr(s2.fundamental_matrix[-1], Phi @ U, floor=1e-2).max() < 1e-8
    with pytest.raises(ValueError):
        CS.integrate(system, DOCS_PT, 2.0, get_fundamental_matrix=True, fundamental_matrix_initial=np.eye(3))


# ---------------------------------------------------------------- surfaces of section
PPAR = [1.0, 0.8, 0.8]   # docs/src/ClassicalDickeExamples.md:86  (w0, w, g)


def test_oracle_section_crossings_lie_on_the_section_and_on_the_energy_shell(O):
    u0 = np.array([[1.0, 0.5, 0.2, 0.0], [0.7, -0.3, 0.1, 0.2]])
    et, eu, ed, cnt, ue, st, ns = O.integrate_events(O.SYS_DICKE, PPAR, u0, 60.0, [0, 0, 0, 1], tol=1e-12, max_events=64)
    assert (st == 0).all() and (cnt >= 8).all()
    for i in range(2):
        k = cnt[i]
        assert np.abs(eu[i, :k, 3]).max() < 1e-13                       # Henon step: g = 0 to rounding
        assert (np.diff(et[i, :k]) > 0).all() and (ed[i, :k][1:] * ed[i, :k][:-1] == -1).all()   # alternate up/down
        h = O.hamiltonian(O.SYS_DICKE, PPAR, eu[i, :k])
        assert np.abs(h - O.hamiltonian(O.SYS_DICKE, PPAR, u0[i:i + 1])[0]).max() < 1e-9
        # the crossing states are states of the trajectory: integrate to the event times and compare
        out, s2, _, _ = O.integrate(O.SYS_DICKE, PPAR, u0[i:i + 1], et[i, :k], tol=1e-12)
        early = et[i, :k] < 15
        assert np.abs(out[0] - eu[i, :k])[early].max() < 1e-9 and np.abs(out[0] - eu[i, :k]).max() < 1e-5   # (chaos amplifies late)
    # direction filter and an offset section (q = 0.25, increasing only)
    et2, eu2, ed2, cnt2, *_ = O.integrate_events(O.SYS_DICKE, PPAR, u0, 60.0, [0, 1, 0, 0], offset=0.25, direction=1, tol=1e-12, max_events=64)
    assert cnt2.sum() >= 4
    for i in range(2):
        if cnt2[i]:
            assert (ed2[i, :cnt2[i]] == 1).all() and np.abs(eu2[i, :cnt2[i], 1] - 0.25).max() < 1e-13
    # capacity: counting continues, storage stops
    et3, eu3, ed3, cnt3, *_ = O.integrate_events(O.SYS_DICKE, PPAR, u0, 60.0, [0, 0, 0, 1], tol=1e-12, max_events=3)
    assert np.array_equal(cnt3, cnt) and np.allclose(et3, et[:, :3])


@pytest.mark.gpu
@pytest.mark.parametrize("alg,kw", [("RK4", dict(dt=2.0 ** -7)), ("RK8", dict(dt=2.0 ** -4)), ("DOP853", dict(tol=1e-11)), ("DP5", dict(tol=1e-11))])
def test_gpu_section_crossings_match_oracle(ctx, pkg, O, alg, kw):
    rng = np.random.default_rng(3)
    u0 = np.column_stack([rng.uniform(-1, 1, 33), rng.uniform(-1, 1, 33), rng.uniform(-0.8, 0.8, 33), rng.uniform(-1, 1, 33)])
    code, ocode = getattr(pkg._lib, "ALG_" + alg), getattr(O, "ALG_" + alg)
    for coef, offset, direction in (([0, 0, 0, 1], 0.0, 0), ([0, 1, 0, 0], 0.25, 1), ([1, 0, -0.5, 0], 0.1, -1)):
        tg, ug, dg, cg, eg, sg, ng, ms = ctx.integrate_events(csys(pkg, 0, PPAR), csol(pkg, code, **kw), u0, 12.0, coef, offset, direction, 32)
        to, uo, do, co, eo, so, no = O.integrate_events(O.SYS_DICKE, PPAR, u0, 12.0, coef, offset, direction, 32, alg=ocode, **kw)
        assert np.array_equal(sg, so) and np.array_equal(cg, co) and np.array_equal(dg, do)
        assert cg.max() <= 32 and cg.sum() > 15
        m = ~np.isnan(to)
        assert np.array_equal(m, ~np.isnan(tg))
        tol_ev = 1e-9 if alg in ("RK4", "RK8") else 2e-8   # adaptive: the step sequences differ at rounding level (float step factor)
        assert np.abs(tg[m] - to[m]).max() < tol_ev and np.abs(ug[m] - uo[m]).max() < tol_ev
        assert relerr(eg, eo).max() < 1e-9
        g = ug[m] @ np.array(coef, float) - offset
        assert np.abs(g).max() < 1e-13
    # LMG: P = 0 section of closed orbits -> crossings repeat with the orbit's period
    v0 = np.array([[0.5, 0.3], [1.2, -0.6]])
    tg, ug, dg, cg, *_ = ctx.integrate_events(csys(pkg, 1, LPAR), csol(pkg, code, **kw), v0, 40.0, [0, 1], 0.0, 1, 64)
    to, uo, do, co, *_ = O.integrate_events(O.SYS_LMG, LPAR, v0, 40.0, [0, 1], 0.0, 1, 64, alg=ocode, **kw)
    assert np.array_equal(cg, co) and (cg >= 3).all()
    for i in range(2):
        per = np.diff(tg[i, :cg[i]])
        assert np.abs(per - per[0]).max() < 1e-5 and np.abs(ug[i, :cg[i], 0] - ug[i, 0, 0]).max() < 1e-5


@pytest.mark.gpu
def test_integrate_with_continuous_callback_like_the_docs(pkg, O):
    """docs/src/ClassicalDickeExamples.md:82-109: save (Q, P) whenever p = 0 and q is the q+ root"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS
    system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    eps = -1.35
    pts = []

    def save(state):
        if CD.q_sign(system, state.u, eps) == "+":
            Q, q, P, p = state.u
            pts.append((Q, P, state.t))
    cb = CS.ContinuousCallback(lambda x, t, _: x[3], save, save_positions=(False, False), abstol=1e-3)
    Q0 = 0.5 * (CD.minimum_nonnegative_Q_for_ϵ(system, eps) + CD.maximum_Q_for_ϵ(system, eps))
    u0 = CD.Point(system, ϵ=eps, P=0, p=0, Q=Q0)
    sol = CS.integrate(system, u0, 200.0, callback=cb, save_everystep=False, tol=1e-10)
    assert sol.retcode == "Success" and len(pts) > 10
    et, eu, ed, cnt, ue, st, ns = O.integrate_events(O.SYS_DICKE, [1.0, 0.8, 0.8], [u0], 200.0, [0, 0, 0, 1], tol=1e-10, max_events=4096)
    want = [(eu[0, k, 0], eu[0, k, 2], et[0, k]) for k in range(cnt[0]) if CD.q_sign(system, eu[0, k], eps) == "+"]
    assert len(want) == len(pts)
    assert np.abs(np.array(want) - np.array(pts)).max() < 1e-6
    assert relerr(sol.u[-1], ue[0]).max() < 1e-6
    with pytest.raises(NotImplementedError, match="affine"):
        CS.integrate(system, u0, 1.0, callback=CS.ContinuousCallback(lambda x, t, _: x[3] ** 2 - 1, save))
    res = CS.poincare_section_batch(system, np.array([u0, u0]), 50.0, coef=[0, 0, 0, 1], tol=1e-10, max_events=8)
    assert np.array_equal(res["count"], [res["count"][0]] * 2) and np.array_equal(res["t"][0], res["t"][1], equal_nan=True)


@pytest.mark.gpu
def test_edge_cases_of_the_widened_entry_points(ctx, pkg):
    """empty batches, bad arguments (the reference's error messages where it has them), failing trajectories"""
    lib = pkg._lib
    sd, sl = csys(pkg, 0, DPAR), csys(pkg, 1, LPAR)
    sol = csol(pkg, lib.ALG_DOP853, tol=1e-9)
    e4, e2 = np.empty((0, 4)), np.empty((0, 2))
    x, F, st, ns = ctx.integrate_fm(sd, sol, e4, [0.0, 1.0])
    assert x.shape == (0, 2, 4) and F.shape == (0, 2, 4, 4)
    lam, xf, it, st, ns, ms = ctx.lyapunov(sl, sol, e2, 10.0, 1e-4, 1e-3, 10)
    assert lam.shape == (0, 2)
    r = ctx.integrate_events(sd, sol, e4, 5.0, [0, 0, 0, 1])
    assert r[0].shape == (0, 1024) and r[3].shape == (0,)
    out, valid, nodes, ms = ctx.shell_quadrature(sd, 0.5, np.empty((0, 2)), ["1"])
    assert out.shape == (0, 1)
    with pytest.raises(pkg.DtwaError, match="interval t must be > 0"):
        ctx.lyapunov(sd, sol, [DOCS_PT], 0.0, 1e-4, 1e-3, 10)
    with pytest.raises(pkg.DtwaError, match="maxiters"):
        ctx.lyapunov(sd, sol, [DOCS_PT], 1.0, 1e-4, 1e-3, 0)
    with pytest.raises(pkg.DtwaError, match="negative time"):
        ctx.integrate_events(sd, sol, [DOCS_PT], -1.0, [0, 0, 0, 1])
    with pytest.raises(pkg.DtwaError, match="non-zero"):
        ctx.integrate_events(sd, sol, [DOCS_PT], 1.0, [0, 0, 0, 0])
    with pytest.raises(pkg.DtwaError, match="direction"):
        ctx.integrate_events(sd, sol, [DOCS_PT], 1.0, [0, 0, 0, 1], direction=2)
    with pytest.raises(pkg.DtwaError, match="Dicke"):
        ctx.shell_quadrature(sl, 0.5, [[0.1, 0.1]], ["1"])
    with pytest.raises(pkg.DtwaError, match="between 1 and 8"):
        ctx.shell_quadrature(sd, 0.5, [[0.1, 0.1]], ["1"] * 9)
    # out-of-bounds initial conditions are reported per trajectory, the rest of the batch is unaffected
    u0 = np.array([DOCS_PT, [1.9, 0.0, 1.9, 0.0], [0.4, -0.7, 0.3, 0.5]])
    x, F, st, ns = ctx.integrate_fm(sd, sol, u0, [0.0, 1.0])
    assert list(st) == [0, lib.TRAJ_BAD_IC, 0] and np.isnan(x[1]).all() and np.isnan(F[1]).all() and np.isfinite(F[[0, 2]]).all()
    lam, xf, it, st, ns, ms = ctx.lyapunov(sd, sol, u0, 5.0, 1e-3, 1e-2, 200)
    assert st[1] == lib.TRAJ_BAD_IC and st[0] == 0 and st[2] == 0
    et, eu, ed, cnt, ue, st, ns, ms = ctx.integrate_events(sd, sol, u0, 10.0, [0, 0, 0, 1], max_events=4)
    assert st[1] == lib.TRAJ_BAD_IC and cnt[1] == 0 and cnt[0] > 0 and cnt[2] > 0
    # t_end = 0: nothing happens
    et, eu, ed, cnt, ue, st, ns, ms = ctx.integrate_events(sd, sol, u0[[0, 2]], 0.0, [0, 0, 0, 1], max_events=4)
    assert (cnt == 0).all() and np.array_equal(ue, u0[[0, 2]]) and (st == 0).all()
    # a section that is never reached
    et, eu, ed, cnt, ue, st, ns, ms = ctx.integrate_events(sd, sol, u0[[0, 2]], 5.0, [0, 1, 0, 0], offset=50.0, max_events=4)
    assert (cnt == 0).all() and np.isnan(et).all()


@pytest.mark.gpu
def test_docs_fotoc_example_grows_at_twice_the_lyapunov_exponent(pkg):
    """docs/src/TWAExamples.md:300-335 end to end: FOTOC = sum of the variances of Q,q,P,p (adaptive, tol=1e-8) starts at
    2ħ and grows like exp(2λt) with λ = lyapunov_exponent(system, x)"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS, TWA
    system = CD.ClassicalDickeSystem(w=1.0, g=1.0, w0=1.0)
    ts = TWA.jrange(0, 0.1, 20)
    j = 1000
    x = CD.Point(system, Q=1, P=0, p=0, ϵ=-0.6)
    W = TWA.coherent_Wigner_HWxSU2(x, j=j, seed=5)
    fotoc = np.array([np.sum(v) for v in TWA.variance(system, observable=["Q", "q", "P", "p"], distribution=W, N=10000, ts=ts, tol=1e-8)])
    lam = CS.lyapunov_exponent(system, x, verbose=False)
    assert abs(fotoc[0] / (2 / j) - 1) < 0.1            # docs/src/TWAExamples.md:334: 2ħ at t = 0
    assert 0.2 < lam < 0.45
    tt = np.asarray(ts)
    slope = np.log(fotoc[60] / fotoc[20]) / (tt[60] - tt[20])
    assert abs(slope / (2 * lam) - 1) < 0.25, (slope, lam)


@pytest.mark.gpu
def test_docs_lyapunov_map_step_with_section_callback(pkg):
    """docs/src/ClassicalDickeExamples.md:180-202: lyapunov_exponent(..., callback=ContinuousCallback(p = 0, save)) returns the
    exponent and fills `pts` with the orbit's Poincare points on the q+ sheet"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, ClassicalSystems as CS
    system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    eps = -1.35
    pts = []

    def save(state):
        if CD.q_sign(system, state.u, eps) == "+":
            Q, q, P, p = state.u
            pts.append((Q, P))
    callback = CS.ContinuousCallback(lambda x, t, _: x[3], save, save_positions=(False, False), abstol=1e-3)
    Q0 = 0.5 * (CD.minimum_nonnegative_Q_for_ϵ(system, eps) + CD.maximum_Q_for_ϵ(system, eps))
    point = CD.Point(system, Q=Q0, P=0.1, p=0, ϵ=eps)
    lam = CS.lyapunov_exponent(system, point, verbose=False, tol=1e-8, λtol=0.00005, callback=callback)
    assert 0.0 <= lam < 0.2 and len(pts) > 20
    rounds = int(CS.last_lyapunov_info["iterations"][0])
    assert rounds >= 2
    Qlo, Qhi, Pm = CD.minimum_nonnegative_Q_for_ϵ(system, eps), CD.maximum_Q_for_ϵ(system, eps), CD.maximum_P_for_ϵ(system, eps)
    arr = np.array(pts)
    assert (np.abs(arr[:, 0]) <= Qhi + 1e-9).all() and (np.abs(arr[:, 1]) <= Pm + 1e-9).all()
    assert lam == CS.lyapunov_exponent(system, point, verbose=False, tol=1e-8, λtol=0.00005)


# ---------------------------------------------------------------- a pin from the reference's own closed forms
def _normal_frequencies_from_phi(Phi, t):
    ang = np.abs(np.angle(np.linalg.eigvals(Phi)))
    return np.sort(ang)[[0, 2]] / t          # the spectrum is {exp(+-i Omega_A t), exp(+-i Omega_B t)}


def test_oracle_linearisation_at_the_ground_state_has_the_references_normal_frequencies(O, pkg):
    """src/ClassicalDicke.jl:139-147 (normal_frequency) and :164-177 (minimum_energy_point) are closed forms of the
    reference; the fundamental matrix at that fixed point is exp(J t), so its eigenvalues are exp(+-i Omega t)"""
    from dickemodel_jl_b200 import ClassicalDicke as CD
    for par in ([1.0, 1.0, 1.0], [1.0, 0.8, 0.8], [2.0, 0.7, 1.3]):
        system = CD.ClassicalDickeSystem(w0=par[0], w=par[1], g=par[2])
        x0 = CD.minimum_energy_point(system)
        want = sorted([CD.normal_frequency(system, "-"), CD.normal_frequency(system, "+")])
        t = 0.5 * np.pi / want[1]
        x, F, st, _ = O.integrate_fm(O.SYS_DICKE, par, [x0], [t], tol=1e-13)
        assert st[0] == 0 and np.abs(x[0, 0] - x0).max() < 1e-12          # a fixed point stays put
        got = _normal_frequencies_from_phi(F[0, 0], t)
        assert np.abs(got - want).max() < 1e-9, (par, got, want)


@pytest.mark.gpu
def test_gpu_linearisation_at_the_ground_state_has_the_references_normal_frequencies(ctx, pkg):
    from dickemodel_jl_b200 import ClassicalDicke as CD
    for par in ([1.0, 1.0, 1.0], [1.0, 0.8, 0.8], [2.0, 0.7, 1.3]):
        system = CD.ClassicalDickeSystem(w0=par[0], w=par[1], g=par[2])
        x0 = CD.minimum_energy_point(system)
        want = sorted([CD.normal_frequency(system, "-"), CD.normal_frequency(system, "+")])
        t = 0.5 * np.pi / want[1]
        for alg, kw in ((pkg._lib.ALG_DOP853, dict(tol=1e-13)), (pkg._lib.ALG_RK4, dict(dt=2.0 ** -9))):
            x, F, st, _ = ctx.integrate_fm(csys(pkg, 0, par), csol(pkg, alg, **kw), [x0], [t])
            assert st[0] == 0 and np.abs(x[0, 0] - x0).max() < 1e-11
            got = _normal_frequencies_from_phi(F[0, 0], t)
            assert np.abs(got - want).max() < 1e-8, (par, alg, got, want)


def _near_tangent_section(O):
    """a section q = q_max - eps just below a local maximum of q(t): two crossings 0.08 apart, inside one DOP853 step"""
    u0 = np.array([[1.0, 0.5, 0.2, 0.0]])
    ts = np.linspace(18.0, 19.6, 1601)
    out, st, _, _ = O.integrate(O.SYS_DICKE, PPAR, u0, ts, tol=1e-13)
    q = out[0, :, 1]
    i = int(np.argmax(q))
    assert 0 < i < len(q) - 1
    return u0, float(ts[i]), float(q[i])


def test_oracle_finds_both_crossings_of_a_near_tangency(O):
    u0, tmax, qmax = _near_tangent_section(O)
    off = qmax - 1e-3
    truth = O.integrate_events(O.SYS_DICKE, PPAR, u0, 30.0, [0, 1, 0, 0], offset=off, tol=1e-13, max_events=64)
    tt = truth[0][0, :truth[3][0]]
    pair = tt[np.abs(tt - tmax) < 0.3]
    assert len(pair) == 2 and 0.05 < pair[1] - pair[0] < 0.12
    for tol in (1e-8, 1e-10):      # steps of 0.1 - 0.25: the pair sits inside one step, no sign change at its ends
        ev = O.integrate_events(O.SYS_DICKE, PPAR, u0, 30.0, [0, 1, 0, 0], offset=off, tol=tol, max_events=64)
        te, de = ev[0][0, :ev[3][0]], ev[2][0, :ev[3][0]]
        got = te[np.abs(te - tmax) < 0.3]
        assert len(got) == 2 and np.abs(got - pair).max() < 0.01, (tol, got, pair)
        assert list(de[np.abs(te - tmax) < 0.3]) == [1, -1]
        assert np.abs(ev[1][0, :ev[3][0], 1] - off).max() < 1e-13
    # direction filters keep one of the two
    up = O.integrate_events(O.SYS_DICKE, PPAR, u0, 30.0, [0, 1, 0, 0], offset=off, direction=1, tol=1e-10, max_events=64)
    tu = up[0][0, :up[3][0]]
    assert len(tu[np.abs(tu - tmax) < 0.3]) == 1 and abs(tu[np.abs(tu - tmax) < 0.3][0] - pair[0]) < 0.01


@pytest.mark.gpu
def test_gpu_finds_both_crossings_of_a_near_tangency(ctx, pkg, O):
    u0, tmax, qmax = _near_tangent_section(O)
    off = qmax - 1e-3
    for alg, ocode, tol in ((pkg._lib.ALG_DOP853, O.ALG_DOP853, 1e-10), (pkg._lib.ALG_DP5, O.ALG_DP5, 1e-10)):
        tg, ug, dg, cg, eg, sg, ng, ms = ctx.integrate_events(csys(pkg, 0, PPAR), csol(pkg, alg, tol=tol), u0, 30.0, [0, 1, 0, 0], off, 0, 64)
        to, uo, do, co, eo, so, no = O.integrate_events(O.SYS_DICKE, PPAR, u0, 30.0, [0, 1, 0, 0], offset=off, tol=tol, max_events=64, alg=ocode)
        assert cg[0] == co[0] and np.array_equal(dg[0, :cg[0]], do[0, :co[0]])
        got = tg[0, :cg[0]]
        assert (np.abs(got - tmax) < 0.3).sum() == 2
        assert np.abs(got - to[0, :co[0]]).max() < 1e-6        # (crossing times near a tangency amplify the rounding-level step differences)


@pytest.mark.gpu
def test_packed_section_output_equals_the_padded_one(ctx, pkg):
    """dtwa_integrate_events_packed: the same events, compacted on the GPU into CSR form (rows of very different
    lengths, a row with no crossing, rows cut at max_events)"""
    CD, CS = pkg.ClassicalDicke, pkg.ClassicalSystems
    system = CD.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    rng = np.random.default_rng(11)
    X = np.column_stack([rng.uniform(0.7, 1.4, 203), rng.uniform(-0.5, 0.5, 203), rng.uniform(-0.4, 0.4, 203), rng.uniform(-0.5, 0.5, 203)])
    X[7] = CD.minimum_energy_point(system, "+")      # the fixed point never crosses p = 0.3
    for kw in (dict(coef=[0, 0, 0, 1], offset=0.3, max_events=15), dict(coef=[0, 1, 0, 0.5], direction=1, max_events=7)):
        a = CS.poincare_section_batch(system, X, 60.0, tol=1e-9, **kw)
        b = CS.poincare_section_batch(system, X, 60.0, tol=1e-9, packed=True, **kw)
        assert np.array_equal(a["count"], b["count"]) and np.array_equal(a["u_end"], b["u_end"]) and np.array_equal(a["nsteps"], b["nsteps"])
        m = np.minimum(a["count"], kw["max_events"])
        assert np.array_equal(np.diff(b["offsets"]), m) and b["offsets"][0] == 0 and len(b["t"]) == m.sum()
        assert m.min() == 0 and m.max() == kw["max_events"] and (a["count"] > kw["max_events"]).any()
        assert kw["max_events"] != 15 or len(set(m.tolist())) > 3
        for i in (0, 7, 50, 202):
            lo, hi = b["offsets"][i], b["offsets"][i + 1]
            assert np.array_equal(b["t"][lo:hi], a["t"][i, :m[i]]) and np.array_equal(b["u"][lo:hi], a["u"][i, :m[i]])
            assert np.array_equal(b["direction"][lo:hi], a["direction"][i, :m[i]])
        mask = np.arange(kw["max_events"])[None, :] < m[:, None]
        assert np.array_equal(a["t"][mask], b["t"]) and np.array_equal(a["u"][mask], b["u"])
    e = CS.poincare_section_batch(system, X[:0], 60.0, coef=[0, 0, 0, 1], packed=True)
    assert len(e["t"]) == 0 and e["offsets"].tolist() == [0]


@pytest.mark.gpu
def test_packed_fetch_rejects_a_stale_or_wrong_total(ctx, pkg):
    import ctypes as C
    L = pkg._lib.load()
    buf = np.empty(64)
    idir = np.empty(64, dtype=np.int32)
    rc = L.dtwa_integrate_events_packed_fetch(ctx._h, 12345, buf.ctypes.data_as(C.POINTER(C.c_double)), buf.ctypes.data_as(C.POINTER(C.c_double)),
                                              idir.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc != 0 and b"packed" in L.dtwa_last_error()
    system = pkg.ClassicalDicke.ClassicalDickeSystem(w=0.8, g=0.8, w0=1)
    r = pkg.ClassicalSystems.poincare_section_batch(system, [[1.0, 0.2, 0.1, 0.0]], 20.0, coef=[0, 0, 0, 1], packed=True, max_events=0)
    assert len(r["t"]) == 0 and r["count"][0] > 0 and r["offsets"].tolist() == [0, 0]
