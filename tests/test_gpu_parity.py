"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (north_star): individual trajectories within 1e-10 relative at matched step/tolerance over
short times; integer work (bin indices, histogram counts, Philox words) bit-exact; observable values
built from + - * / sqrt and integer powers bit-exact (single IEEE operations on both sides).
"""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

DPAR = [1.0, 1.0, 1.0]
LPAR = [1.0, -1.0]
DOCS_PT = [1.0, 0.31783724519578205, 1.0, 0.0]


def csys(pkg, kind, par):
    s = pkg._lib.System()
    s.kind = kind
    for i, v in enumerate(par):
        s.params[i] = v
    return s


def csol(pkg, alg, dt=0.0, tol=1e-12, backwards=False, dtmin=0.0, maxiters=0):
    s = pkg._lib.Solver()
    s.alg, s.backwards, s.dt, s.tol, s.dtmin, s.maxiters = alg, int(backwards), dt, tol, dtmin, maxiters
    return s


def csmp(pkg, kind, center, hbar, seed=0, offset=0):
    s = pkg._lib.Sampler()
    s.kind = kind
    for i, v in enumerate(center):
        s.center[i] = v
    s.hbar, s.seed, s.offset = hbar, seed, offset
    return s


def test_device_is_blackwell(ctx):
    info = ctx.device_info()
    assert info["cc"] >= 100 and info["sm_count"] > 0


def test_hamiltonian_bit_exact(ctx, pkg, O, gold):
    u = np.array(gold["truth_ld"]["dicke"]["u0"])
    assert np.array_equal(ctx.hamiltonian(csys(pkg, 0, DPAR), u), O.hamiltonian(O.SYS_DICKE, DPAR, u))
    v = np.array(gold["truth_ld"]["lmg"]["u0"])
    assert np.array_equal(ctx.hamiltonian(csys(pkg, 1, LPAR), v), O.hamiltonian(O.SYS_LMG, LPAR, v))


@pytest.mark.parametrize("kind,par,name", [(0, DPAR, "dicke"), (1, LPAR, "lmg")])
@pytest.mark.parametrize("alg,dt", [("RK4", 2.0 ** -7), ("RK8", 2.0 ** -4)])
def test_fixed_step_trajectories_match_oracle(ctx, pkg, O, gold, kind, par, name, alg, dt):
    g = gold["truth_ld"][name]
    ts = np.array(g["ts"])
    code = getattr(pkg._lib, "ALG_" + alg)
    out, st, ns = ctx.integrate(csys(pkg, kind, par), csol(pkg, code, dt=dt), g["u0"], ts)
    ref, rst, acc, _ = O.integrate(kind, par, g["u0"], ts, alg=getattr(O, "ALG_" + alg), dt=dt)
    assert np.array_equal(st, rst) and (st == 0).all()
    assert np.array_equal(ns, acc)
    # same algorithm, same step table: the only differences are FMA contraction and the rsqrt polish
    assert relerr(out, ref).max() < 1e-10


@pytest.mark.parametrize("kind,par,name", [(0, DPAR, "dicke"), (1, LPAR, "lmg")])
@pytest.mark.parametrize("alg,tol", [("DOP853", 1e-12), ("DP5", 1e-10), ("DOP853", 1e-8), ("TSIT5", 1e-10), ("TSIT5", 1e-12)])
def test_adaptive_trajectories_match_oracle_and_truth(ctx, pkg, O, gold, kind, par, name, alg, tol):
    g = gold["truth_ld"][name]
    ts = np.array(g["ts"])
    out, st, ns = ctx.integrate(csys(pkg, kind, par), csol(pkg, getattr(pkg._lib, "ALG_" + alg), tol=tol), g["u0"], ts)
    ref, rst, acc, rej = O.integrate(kind, par, g["u0"], ts, alg=getattr(O, "ALG_" + alg), tol=tol)
    assert (st == 0).all() and (rst == 0).all()
    short = ts <= 2.0
    # matched tolerance: both follow the same controller; a last-bit difference in an error norm may
    # flip one accept/reject decision, so compare with the tolerance-level bound, not bit for bit
    assert relerr(out[:, short], ref[:, short]).max() < max(1e-10, 100 * tol)
    assert np.abs(ns - (acc + rej)).max() <= max(3, 0.02 * (acc + rej).max())
    if tol <= 1e-12:
        truth = np.array(g["u"])
        assert relerr(out[:, short], truth[:, short]).max() < 1e-10   # north_star bound


def test_hermite_output_mode_matches_the_oracle(ctx, pkg, O, gold):
    """DTWA_OUTPUT_HERMITE (the reference's interpolated outputs, src/TWA.jl:180): same un-clipped step sequence and the same
    interpolant as the oracle's restatement -> same values to the tolerance level, same step counts to within 2 %"""
    g = gold["truth_ld"]["dicke"]
    u0 = np.array(g["u0"])[:96]
    ts = np.arange(0, 161) * 0.05
    for alg, oalg in ((pkg._lib.ALG_DOP853, O.ALG_DOP853), (pkg._lib.ALG_DP5, O.ALG_DP5), (pkg._lib.ALG_TSIT5, O.ALG_TSIT5)):
        sol = csol(pkg, alg, tol=1e-10)
        sol.output = pkg._lib.OUTPUT_HERMITE
        out, st, ns = ctx.integrate(csys(pkg, 0, g["par"]), sol, u0, ts)
        ref, rst, acc, rej = O.integrate(O.SYS_DICKE, g["par"], u0, ts, alg=oalg, tol=1e-10, output="hermite")
        assert (st == 0).all() and (rst == 0).all()
        assert relerr(out, ref).max() < 2e-5       # (Hermite error ~1e-6 at the outputs; it moves with the last-digit differences of the step sizes)
        assert relerr(out[:, -1], ref[:, -1]).max() < 1e-8
        assert abs(ns.sum() - (acc + rej).sum()) <= 0.02 * (acc + rej).sum()
        stepped, _, ns2 = ctx.integrate(csys(pkg, 0, g["par"]), csol(pkg, alg, tol=1e-10), u0, ts)
        assert ns.sum() < 0.92 * ns2.sum() and 1e-9 < np.abs(out - stepped).max() < 1e-3


def test_docs_point_against_mpmath(ctx, pkg, gold):
    c = gold["truth_mpmath"]["cases"][0]
    ts = np.array([float(t) for t in c["times"]])
    out, st, _ = ctx.integrate(csys(pkg, 0, c["par"]), csol(pkg, pkg._lib.ALG_DOP853, tol=1e-13), [c["u0"]], ts)
    ref = np.array([[float(x) for x in row] for row in c["u"]])
    err = relerr(out[0], ref).max(axis=1)
    assert st[0] == 0 and err[ts <= 10].max() < 1e-10


def test_backwards_returns_to_start(ctx, pkg, gold):
    u0 = np.array(gold["truth_ld"]["dicke"]["u0"])[:64]
    sysd = csys(pkg, 0, DPAR)
    fwd, st, _ = ctx.integrate(sysd, csol(pkg, pkg._lib.ALG_DOP853, tol=1e-12), u0, [5.0])
    back, st2, _ = ctx.integrate(sysd, csol(pkg, pkg._lib.ALG_DOP853, tol=1e-12, backwards=True), fwd[:, 0], [5.0])
    assert (st == 0).all() and (st2 == 0).all()
    assert relerr(back[:, 0], u0).max() < 1e-8


def test_out_of_bounds_and_domain_status(ctx, pkg):
    sysd = csys(pkg, 0, DPAR)
    u0 = np.array([[1.5, 0.0, 1.5, 0.0], [0.1, 0.1, 0.1, 0.1], [np.nan, 0, 0, 0]])
    out, st, _ = ctx.integrate(sysd, csol(pkg, pkg._lib.ALG_RK4, dt=0.01), u0, [0.0, 1.0])
    assert list(st) == [pkg._lib.TRAJ_BAD_IC, pkg._lib.TRAJ_OK, pkg._lib.TRAJ_BAD_IC]
    assert np.isnan(out[0]).all() and np.isfinite(out[1]).all()


def test_invalid_arguments_raise(ctx, pkg):
    sysd = csys(pkg, 0, DPAR)
    with pytest.raises(pkg.DtwaError, match="negative time"):
        ctx.integrate(sysd, csol(pkg, pkg._lib.ALG_RK4, dt=0.01), [[0, 0, 0, 0]], [-1.0])
    with pytest.raises(pkg.DtwaError, match="sorted"):
        ctx.integrate(sysd, csol(pkg, pkg._lib.ALG_RK4, dt=0.01), [[0, 0, 0, 0]], [1.0, 0.5])
    with pytest.raises(pkg.DtwaError, match="dt > 0"):
        ctx.integrate(sysd, csol(pkg, pkg._lib.ALG_RK4, dt=0.0), [[0, 0, 0, 0]], [1.0])
    with pytest.raises(pkg.DtwaError, match="unknown variable"):
        ctx.eval_expr(sysd, "Q + z", [[0, 0, 0, 0]])


@pytest.mark.parametrize("kind,center", [("HWXSU2", [0.0, 0.0, 0.0, 0.0]), ("HWXSU2", DOCS_PT), ("SU2", [0.3, -0.4]), ("HW", [0.5, -0.2])])
def test_sampler_matches_oracle(ctx, pkg, O, kind, center):
    n, hb = 20000, 1 / 100
    smp = csmp(pkg, getattr(pkg._lib, "SAMPLER_COHERENT_" + kind), center, hb, seed=1234567890123, offset=2 ** 33 + 5)
    got = ctx.sample(smp, n, attempt=2)
    want = O.sample(getattr(O, "SAMPLER_" + kind), center, hb, n, seed=1234567890123, offset=2 ** 33 + 5, attempt=2)
    # identical Philox words and 53-bit uniforms; CUDA's and glibc's log/sin/cos/atan2 differ by an ulp or two
    assert np.abs(got - want).max() < 5e-14


def test_pdf_matches_oracle(ctx, pkg, O):
    hb = 1 / 100
    u = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, hb, 5000, seed=4)
    got = ctx.pdf(csmp(pkg, pkg._lib.SAMPLER_COHERENT_HWXSU2, DOCS_PT, hb), u)
    want = O.pdf(O.SAMPLER_HWXSU2, DOCS_PT, hb, u)
    assert relerr(got, want, floor=1e-300).max() < 1e-10
    assert ctx.pdf(csmp(pkg, pkg._lib.SAMPLER_COHERENT_SU2, [0.3, -0.4], hb), [[3.0, 0.0]])[0] == 0.0


def test_expressions_bit_exact(ctx, pkg, O, gold):
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    u = np.array(gold["truth_ld"]["dicke"]["u0"])
    sysd = csys(pkg, 0, DPAR)
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j = 100
    cases = [TWA.Weyl.Jz(j) / j, TWA.Weyl.Jx(j), TWA.Weyl.Jy(j), TWA.Weyl.Jz2(j), TWA.Weyl.Jx2(j), TWA.Weyl.Jy2(j),
             TWA.Weyl.n(j), TWA.Weyl.n2(j), "q+p^2-Q", "2Q - 3(q+1)", "-Q^2", "Q^5 - P^-3 + q^7/p^4",
             TWA._expr_of(system, CD.hamiltonian(system)), "(1/2)*sqrt(4 - Q^2 - P^2)"]
    for e in cases:
        e = str(e)
        got = ctx.eval_expr(sysd, e, u)
        want = O.eval_expr(e.replace("2Q", "2*Q").replace("3(", "3*("), O.varnames(O.SYS_DICKE), u)
        assert np.array_equal(got, want), e
    # the oracle's own Weyl strings (written independently from the source formulas) agree as well
    for a, b in ((TWA.Weyl.Jz(j), O.Weyl.Jz(j)), (TWA.Weyl.Jx2(j), O.Weyl.Jx2(j)), (TWA.Weyl.n2(j), O.Weyl.n2(j))):
        assert np.array_equal(ctx.eval_expr(sysd, str(a), u), O.eval_expr(b, O.varnames(O.SYS_DICKE), u))
    # transcendental functions: a few ulp between CUDA and glibc
    got = ctx.eval_expr(sysd, "sin(Q)*cos(q)+exp(-P^2)+log(4-p^2)+atan(Q,P)+acos(Q/2)", u)
    want = np.sin(u[:, 0]) * np.cos(u[:, 1]) + np.exp(-u[:, 2] ** 2) + np.log(4 - u[:, 3] ** 2) + np.arctan2(u[:, 0], u[:, 2]) + np.arccos(u[:, 0] / 2)
    assert relerr(got, want).max() < 1e-13


def test_scale_to_index_bit_exact(ctx, O):
    rng = np.random.default_rng(7)
    v = np.concatenate([rng.uniform(-1.3, 1.3, 200000), np.linspace(-1.1, 1.1, 221), np.linspace(-1.1, 1.1, 441),
                        [np.nan, np.inf, -np.inf, -1.1, 1.1]])
    for (a, b, n) in ((-1.1, 1.1, 221), (-4.3, 4.3, 431), (0.0, 1.0, 1), (-0.8, 0.0, 41)):
        assert np.array_equal(ctx.scale_to_index(v, a, b, n), O.scale_to_index(v, a, b, n))


def test_scale_to_index_fast_path_never_disagrees_with_the_formula(ctx, O):
    """the kernels take rint(fma(v - start, c, 1)) unless that sits within 16 ulp of a rounding boundary, where the
    reference's division decides (csrc/dtwa_device.cuh scale_to_index): values AT the boundaries k + 1/2 and a few ulp
    either side of them, for short and very long axes, must give the reference's index (src/TWA.jl:53-64)"""
    rng = np.random.default_rng(11)
    for (a, b, n) in ((-1.1, 1.1, 221), (-4.3, 4.3, 431), (-4.3, 4.3, 8601), (0.0, 1.0, 2), (-7.25, 13.5, 100001), (1e-3, 2e-3, 1001),
                      (-2.0, 2.0, 2 ** 31 - 1), (0.5, 0.5, 7)):
        mids = a + (rng.integers(0, max(1, n - 1), 20000) + 0.5) * (b - a) / max(1, n - 1)
        vs = [rng.uniform(a - 0.1 * (b - a + 1), b + 0.1 * (b - a + 1), 50000), mids, np.array([a, b, np.nextafter(a, -np.inf), np.nextafter(b, np.inf)])]
        for s in range(1, 5):
            up, dn = mids.copy(), mids.copy()
            for _ in range(s):
                up, dn = np.nextafter(up, np.inf), np.nextafter(dn, -np.inf)
            vs += [up, dn]
        v = np.concatenate(vs)
        with np.errstate(all="ignore"):
            got, want = ctx.scale_to_index(v, a, b, n), O.scale_to_index(v, a, b, n)
            bad = np.nonzero(got != want)[0]
            assert len(bad) == 0, (a, b, n, [(float(v[i]).hex(), int(got[i]), int(want[i])) for i in bad[:5]])


def _random_expression(rng, depth):
    """a random arithmetic expression over Q,q,P,p with + - * / integer powers and sqrt(abs-free positive arguments)"""
    if depth == 0 or rng.random() < 0.2:
        r = rng.random()
        if r < 0.6:
            return str(rng.choice(["Q", "q", "P", "p"]))
        if r < 0.8:
            return repr(float(np.round(rng.uniform(0.1, 3.0), 3)))
        return str(int(rng.integers(1, 6)))
    op = rng.choice(["+", "-", "*", "/", "^", "sqrt", "neg"], p=[0.25, 0.2, 0.25, 0.1, 0.1, 0.05, 0.05])
    a = _random_expression(rng, depth - 1)
    if op == "sqrt":
        return f"sqrt(({a})^2 + 1.5)"
    if op == "neg":
        return f"(-({a}))"
    if op == "^":
        return f"({a})^{int(rng.integers(2, 6))}"
    b = _random_expression(rng, depth - 1)
    if op == "/":
        return f"({a})/(({b})^2 + 0.5)"
    return f"({a}) {op} ({b})"


def test_random_expressions_bit_exact_in_the_vm_and_when_specialised(ctx, pkg, O):
    """200 random expression trees: the interpreter (dtwa_eval_expr), the NVRTC-specialised reducer and the oracle's
    evaluator execute the same single-rounding operations -> identical bits"""
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    rng = np.random.default_rng(2026)
    u = np.column_stack([rng.uniform(-1.3, 1.3, 64), rng.normal(0, 1, 64), rng.uniform(-1.3, 1.3, 64), rng.normal(0, 1, 64)])
    sysd = csys(pkg, 0, DPAR)
    names = O.varnames(O.SYS_DICKE)
    exprs = [_random_expression(rng, 4) for _ in range(200)]
    for e in exprs:
        got = ctx.eval_expr(sysd, e, u)
        want = O.eval_expr(e, names, u)
        assert np.array_equal(got, want, equal_nan=True), e
    # through the ensemble: one trajectory per call would be slow -- average 8 observables at a time over the 64 states
    # at ts = [0] (sums of identical terms in a fixed order: interpreter == specialised kernel bit for bit)
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    try:
        for i in range(0, 24, 8):
            outs = []
            for jit in (0, 1):
                ctx.set_jit(jit)
                outs.append(TWA.average(system, observable=exprs[i:i + 8], N=64, ts=[0.0], distribution=TWA.samples_from_array(u, j=10)))
            assert np.array_equal(np.asarray(outs[0]), np.asarray(outs[1]), equal_nan=True), exprs[i:i + 8]
            ref = np.array([O.eval_expr(e, names, u) for e in exprs[i:i + 8]]).mean(axis=1)
            assert np.allclose(np.asarray(outs[0]).ravel(), ref, rtol=1e-12, atol=1e-12, equal_nan=True)
    finally:
        ctx.set_jit(2)
