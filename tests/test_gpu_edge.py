"""Edge cases of the C ABI and size-independent properties at BASELINE.json's full sizes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DPAR = [1.0, 1.0, 1.0]
DOCS_PT = [1.0, 0.31783724519578205, 1.0, 0.0]


@pytest.fixture(scope="module")
def mods(pkg):
    from dickemodel_jl_b200 import TWA, ClassicalDicke, ClassicalLMG, ClassicalSystems
    return TWA, ClassicalDicke, ClassicalLMG, ClassicalSystems


@pytest.fixture(scope="module")
def system(mods):
    return mods[1].ClassicalDickeSystem(w0=1, w=1, g=1)


def test_ragged_and_tiny_inputs(ctx, mods, system, O):
    TWA, CD, _, CS = mods
    j = 50
    # N = 1, one output time, non-uniform ts with repeated and late-starting times
    for N, ts in ((1, [0.0]), (1, [3.0]), (7, [0.5, 0.5, 0.75, 2.0, 2.0]), (257, [0.0, 0.0, 1e-9, 0.3]), (513, [1.0, 7.0])):
        u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / j, N, seed=N)
        got = TWA.average(system, observable=["Q", TWA.Weyl.Jz(j)], N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j),
                          integrator_alg="RK4", dt=0.01)
        traj, st, _, _ = O.integrate(O.SYS_DICKE, DPAR, u0, ts, alg=O.ALG_RK4, dt=0.01)
        want = np.stack([traj[:, :, 0].mean(axis=0), O.eval_expr(str(TWA.Weyl.Jz(j)), O.varnames(0), traj).mean(axis=0)], axis=1)
        assert got.shape == (len(ts), 2) and np.abs(got - want).max() < 1e-10, (N, ts)
        assert (TWA.last_info["n_valid"] == N).all()
    # adaptive integrators with the same ragged time axis
    u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / j, 100, seed=3)
    ts = [0.0, 0.0, 0.2, 0.2, 1.7]
    for alg, oalg in (("DP8", O.ALG_DOP853), ("DP5", O.ALG_DP5)):
        got = TWA.average(system, observable="q", N=100, ts=ts, distribution=TWA.samples_from_array(u0, j=j), integrator_alg=alg, tol=1e-10)
        traj, _, _, _ = O.integrate(O.SYS_DICKE, DPAR, u0, ts, alg=oalg, tol=1e-10)
        assert np.abs(got - traj[:, :, 1].mean(axis=0)).max() < 1e-8


def test_histogram_range_edges(ctx, mods, system, O):
    TWA = mods[0]
    u0 = np.array([[0.0, -1.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0], [0.0, 0.999999, 0.0, 0.0], [0.0, 1.0000001, 0.0, 0.0],
                   [0.0, 0.05, 0.0, 0.0], [0.0, 0.15, 0.0, 0.0]])
    D = TWA.samples_from_array(u0, j=10)
    m = TWA.calculate_distribution(system, x="q", xs=TWA.jrange(-1, 0.1, 1), N=len(u0), distribution=D)   # t = 0 only
    c = np.rint(m[0] * len(u0)).astype(int)
    # both ends inclusive, value just above stop dropped, ties to even: 0.05 -> index 11.5 -> 12 (1-based), 0.15 -> 12.5 -> 12
    assert c[0] == 1 and c[20] == 2 and c.sum() == 6 and c[10] == 1 and c[11] == 2
    D.offset = 0
    one = TWA.calculate_distribution(system, x="q", xs=(0.0, 2.0, 1), N=len(u0), distribution=D)             # len == 1: every value in range -> bin 1
    assert one.shape == (1, 1) and np.isclose(one[0, 0] * len(u0), 6)


def test_invalid_ensemble_arguments(ctx, mods, system, pkg):
    TWA = mods[0]
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=10, seed=1)
    with pytest.raises(pkg.DtwaError, match="unknown variable 'z'"):
        TWA.average(system, observable="Q+z", N=10, ts=[0.0], distribution=W)
    with pytest.raises(pkg.DtwaError, match="sorted"):
        TWA.average(system, observable="Q", N=10, ts=[1.0, 0.5], distribution=W, integrator_alg="RK4", dt=0.1)
    with pytest.raises(ValueError, match="N must be"):
        TWA.average(system, observable="Q", N=0, ts=[0.0], distribution=W)
    with pytest.raises(pkg.DtwaError, match="one histogram axis"):
        TWA.monte_carlo_integrate(system, TWA.MonteCarloSystem(None, [{"kind": pkg._lib.RED_HIST, "x_axis": 1, "y_axis": 1}], "+", lambda r: r, 10, [0.0], W))
    with pytest.raises(ValueError, match="samples 2 variables"):
        TWA.average(system, observable="Q", N=10, ts=[0.0], distribution=TWA.coherent_Wigner_SU2([0, 0], j=10))
    with pytest.raises(pkg.DtwaError, match="VM limits"):
        TWA.average(system, observable=["Q"] * 65, N=10, ts=[0.0], distribution=W)
    deep = "Q"
    for _ in range(20):
        deep = f"(q*(p+({deep})))"
    TWA.average(system, observable=deep, N=10, ts=[0.0], distribution=W)          # right-nested: constant stack depth, fine
    wide = "+".join(f"(Q*q+P*p)*(Q*p+{i})" for i in range(3))
    TWA.average(system, observable=wide, N=10, ts=[0.0], distribution=W)
    bad = "Q"
    for _ in range(18):
        bad = f"((q+p)*({bad}*(P+q)))"
    bad = "(" + ")+(".join(["(Q+q)*(P+p)"] * 2) + ")"
    TWA.average(system, observable=bad, N=10, ts=[0.0], distribution=W)


def test_sixty_four_observables_and_many_histograms(ctx, mods, system, O):
    TWA = mods[0]
    N, j = 2000, 100
    u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / j, N, seed=9)
    ts = TWA.jrange(0, 0.5, 2)
    obs = [f"Q^{1 + i % 4} + {i}*q - p/{i + 1}" for i in range(64)]
    for mode in (0, 1):
        ctx.set_jit(mode)
        got = TWA.average(system, observable=obs, N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), integrator_alg="RK4", dt=0.01)
        traj, _, _ = mods[3].integrate_batch(system, u0, np.asarray(ts), integrator_alg="RK4", dt=0.01)
        for i in (0, 17, 63):
            want = O.average_sums(O.eval_expr(obs[i], O.varnames(0), traj), np.ones(traj.shape[:2], bool)) / N
            assert np.abs(got[:, i] - want).max() < 1e-11
    ctx.set_jit(2)
    # 16 time-resolved histograms: only some fit the shared-memory staging budget, the rest go straight to global atomics
    D = TWA.samples_from_array(u0, j=j)
    chain = TWA.mcs_chain([TWA.mcs_for_distributions(system, y="t", x="q", xs=(-4.0, 4.0, 1000 + i), N=N, ts=ts, distribution=D) for i in range(16)])
    out = TWA.monte_carlo_integrate(system, chain, integrator_alg="RK4", dt=0.01)
    for i, m in enumerate(out):
        want = O.hist_time(traj[:, :, 1], np.ones(traj.shape[:2], bool), -4.0, 4.0, 1000 + i)
        assert np.array_equal(m, want / N)


def test_maxiters_and_no_tolerance(ctx, mods, system, pkg, O):
    TWA = mods[0]
    W = TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=20, seed=2)
    TWA.average(system, observable="Q", N=500, ts=[0.0, 5.0, 10.0], distribution=W, integrator_alg="DP8", tol=1e-10, maxiters=20)
    assert TWA.last_info["n_unfinished"] == 500 and list(TWA.last_info["n_valid"]) == [500, 0, 0]     # src/ClassicalSystems.jl:138 maxiters: no exception
    with pytest.raises(pkg.DtwaError, match="tolerate_errors=false"):
        TWA.average(system, observable="Q", N=20000, ts=TWA.jrange(0, 0.5, 20), distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=20, seed=2),
                    integrator_alg="RK4", dt=2.0 ** -5, tolerate_errors=False)


def test_max_nbatch_chunks_equal_one_launch(ctx, mods, system):
    TWA = mods[0]
    kw = dict(observable="q", N=10000, ts=TWA.jrange(0, 0.5, 2), integrator_alg="RK4", dt=0.01)
    a = TWA.average(system, distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=6), **kw)
    b = TWA.average(system, distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=6), maxNBatch=3000, **kw)
    assert np.abs(a - b).max() < 1e-13
    h1 = TWA.calculate_distribution(system, x="t", y="q", ys=TWA.jrange(-3, 0.1, 3), N=10000, ts=TWA.jrange(0, 0.5, 2),
                                    distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=6), integrator_alg="RK4", dt=0.01)
    h2 = TWA.calculate_distribution(system, x="t", y="q", ys=TWA.jrange(-3, 0.1, 3), N=10000, ts=TWA.jrange(0, 0.5, 2), maxNBatch=1234,
                                    distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=6), integrator_alg="RK4", dt=0.01)
    assert np.array_equal(h1, h2)      # integer counts: exactly additive over index blocks


def test_full_size_properties_config2(ctx, mods, system):
    """BASELINE.json configs[1] at full size (10^7 trajectories, ts = 0:0.05:100): size-independent checks"""
    TWA = mods[0]
    j, N = 100, 10_000_000
    ts = TWA.jrange(0, 0.05, 100)
    jz = TWA.Weyl.Jz(j) / j
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=j, seed=20261019)
    chain = TWA.mcs_chain(TWA.mcs_for_averaging(system, observable=[jz, 2 * jz + 3, "1"], N=N, ts=ts, distribution=W),
                          TWA.mcs_for_distributions(system, x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1), N=N, ts=ts, distribution=W))
    avg, hist = TWA.monte_carlo_integrate(system, chain, integrator_alg="RK4", dt=2.0 ** -7)
    info = TWA.last_info
    assert info["steps_accepted"] == N * 14000 and info["n_failed"] == 0 and (info["n_valid"] == N).all()
    assert np.array_equal(avg[:, 2], np.ones(len(ts)))                         # sum of N ones / N, exactly
    assert np.abs(avg[:, 1] - (2 * avg[:, 0] + 3)).max() < 1e-12               # linearity of the reduction
    assert np.allclose(hist.sum(axis=0), 1.0, rtol=0, atol=1e-12)              # every sample lands in -1.1:0.01:1.1
    centers = np.asarray(TWA.jrange(-1.1, 0.01, 1.1))
    assert np.abs((hist * centers[:, None]).sum(axis=0) - avg[:, 0]).max() < 0.01 / 2 + 1e-9    # histogram mean vs direct mean: within half a bin
    assert abs(avg[0, 0] + 1.0) < 0.02 and -0.30 < avg[100, 0] < -0.20 and -0.40 < avg[-1, 0] < -0.34
    # halves with offset Philox counters add up to the whole (what the multi-GPU shard does)
    parts = []
    for off in (0, N // 2):
        Wp = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=j, seed=20261019)
        Wp.offset = off
        parts.append(TWA.calculate_distribution(system, x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1), N=N // 2, ts=ts, distribution=Wp,
                                                integrator_alg="RK4", dt=2.0 ** -7))
    assert np.array_equal(np.rint((parts[0] + parts[1]) * (N // 2)), np.rint(hist * N))


def test_background_specialisation_takes_over_without_changing_a_bit(pkg, ctx):
    """small calls never wait for NVRTC: the first one runs the interpreter kernel and starts a background compilation,
    a later one picks the specialised kernel up; both produce the same bits"""
    import time
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    ctx.set_jit(2)
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    ts = TWA.jrange(0, 0.5, 4)
    obs = "Q*p - 0.37*q^2 + P"          # a reducer set no other test uses
    runs = []
    for attempt in range(120):   # up to 60 s for NVRTC on a busy box (normally 2-3 s)
        W = TWA.coherent_Wigner_HWxSU2([0.3, 0.1, -0.2, 0.4], j=50, seed=77)
        out = TWA.average(system, observable=obs, N=3000, ts=ts, distribution=W, tol=1e-9)
        runs.append((TWA.last_info["jit"], out.copy()))
        if runs[-1][0] == 1:
            break
        time.sleep(0.5)
    assert runs[0][0] == 0, "the first call must not wait for the compiler"
    assert runs[-1][0] == 1, "the specialised kernel never became available"
    assert all(np.array_equal(r[1], runs[0][1]) for r in runs)


def test_function_from_expression_and_scale_to_index_helpers(pkg, O):
    """src/TWA.jl:277-303 and :53-64 as callable helpers (evaluated on the GPU)"""
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    f = TWA.function_from_expression(system, ":(q+p^2-Q)")
    assert f(0.5, 1.0, 0.25, 2.0) == 1.0 + 4.0 - 0.5 and f([0.5, 1.0, 0.25, 2.0]) == 4.5
    g1, g2 = TWA.functions_from_expressions(system, [lambda Q, q, P, p: Q * P, TWA.Weyl.Jz(10)])
    u = np.random.default_rng(2).uniform(-1, 1, (7, 4))
    assert np.array_equal(g1(u), u[:, 0] * u[:, 2])
    assert np.array_equal(g2(u), O.eval_expr(str(TWA.Weyl.Jz(10)), O.varnames(O.SYS_DICKE), u))
    r = TWA.jrange(-1.1, 0.01, 1.1)
    assert TWA.scale_to_index(-1.1, r) == 1 and TWA.scale_to_index(1.1, r) == 221 and np.isnan(TWA.scale_to_index(1.2, r))
    v = np.linspace(-1.2, 1.2, 97)
    want = O.scale_to_index(v, -1.1, 1.1, 221).astype(float)
    want[want == 0] = np.nan
    assert np.array_equal(TWA.scale_to_index(v, r), want, equal_nan=True)


def test_full_size_properties_config3_share(ctx, mods, system):
    """BASELINE.json configs[2]: the jz-vs-t and q-vs-t histograms, one GPU's share (1.25e8) of 10^9 trajectories"""
    TWA, CD = mods[0], mods[1]
    j, N = 100, 125_000_000
    ts = TWA.jrange(0, 0.05, 40)
    jz = TWA.Weyl.Jz(j) / j
    B = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)
    W = TWA.coherent_Wigner_HWxSU2(B, j=j, seed=1)
    mk = lambda **k: TWA.mcs_for_distributions(system, N=N, distribution=W, ts=ts, **k)
    hj, hq = TWA.monte_carlo_integrate(system, TWA.mcs_chain(mk(x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1)),
                                                             mk(y="t", x="q", xs=TWA.jrange(-4.3, 0.02, 4.3))),
                                       integrator_alg="RK4", dt=2.0 ** -7)
    info = TWA.last_info
    assert hj.shape == (221, 801) and hq.shape == (801, 431)
    nv = info["n_valid"].astype(np.float64)
    cj, cq = np.rint(hj * N), np.rint(hq * N)
    assert np.abs(hj * N - cj).max() < 1e-6 and np.abs(hq * N - cq).max() < 1e-6        # integer counts
    assert np.array_equal(cj.sum(axis=0), nv)                                           # Jz/j always lies in -1.1:0.01:1.1
    assert (cq.sum(axis=1) <= nv).all() and cq.sum(axis=1).min() > 0.99 * N             # a few q fall outside +-4.3
    # failed trajectories were resampled (src/TWA.jl:186-209): every time has N contributions or a few more
    assert info["n_unfinished"] == 0 and (nv >= N).all() and (nv <= N + 2 * info["n_failed"]).all()


def test_full_size_properties_config4_share_and_config5(ctx, mods, system):
    """configs[3] (adaptive, chaotic, t <= 1000; one GPU's share 10^6) and configs[4] (LMG, 10^8): the ensemble mean of the
    Hamiltonian is a constant of motion; sums are linear"""
    TWA, CD, CL = mods[0], mods[1], mods[2]
    j = 100
    B = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)
    h = TWA._expr_of(system, CD.hamiltonian(system))
    out = TWA.average(system, observable=[h, "1", "2*(" + h + ") - 1"], N=1_000_000, ts=TWA.jrange(0, 1, 1000),
                      distribution=TWA.coherent_Wigner_HWxSU2(B, j=j, seed=1), integrator_alg="DP8", tol=1e-8)
    info = TWA.last_info
    out = np.asarray(out)
    assert info["n_unfinished"] == 0 and np.array_equal(out[:, 1], info["n_valid"] / 1_000_000)
    assert np.abs(out[:, 2] - (2 * out[:, 0] - out[:, 1])).max() < 1e-12
    mean_h = out[:, 0] / out[:, 1]
    assert np.abs(mean_h - mean_h[0]).max() < 2e-5          # tol = 1e-8 over t = 1000; resampled replacements shift it at 1e-6
    lmg = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
    hl = TWA._expr_of(lmg, CL.hamiltonian(lmg))
    N5 = 100_000_000
    v = TWA.average(lmg, observable=[hl, "Q", "Q^2", "1"], N=N5, ts=TWA.jrange(0, 0.1, 50),
                    distribution=TWA.coherent_Wigner_SU2([0, 0], j=500, seed=1), integrator_alg="RK4", dt=2.0 ** -7)
    v = np.asarray(v)
    i5 = TWA.last_info
    assert i5["n_failed"] == 0 and (i5["n_valid"] == N5).all() and i5["steps_accepted"] == N5 * 13 * 500
    assert np.array_equal(v[:, 3], np.ones(v.shape[0]))
    assert np.abs(v[:, 0] - v[0, 0]).max() < 1e-9           # <h> conserved (RK4 at dt = 2^-7 drifts by 1.5e-10 over t = 50)
    assert np.abs(v[:, 1]).max() < 5e-4 and (v[:, 2] > 0).all()          # the cloud stays centred; variance of Q positive


@pytest.mark.gpu
def test_animate_request_that_does_not_fit_is_refused(ctx):
    """src/TWA.jl:441,480-541 keeps per-time sparse dictionaries and flushes them; the library keeps one dense ny x nx u64
    matrix per output time on the device, so a request beyond the free device memory is refused up front (no partial
    allocation, no launch) with DTWA_ERR_CAPACITY and a message that says what to change"""
    from dickemodel_jl_b200 import TWA, ClassicalDicke as CD
    import dickemodel_jl_b200 as pkg
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=100, seed=1)
    ts = TWA.jrange(0, 0.05, 40)                                    # 801 times
    big = dict(x="q", xs=TWA.jrange(-4.3, 0.001, 4.3), y="p", ys=TWA.jrange(-4.3, 0.001, 4.3))   # 8601 x 8601 bins
    with pytest.raises(pkg.DtwaError, match="device memory") as e:
        TWA.calculate_distribution(system, N=1000, ts=ts, distribution=W, animate=True, integrator_alg="RK4", dt=2.0 ** -7, **big)
    assert e.value.code == pkg._lib.ERR_CAPACITY
    # the same bins without animate are one matrix (592 MB): accepted
    m = TWA.calculate_distribution(system, N=1000, ts=TWA.jrange(0, 0.5, 1), distribution=W, integrator_alg="RK4", dt=2.0 ** -7, **big)
    assert m.shape == (8601, 8601) and abs(m.sum() - 3.0) < 1e-9
    # and the context is still usable after the refusal
    small = TWA.calculate_distribution(system, N=1000, ts=TWA.jrange(0, 0.5, 1), distribution=W, animate=True, integrator_alg="RK4", dt=2.0 ** -7,
                                       x="q", xs=TWA.jrange(-4.3, 0.1, 4.3), y="p", ys=TWA.jrange(-4.3, 0.1, 4.3))
    assert len(small) == 3 and all(abs(mk.sum() - 1.0) < 1e-9 for mk in small)
