"""The fused ensemble kernel through the reference-shaped Python API, against the oracle."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

DPAR = [1.0, 1.0, 1.0]
DOCS_PT = [1.0, 0.31783724519578205, 1.0, 0.0]


@pytest.fixture(scope="module")
def mods(pkg):
    from dickemodel_jl_b200 import TWA, ClassicalDicke, ClassicalLMG, ClassicalSystems
    return TWA, ClassicalDicke, ClassicalLMG, ClassicalSystems


def test_average_matches_oracle_on_host_supplied_ics(ctx, mods, O):
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 100, 3000
    u0 = O.sample(O.SAMPLER_HWXSU2, [0, 0, 0, 0], 1 / j, N, seed=21)
    ts = TWA.jrange(0, 0.05, 10)
    obs = [TWA.Weyl.Jz(j) / j, "q", "p^2"]
    got = TWA.average(system, observable=obs, N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), integrator_alg="RK4", dt=2.0 ** -7)
    traj, st, acc, _ = O.integrate(O.SYS_DICKE, DPAR, u0, np.asarray(ts), alg=O.ALG_RK4, dt=2.0 ** -7)
    assert (st == 0).all()
    valid = np.isfinite(traj).all(axis=2)
    for i, e in enumerate(obs):
        want = O.average_sums(O.eval_expr(str(e), O.varnames(0), traj), valid) / N
        assert np.abs(got[:, i] - want).max() < 1e-10, e
    assert TWA.last_info["steps_accepted"] == int(acc.sum())
    assert (TWA.last_info["n_valid"] == N).all() and TWA.last_info["n_failed"] == 0
    # scalar observable -> vector of scalars (src/TWA.jl:363-365)
    one = TWA.average(system, observable=obs[0], N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), integrator_alg="RK4", dt=2.0 ** -7)
    assert one.shape == (len(ts),) and np.array_equal(one, got[:, 0])


def test_overlapped_upload_of_host_initial_conditions_is_bit_identical(ctx, mods, monkeypatch):
    """Large host-supplied ensembles under a fixed-step integrator are uploaded while the first batch of every CTA is already
    being integrated (two launches, dtwa_ensemble_launch); every trajectory still meets the same thread in the same order, so
    sums, counts and counters equal those of the single-launch path bit for bit -- also when N is not a multiple of anything"""
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=0.9, g=1.1)
    j, N = 80, 8 * 148 * 4 * 256 + 12345      # >= 8 first passes (148 SMs x 4 CTAs x 256 threads)
    W = TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=j, seed=77)
    u0 = W.sample_many(N)
    ts = TWA.jrange(0, 0.05, 0.5)

    def run():
        avg = TWA.average(system, observable=[TWA.Weyl.Jz(j) / j, "q*p"], N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), integrator_alg="RK4", dt=0.01)
        ia = dict(TWA.last_info)
        D = TWA.samples_from_array(u0, j=j)
        h = TWA.monte_carlo_integrate(system, TWA.mcs_chain(TWA.mcs_for_distributions(system, N=N, distribution=D, ts=ts, x="t", y="q", ys=TWA.jrange(-3, 0.1, 3)),
                                                            TWA.mcs_for_distributions(system, N=N, distribution=D, ts=ts, x="Q", xs=TWA.jrange(-2, 0.1, 2), y="P", ys=TWA.jrange(-2, 0.1, 2))),
                                      integrator_alg="RK4", dt=0.01)
        return avg, ia, h

    a1, i1, h1 = run()
    monkeypatch.setenv("DTWA_NO_COPY_OVERLAP", "1")
    a0, i0, h0 = run()
    assert i1["launches"] == i0["launches"] + 1            # the split launch really ran
    assert np.array_equal(a1, a0)
    assert all(np.array_equal(x, y) for x, y in zip(h1, h0))
    assert i1["steps_accepted"] == i0["steps_accepted"] and i1["n_failed"] == i0["n_failed"] and np.array_equal(i1["n_valid"], i0["n_valid"])


def test_runs_are_bitwise_reproducible(ctx, mods):
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    a = TWA.average(system, observable=TWA.Weyl.Jz(50), N=20000, ts=TWA.jrange(0, 0.5, 5), distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=5), integrator_alg="RK4", dt=0.01)
    b = TWA.average(system, observable=TWA.Weyl.Jz(50), N=20000, ts=TWA.jrange(0, 0.5, 5), distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=5), integrator_alg="RK4", dt=0.01)
    assert np.array_equal(a, b)


def test_histograms_bit_exact_given_identical_trajectories(ctx, mods, O, pkg):
    """counts from the fused kernel == oracle binning of the states dtwa_integrate reports for the
    same initial conditions (both kernels run the same stepper instruction sequence)"""
    TWA, CD, _, CS = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 100, 4000
    u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / j, N, seed=33)
    ts = TWA.jrange(0, 0.1, 8)
    jz = TWA.Weyl.Jz(j) / j
    ys, qs, ps = TWA.jrange(-1.1, 0.01, 1.1), TWA.jrange(-4.3, 0.02, 4.3), TWA.jrange(-2.4, 0.02, 2.4)
    mk = lambda **kw: TWA.mcs_for_distributions(system, N=N, distribution=dist, ts=ts, **kw)
    dist = TWA.samples_from_array(u0, j=j)
    chain = TWA.mcs_chain(mk(x="t", y=jz, ys=ys), mk(y="t", x="q", xs=qs), mk(x="q", xs=qs, y="p", ys=ps),
                          mk(x="q", xs=qs, y="p", ys=ps, animate=True))
    m_jz, m_q, m_qp, m_anim = TWA.monte_carlo_integrate(system, chain, integrator_alg="RK4", dt=2.0 ** -7)
    traj, st, _ = CS.integrate_batch(system, u0, np.asarray(ts), integrator_alg="RK4", dt=2.0 ** -7)
    valid = (st == 0)[:, None] & np.isfinite(traj).all(axis=2)
    vjz = O.eval_expr(str(jz), O.varnames(0), traj)
    c_jz = O.hist_time(vjz, valid, *ys.linrange())
    c_q = O.hist_time(traj[:, :, 1], valid, *qs.linrange())
    c_qp = O.hist_2d(traj[:, :, 1], traj[:, :, 3], valid, qs.linrange(), ps.linrange(), animate=False)
    c_an = O.hist_2d(traj[:, :, 1], traj[:, :, 3], valid, qs.linrange(), ps.linrange(), animate=True)
    assert m_jz.shape == (len(ys), len(ts)) and np.array_equal(m_jz, c_jz.T / N)       # [iy, i_t]  (src/TWA.jl:459)
    assert m_q.shape == (len(ts), len(qs)) and np.array_equal(m_q, c_q / N)            # [i_t, ix]  (:466)
    assert m_qp.shape == (len(ps), len(qs)) and np.array_equal(m_qp, c_qp / N)         # [iy, ix]   (:474)
    assert len(m_anim) == len(ts) and all(np.array_equal(m, c / N) for m, c in zip(m_anim, c_an))
    assert c_jz.sum() == N * len(ts)   # jz in [-1,1] always lands in -1.1:0.01:1.1


def test_no_time_evolution_path_bit_exact(ctx, mods, O):
    """ts == [0.0]: sample -> observable -> bin, no integration (src/TWA.jl:177,187-188); the LDoS
    example passes the Hamiltonian closure as observable (docs/src/TWAExamples.md:369-376)"""
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 1000, 50000
    x = CD.Point(system, Q=-1, P=0, p=0, ϵ=-0.5)
    u0 = O.sample(O.SAMPLER_HWXSU2, x, 1 / j, N, seed=8)
    es = TWA.jrange(-0.8, 0.02, 0)
    rho = TWA.calculate_distribution(system, distribution=TWA.samples_from_array(u0, j=j), N=N, x=CD.hamiltonian(system), xs=es)
    assert rho.shape == (1, len(es))
    h = O.eval_expr(TWA._expr_of(system, CD.hamiltonian(system)), O.varnames(0), u0)
    want = O.hist_time(h[:, None], np.ones((N, 1), bool), *es.linrange())
    assert np.array_equal(rho, want / N)
    assert abs((rho[0] * np.asarray(es)).sum() / rho.sum() + 0.5) < 0.01      # centred on the state's energy


def test_variance_and_fotoc_initial_value(ctx, mods, O):
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 1000, 40000
    x = CD.Point(system, Q=1, P=0, p=0, ϵ=-0.6)
    W = TWA.coherent_Wigner_HWxSU2(x, j=j, seed=77)
    ts = TWA.jrange(0, 0.5, 3)
    var, mean = TWA.variance(system, observable=["Q", "q", "P", "p"], distribution=W, N=N, ts=ts, tol=1e-8, return_average=True)
    fotoc = var.sum(axis=1)
    assert abs(fotoc[0] - 2 / j) < 0.05 * 2 / j          # docs/src/TWAExamples.md:334
    assert np.abs(mean[0] - x).max() < 5 * np.sqrt(1 / (2 * j) / N) + 1e-3
    # against the oracle on the same Philox samples (adaptive DOP853 at the docs' tol)
    u0 = O.sample(O.SAMPLER_HWXSU2, x, 1 / j, N, seed=77)
    traj, st, _, _ = O.integrate(O.SYS_DICKE, DPAR, u0, np.asarray(ts), alg=O.ALG_DOP853, tol=1e-8)
    assert relerr(var, traj.var(axis=0), floor=1e-6).max() < 1e-5
    one = TWA.variance(system, observable="Q", distribution=TWA.coherent_Wigner_HWxSU2(x, j=j, seed=77), N=N, ts=ts, tol=1e-8)
    assert one.shape == (len(ts),) and np.allclose(one, var[:, 0], rtol=1e-12)


def test_survival_probability(ctx, mods, O):
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 30, 20000
    x = CD.Point(system, Q=1.75, P=0, p=0, ϵ=-0.5)
    ts = np.concatenate([[0.0], 10.0 ** np.arange(-2, 0.51, 0.25)])
    sp = TWA.survival_probability(system, distribution=TWA.coherent_Wigner_HWxSU2(x, j=j, seed=3), N=N, ts=ts, tol=1e-9)
    # t = 0: (2 pi hbar)^2 E_W[W] = (2 pi hbar)^2 int W^2 = 1 for Gaussians of variance hbar/2 per coordinate
    assert abs(sp[0] - 1.0) < 0.03
    u0 = O.sample(O.SAMPLER_HWXSU2, x, 1 / j, N, seed=3)
    traj, st, _, _ = O.integrate(O.SYS_DICKE, DPAR, u0, ts, alg=O.ALG_DOP853, tol=1e-9, backwards=True)
    w = O.pdf(O.SAMPLER_HWXSU2, x, 1 / j, traj.reshape(-1, 4)).reshape(N, len(ts))
    want = w.sum(axis=0) * (2 * np.pi / j) ** 2 / N
    assert relerr(sp, want, floor=1e-4).max() < 1e-6


def test_philox_ensemble_matches_oracle_ensemble(ctx, mods, O):
    """fused Philox sampling: same seed => same sample set as the oracle => same sums up to rounding"""
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 100, 5000
    ts = TWA.jrange(0, 0.25, 5)
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=j, seed=424242)
    got = TWA.average(system, observable=TWA.Weyl.Jz(j) / j, N=N, ts=ts, distribution=W, integrator_alg="RK4", dt=2.0 ** -7)
    sp = O.make_sampler(O.SAMPLER_HWXSU2, [0, 0, 0, 0], 1 / j, seed=424242)
    c = float(np.sqrt(j * (j + 1)))
    ref = O.ensemble(O.SYS_DICKE, DPAR, sp, N, np.asarray(ts), values=[(2, 0, c)], alg=O.ALG_RK4, dt=2.0 ** -7, workers=4)
    assert np.abs(got - ref["sums"][:, 0] / j / N).max() < 1e-9
    assert W.offset == N      # the next ensemble draws fresh trajectories
    again = TWA.average(system, observable=TWA.Weyl.Jz(j) / j, N=N, ts=ts, distribution=W, integrator_alg="RK4", dt=2.0 ** -7)
    assert not np.array_equal(again, got) and np.abs(again - got).max() < 0.05


def test_resampling_of_failed_trajectories(ctx, mods, O):
    """tolerate_errors (src/TWA.jl:186-209): a broad (j=20) chaotic ensemble, ~5% of whose trajectories reach the chart
    singularity; replacements contribute from the failure index on, so every time has >= N samples"""
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    N = 20000
    W = TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=20, seed=2)
    ts = TWA.jrange(0, 0.5, 20)
    TWA.average(system, observable="Q", N=N, ts=ts, distribution=W, integrator_alg="RK4", dt=2.0 ** -5)
    info = TWA.last_info
    assert info["n_failed"] > 0 and info["resample_rounds"] >= 1 and info["n_unfinished"] == 0
    assert info["n_valid"].min() >= N
    sp = O.make_sampler(O.SAMPLER_HWXSU2, DOCS_PT, 1 / 20, seed=2)
    ref = O.ensemble(O.SYS_DICKE, DPAR, sp, N, np.asarray(ts), values=[(0, 0, 0.0)], alg=O.ALG_RK4, dt=2.0 ** -5, workers=1)
    # same trajectories, same replacement draws (attempt counter), same failure indices
    # (trajectories that graze the singularity amplify last-bit differences, so allow a few to differ)
    assert abs(info["n_failed"] - ref["n_failed"]) <= max(3, 0.03 * ref["n_failed"])
    assert np.abs(info["n_valid"].astype(np.int64) - ref["n_valid"].astype(np.int64)).max() <= max(3, 0.03 * ref["n_failed"])


def test_resampling_with_known_failing_indices(ctx, mods, O):
    """Deterministic twin of the test above: LMG (polynomial field, no singular passages) sampled from a planar Gaussian so
    broad (hbar = 0.5 around Q = 1.5) that many draws fall outside the disk Q^2 + P^2 < 4.  Those -- and only those -- fail
    (the reference's integrate raises "out of bounds", src/ClassicalSystems.jl:100-102, and monte_carlo_integrate redraws,
    src/TWA.jl:186-209), so the failing indices, the number of rounds and the final ensemble follow from the samples alone."""
    TWA, _, CL, _ = mods
    system = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
    N, hbar, seed = 6000, 0.5, 31
    ts = np.array([0.0, 0.5, 1.0, 1.5])
    W = TWA.coherent_Wigner_HW(q0=1.5, p0=0.0, ħ=hbar, seed=seed)
    got = TWA.average(system, observable=["Q", "P"], N=N, ts=ts, distribution=W, integrator_alg="RK4", dt=2.0 ** -6)
    info = dict(TWA.last_info)
    # replay the rounds on the host from the library's own sampler hook (dtwa_sample with the attempt counter)
    final = np.full((N, 2), np.nan)
    pending, n_failed, rounds = np.arange(N), 0, 0
    for attempt in range(50):
        u = ctx.sample(TWA.coherent_Wigner_HW(q0=1.5, p0=0.0, ħ=hbar, seed=seed)._c_sampler(), N, attempt)
        ok = 4 - u[pending, 1] ** 2 - u[pending, 0] ** 2 > 0
        final[pending[ok]] = u[pending[ok]]
        n_failed += int((~ok).sum())
        pending = pending[~ok]
        if len(pending) == 0:
            break
        rounds += 1
    assert n_failed > 0.2 * N and rounds >= 3                      # the test does exercise several rounds
    assert info["n_failed"] == n_failed and info["resample_rounds"] == rounds and info["n_unfinished"] == 0
    assert (info["n_valid"] == N).all()                              # a redraw at k = 0 contributes at every time, exactly once
    traj, st, _, _ = O.integrate(O.SYS_LMG, [1.0, -1.0], final, ts, alg=O.ALG_RK4, dt=2.0 ** -6)
    assert (st == 0).all()
    assert np.abs(got - traj.mean(axis=0)).max() < 1e-11
    # and the oracle's own driver replays the same rounds
    ref = O.ensemble(O.SYS_LMG, [1.0, -1.0], O.make_sampler(O.SAMPLER_HW, [1.5, 0.0], hbar, seed=seed), N, ts, values=[(0, 0, 0.0), (0, 1, 0.0)],
                     alg=O.ALG_RK4, dt=2.0 ** -6)
    assert ref["n_failed"] == n_failed and np.abs(ref["sums"] / N - got).max() < 1e-11


def test_lmg_variance_example(ctx, mods, O):
    """docs/src/ClassicalLMGExamples.md:22-38"""
    TWA, _, CL, _ = mods
    system = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
    j, N = 500, 5000
    ts = TWA.jrange(0, 0.1, 50)
    W = TWA.coherent_Wigner_SU2([0, 0], j=j, seed=12)
    var = TWA.variance(system, observable=["Q", "P"], N=N, ts=ts, distribution=W, integrator_alg="RK4", dt=2.0 ** -7)
    u0 = O.sample(O.SAMPLER_SU2, [0, 0], 1 / j, N, seed=12)
    traj, st, _, _ = O.integrate(O.SYS_LMG, [1.0, -1.0], u0, np.asarray(ts), alg=O.ALG_RK4, dt=2.0 ** -7)
    assert (st == 0).all()
    assert relerr(var, traj.var(axis=0), floor=1e-9).max() < 1e-6
    assert abs(var[0].sum() - 1 / j) < 0.1 / j


def test_single_trajectory_integrate_api(ctx, mods, O):
    """docs/src/ClassicalDickeExamples.md:43-51"""
    _, CD, _, CS = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    u0 = CD.Point(system, Q=1, P=1, p=0, ϵ=0.5)
    sol = CS.integrate(system, u0, 10.0)
    ref, st, _, _ = O.integrate(O.SYS_DICKE, DPAR, [u0], [10.0], alg=O.ALG_DOP853, tol=1e-12)
    assert sol.retcode == "Success" and relerr(sol.u[-1], ref[0, 0]).max() < 1e-9
    assert relerr(sol(5.0), O.integrate(O.SYS_DICKE, DPAR, [u0], [5.0], alg=O.ALG_DOP853, tol=1e-12)[0][0, 0]).max() < 1e-9
    with pytest.raises(ValueError, match="out of bounds"):
        CS.integrate(system, [1.5, 0, 1.5, 0], 1.0)
    with pytest.raises(ValueError, match="negative time"):
        CS.integrate(system, u0, -1.0)
    with pytest.raises(ValueError, match="should have length 4"):
        CS.integrate(system, [0.0, 0.0], 1.0)


def test_specialised_kernel_is_bit_identical_to_the_interpreter(ctx, mods, O):
    """dtwa_set_jit: the NVRTC-specialised reducers execute the interpreter's operation sequence"""
    TWA, CD, CL, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 100, 6000
    ts = TWA.jrange(0, 0.1, 4)
    jz = TWA.Weyl.Jz(j) / j

    def run(mode, **solver):
        ctx.set_jit(mode)
        try:
            W = TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=j, seed=31)
            mk = lambda **kw: TWA.mcs_for_distributions(system, N=N, distribution=W, ts=ts, **kw)
            chain = TWA.mcs_chain(TWA.mcs_for_averaging(system, observable=[jz, TWA.Weyl.Jx2(j), "sin(q)*exp(-p^2)", "Q^5/(1+P^4)"], N=N, ts=ts, distribution=W),
                                  mk(x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1)), mk(x="q", xs=TWA.jrange(-4.3, 0.1, 4.3), y="p", ys=TWA.jrange(-2.4, 0.1, 2.4)),
                                  TWA.mcs_for_survival_probability(system, N=N, ts=ts, distribution=W))
            out = TWA.monte_carlo_integrate(system, chain, **solver)
            return out, TWA.last_info["jit"]
        finally:
            ctx.set_jit(2)

    for solver in (dict(integrator_alg="RK4", dt=2.0 ** -6), dict(integrator_alg="DP8", tol=1e-9), dict(integrator_alg="DP5", tol=1e-7)):
        a, ja = run(0, **solver)
        b, jb = run(1, **solver)
        assert (ja, jb) == (0, 1)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    # LMG through the same path
    lmg = CL.ClassicalLMGSystem(Ω=1, ξ=-1)
    res = []
    for mode in (0, 1):
        ctx.set_jit(mode)
        W = TWA.coherent_Wigner_SU2([0.2, 0.1], j=50, seed=4)
        res.append(TWA.variance(lmg, observable=["Q", "P", "Q*P"], N=3000, ts=TWA.jrange(0, 0.5, 5), distribution=W, integrator_alg="RK8", dt=0.05))
    ctx.set_jit(2)
    assert np.array_equal(res[0], res[1])


@pytest.mark.parametrize("solver", [dict(integrator_alg="RK4", dt=2.0 ** -6), dict(integrator_alg="DP8", tol=1e-9)])
def test_multi_device_context_matches_single_device(pkg, mods, O, solver):
    """dtwa_create(ndev=2): contiguous index blocks + one grouped ncclAllReduce; integer counts are
    bit-identical to the single-device run, sums agree to rounding"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 100, 50001
    ts = TWA.jrange(0, 0.1, 5)
    jz = TWA.Weyl.Jz(j) / j

    def run():
        W = TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=j, seed=5)
        chain = TWA.mcs_chain(TWA.mcs_for_averaging(system, observable=jz, N=N, ts=ts, distribution=W),
                              TWA.mcs_for_distributions(system, x="t", y=jz, ys=TWA.jrange(-1.1, 0.01, 1.1), N=N, ts=ts, distribution=W))
        out = TWA.monte_carlo_integrate(system, chain, **solver)
        return out, dict(TWA.last_info)

    one, i1 = run()
    old = pkg.default_context()
    try:
        pkg.set_default_context(pkg.Context([0, 1]))
        two, i2 = run()
    finally:
        pkg.set_default_context(old)
    assert np.array_equal(one[1], two[1])
    assert np.abs(one[0] - two[0]).max() < 1e-12
    assert i1["steps_accepted"] == i2["steps_accepted"] and np.array_equal(i1["n_valid"], i2["n_valid"])


def test_multi_device_context_with_large_host_supplied_arrays(pkg, mods):
    """dtwa_create(ndev=2) + host-supplied initial conditions large enough that every device overlaps its upload with its first
    batch (two launches per device, each on its own copy stream): counts bit-identical to one device, sums to rounding"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    TWA, CD, _, _ = mods
    system = CD.ClassicalDickeSystem(w0=1, w=0.9, g=1.1)
    j, N = 80, 2 * (8 * 148 * 4 * 256) + 4321
    u0 = TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=j, seed=78).sample_many(N)
    ts = TWA.jrange(0, 0.05, 0.5)

    def run():
        D = TWA.samples_from_array(u0, j=j)
        chain = TWA.mcs_chain(TWA.mcs_for_averaging(system, observable=["q*p", TWA.Weyl.Jz(j) / j], N=N, ts=ts, distribution=D),
                              TWA.mcs_for_distributions(system, x="t", y="q", ys=TWA.jrange(-3, 0.1, 3), N=N, ts=ts, distribution=D))
        out = TWA.monte_carlo_integrate(system, chain, integrator_alg="RK4", dt=0.01)
        return out, dict(TWA.last_info)

    one, i1 = run()
    old = pkg.default_context()
    try:
        pkg.set_default_context(pkg.Context([0, 1]))
        two, i2 = run()
    finally:
        pkg.set_default_context(old)
    assert i1["launches"] >= 3 and i2["launches"] >= 2 * 3 - 1     # split launches (+ the merge kernels) on one and on both devices
    assert np.array_equal(one[1], two[1])
    assert np.abs(one[0] - two[0]).max() < 1e-12
    assert i1["steps_accepted"] == i2["steps_accepted"] and np.array_equal(i1["n_valid"], i2["n_valid"])


def test_user_defined_phase_space_distribution(ctx, mods, O):
    """src/TWA.jl:22-26: a distribution given by user closures.  sample() feeds host-drawn points to the kernel; the traced
    probability_density is the survival-probability observable (src/TWA.jl:656-665)"""
    TWA, CD, _, _ = mods
    from dickemodel_jl_b200 import symbolic as S
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 40, 3000
    hb = 1 / j
    x0 = np.array(CD.Point(system, Q=1.0, P=0, p=0, ϵ=-0.6))
    pts = x0 + np.random.default_rng(5).normal(0, np.sqrt(hb / 2), (N, 4))
    it = iter(pts)
    pdf = lambda u: S.exp(-((u[0] - x0[0]) ** 2 + (u[1] - x0[1]) ** 2 + (u[2] - x0[2]) ** 2 + (u[3] - x0[3]) ** 2) / hb) / (np.pi * hb) ** 2   # noqa: E731
    D = TWA.PhaseSpaceDistribution(pdf, lambda: next(it), hb)
    ts = np.array([0.0, 0.5, 1.0, 2.0])
    got = TWA.average(system, observable=["Q", "q"], N=N, ts=ts, distribution=D, integrator_alg="RK4", dt=2.0 ** -6)
    same = TWA.average(system, observable=["Q", "q"], N=N, ts=ts, distribution=TWA.samples_from_array(pts, j=j), integrator_alg="RK4", dt=2.0 ** -6)
    assert np.array_equal(got, same)                     # the same points, the same kernel
    it = iter(pts)
    sp = TWA.survival_probability(system, distribution=D, N=N, ts=ts, integrator_alg="RK4", dt=2.0 ** -6)
    traj, st, _, _ = O.integrate(O.SYS_DICKE, DPAR, pts, ts, alg=O.ALG_RK4, dt=2.0 ** -6, backwards=True)
    assert (st == 0).all()
    w = np.exp(-((traj - x0) ** 2).sum(axis=2) / hb) / (np.pi * hb) ** 2
    want = w.sum(axis=0) * (2 * np.pi * hb) ** 2 / N
    assert relerr(sp, want, floor=1e-6).max() < 1e-9
    assert abs(sp[0] - 1.0) < 0.1                        # (2 pi hbar)^2 E_W[W] = 1 for this Gaussian


def test_small_2d_histogram_tile_in_shared_memory_is_bit_exact(ctx, mods, O):
    """2-D histograms whose matrix fits the per-CTA shared-memory tile (<= 8192 cells, all times in one matrix) are staged
    there and flushed with one global RED per non-empty cell; the counts must equal those of the straight-to-global path
    (DTWA_NO_STAGE2D=1) and the oracle's binning of the same trajectories"""
    import os
    TWA, CD, _, CS = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 60, 6000
    u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / j, N, seed=21)
    ts = TWA.jrange(0, 0.25, 6)
    qs, ps = TWA.jrange(-4.0, 0.1, 4.0), TWA.jrange(-4.0, 0.1, 4.0)     # 81 x 81 = 6561 cells
    run = lambda: TWA.calculate_distribution(system, N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), x="q", xs=qs, y="p", ys=ps,   # noqa: E731
                                             integrator_alg="RK4", dt=2.0 ** -6)
    staged = run()
    os.environ["DTWA_NO_STAGE2D"] = "1"
    try:
        direct = run()
    finally:
        del os.environ["DTWA_NO_STAGE2D"]
    assert np.array_equal(staged, direct)
    traj, st, _ = CS.integrate_batch(system, u0, np.asarray(ts), integrator_alg="RK4", dt=2.0 ** -6)
    valid = (st == 0)[:, None] & np.isfinite(traj).all(axis=2)
    want = O.hist_2d(traj[:, :, 1], traj[:, :, 3], valid, qs.linrange(), ps.linrange(), animate=False)
    assert np.array_equal(staged, want / N) and want.sum() > 0.99 * N * len(ts)
    # many batches per CTA and a chain of two tiles + a 1-D histogram: Philox sampling, larger N
    W = TWA.coherent_Wigner_HWxSU2([0, 0, 0, 0], j=100, seed=3)
    mk = lambda **kw: TWA.mcs_for_distributions(system, N=400000, distribution=W, ts=TWA.jrange(0, 0.5, 2), **kw)   # noqa: E731
    chain = lambda: TWA.monte_carlo_integrate(system, TWA.mcs_chain(mk(x="q", xs=TWA.jrange(-1, 0.05, 1), y="p", ys=TWA.jrange(-1, 0.05, 1)),   # noqa: E731
                                                                    mk(x="Q", xs=TWA.jrange(-1, 0.05, 1), y="P", ys=TWA.jrange(-1, 0.05, 1)),
                                                                    mk(y="t", x="q", xs=TWA.jrange(-1, 0.01, 1))), integrator_alg="RK4", dt=2.0 ** -5)
    a = chain()
    W.offset = 0
    os.environ["DTWA_NO_STAGE2D"] = "1"
    try:
        b = chain()
    finally:
        del os.environ["DTWA_NO_STAGE2D"]
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and 4.9 < a[0].sum() <= 5.0 + 1e-9    # (a few samples leave q in [-1, 1] by t = 2)


def test_bloch_chart_is_the_same_dynamics(ctx, mods, O, gold):
    """chart="bloch" (opt-in): Hamilton's equations in (jx, jy, jz, q, p), bilinear, no square root.  Coordinate-free checks:
    against the 40-digit mpmath truth at tight tolerance, against the reference's chart at the truncation-error level with
    the method's convergence order, and through the ensemble kernels (same Philox samples, same observables)"""
    TWA, CD, _, CS = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    for c in gold["truth_mpmath"]["cases"]:
        if c["system"] != "dicke":
            continue
        keep = [i for i, t in enumerate(c["times"]) if float(t) <= 10.0]
        ts = np.array([float(c["times"][i]) for i in keep])
        sysc = CD.ClassicalDickeSystem(w0=c["par"][0], w=c["par"][1], g=c["par"][2])
        out, st, _ = CS.integrate_batch(sysc, [c["u0"]], ts, integrator_alg="DP8", tol=1e-13, chart="bloch")
        assert st[0] == 0
        want = np.array([[float(x) for x in c["u"][i]] for i in keep])
        assert relerr(out[0], want).max() < 2e-10, relerr(out[0], want).max()
    u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / 50, 64, seed=4)
    ts = np.array([0.0, 1.0, 2.0])
    ref, st, _ = CS.integrate_batch(system, u0, ts, integrator_alg="DP8", tol=1e-13)
    errs = []
    for dt in (2.0 ** -5, 2.0 ** -6):
        b, stb, _ = CS.integrate_batch(system, u0, ts, integrator_alg="RK4", dt=dt, chart="bloch")
        assert (stb == 0).all()
        errs.append(np.abs(b - ref).max())
    assert errs[0] < 1e-5 and 10 < errs[0] / errs[1] < 24          # 4th order in the other chart as well
    # ensembles: same Philox samples, Bloch chart vs (Q,P) chart, adaptive at a tight tolerance -> same averages
    kw = dict(observable=[TWA.Weyl.Jz(50) / 50, "q", "Q*p"], N=4096, ts=TWA.jrange(0, 0.5, 3), integrator_alg="DP8", tol=1e-12)
    a = TWA.average(system, distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=9), **kw)
    b = TWA.average(system, distribution=TWA.coherent_Wigner_HWxSU2(DOCS_PT, j=50, seed=9), chart="bloch", **kw)
    assert TWA.last_info["n_failed"] == 0 and np.abs(a - b).max() < 1e-9
    with pytest.raises(pkg_error(), match="Dicke"):
        from dickemodel_jl_b200 import ClassicalLMG as CL
        TWA.average(CL.ClassicalLMGSystem(Ω=1, ξ=-1), observable="Q", N=10, ts=[0.0, 1.0], distribution=TWA.coherent_Wigner_SU2([0, 0], j=10),
                    integrator_alg="RK4", dt=0.1, chart="bloch")


def pkg_error():
    import dickemodel_jl_b200 as pkg
    return pkg.DtwaError


def test_hermite_output_in_the_ensemble_kernels(ctx, mods, O):
    """output="hermite" through TWA.average / calculate_distribution: same trajectories as dtwa_integrate in that mode
    (histogram counts bit-exact against the binned dense output), averages equal to the oracle's, fewer attempts than
    stepping onto every output time"""
    TWA, CD, _, CS = mods
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    j, N = 100, 3000
    u0 = O.sample(O.SAMPLER_HWXSU2, DOCS_PT, 1 / j, N, seed=17)
    ts = TWA.jrange(0, 0.05, 6)
    kw = dict(integrator_alg="DP8", tol=1e-9, output="hermite")
    qs = TWA.jrange(-4.3, 0.02, 4.3)
    mk = lambda **k: TWA.mcs_for_distributions(system, N=N, distribution=TWA.samples_from_array(u0, j=j), ts=ts, **k)   # noqa: E731
    D = TWA.samples_from_array(u0, j=j)
    m_q, avg = TWA.monte_carlo_integrate(system, TWA.mcs_chain(TWA.mcs_for_distributions(system, N=N, distribution=D, ts=ts, y="t", x="q", xs=qs),
                                                               TWA.mcs_for_averaging(system, N=N, distribution=D, ts=ts, observable=["Q", "q*p"])), **kw)
    n_h = TWA.last_info["steps_accepted"] + TWA.last_info["steps_rejected"]
    traj, st, _ = CS.integrate_batch(system, u0, np.asarray(ts), **kw)
    assert (st == 0).all()
    assert np.array_equal(m_q, O.hist_time(traj[:, :, 1], np.ones(traj.shape[:2], bool), *qs.linrange()) / N)
    ref, rst, _, _ = O.integrate(O.SYS_DICKE, DPAR, u0, np.asarray(ts), alg=O.ALG_DOP853, tol=1e-9, output="hermite")
    assert np.abs(avg[:, 0] - ref[:, :, 0].mean(axis=0)).max() < 1e-8
    TWA.average(system, observable="Q", N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), integrator_alg="DP8", tol=1e-9)
    assert n_h < 0.7 * (TWA.last_info["steps_accepted"] + TWA.last_info["steps_rejected"])
    # the interpreter kernel (no run-time specialisation) takes the same path
    ctx.set_jit(0)
    try:
        avg0 = TWA.average(system, observable=["Q", "q*p"], N=N, ts=ts, distribution=TWA.samples_from_array(u0, j=j), **kw)
    finally:
        ctx.set_jit(2)
    assert np.array_equal(avg0, avg)
