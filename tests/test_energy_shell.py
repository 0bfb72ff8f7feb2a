"""the classical energy-shell quadrature of src/EnergyShellProjections.jl:49-96,159-199,476-508."""
import math

import numpy as np
import pytest

PAR = [1.0, 1.0, 1.0]   # w0, w, g


def test_oracle_shell_quadrature_reproduces_the_shell_volume(O):
    """sum over a (Q,P) grid of ∫∫dqdpδϵ[f=1] · res² -> ν-volume of the shell; for ϵ >= ω₀ that is 8π²/ω
    (src/ClassicalDicke.jl:115-117)"""
    res, eps = 0.1, 1.5
    Qs = np.arange(-2, 2 + 1e-9, res)
    tot = 0.0
    for Q in Qs:
        for P in Qs:
            v = O.shell_quadrature(PAR, eps, Q, P, ["1"], p_res=res)
            if v is not None:
                tot += v[0] * res * res
    assert abs(tot / (8 * math.pi ** 2 / PAR[1]) - 1) < 0.02
    assert O.shell_quadrature(PAR, eps, 1.9, 1.9, ["1"]) is None                 # outside the Bloch disk
    assert O.shell_quadrature(PAR, -1.9, 0.0, 0.0, ["1"]) is None                # below the available energy
    # f = p is odd in p: the Chebyshev nodes are symmetric, the integral vanishes; f = p^2 is positive
    v = O.shell_quadrature(PAR, 0.5, 0.4, -0.3, ["p", "p^2", "q"], p_res=0.01)
    assert abs(v[0]) < 1e-12 and v[1] > 0
    vp = O.shell_quadrature(PAR, 0.5, 0.4, -0.3, ["q"], p_res=0.01, onlyqroot=+1)
    vm = O.shell_quadrature(PAR, 0.5, 0.4, -0.3, ["q"], p_res=0.01, onlyqroot=-1)
    assert abs(vp[0] + vm[0] - v[2]) < 1e-12


def csys(pkg, par):
    s = pkg._lib.System()
    s.kind = 0
    for i, v in enumerate(par):
        s.params[i] = v
    return s


@pytest.mark.gpu
def test_gpu_shell_quadrature_matches_oracle(ctx, pkg, O):
    rng = np.random.default_rng(11)
    QP = np.column_stack([rng.uniform(-2.1, 2.1, 64), rng.uniform(-2.1, 2.1, 64)])
    QP[0] = [0.0, 0.0]
    QP[1] = [2.0, 0.0]
    exprs = ["1", "q^2 + p^2", "Q*q*sqrt(4 - Q^2 - P^2)", "cos(q)*p^2"]
    for eps, p_res, root in ((0.5, 0.01, 0), (-1.2, 0.02, 0), (1.5, 0.05, +1), (0.5, 0.013, -1)):
        out, valid, nodes, ms = ctx.shell_quadrature(csys(pkg, PAR), eps, QP, exprs, p_res, root)
        nin = 0
        for i, (Q, P) in enumerate(QP):
            want = O.shell_quadrature(PAR, eps, Q, P, exprs, p_res=p_res, onlyqroot=root)
            if want is None:
                assert not valid[i] and np.isnan(out[i]).all()
            else:
                nin += 1
                assert valid[i]
                w = np.array(want)
                assert np.abs(out[i] - w).max() <= 1e-11 * max(1.0, np.abs(w).max()), (eps, i, out[i], w)
        assert nin >= 5
    with pytest.raises(pkg.DtwaError, match="unknown variable"):
        ctx.shell_quadrature(csys(pkg, PAR), 0.5, QP, ["z"], 0.01, 0)


@pytest.mark.gpu
def test_energy_shell_projections_api(pkg, O):
    """the reference's entry points: ∫∫dqdpδϵ, matrix_QP∫∫dqdpδϵ (orientation, mirroring), energy_shell_average"""
    from dickemodel_jl_b200 import ClassicalDicke as CD, EnergyShellProjections as ESP
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    f1 = lambda x: 1
    v = getattr(ESP, "∫∫dqdpδϵ")(system, ϵ=0.5, Q=0.4, P=-0.3, f=lambda x: x[1] ** 2 + x[3] ** 2, p_res=0.01)
    assert abs(v - O.shell_quadrature(PAR, 0.5, 0.4, -0.3, ["q^2 + p^2"])[0]) < 1e-10
    assert math.isnan(ESP.iint_dqdp_delta(system, ϵ=0.5, Q=1.9, P=1.9, f=f1))
    assert ESP.iint_dqdp_delta(system, ϵ=0.5, Q=1.9, P=1.9, f=f1, nonvalue=None) is None
    eps, res = -0.5, 0.1
    Qs, Ps, mat = getattr(ESP, "matrix_QP∫∫dqdpδϵ")(system, f=f1, ϵ=eps, res=res, show_progress=False)
    assert mat.shape == (len(Ps), len(Qs))
    i, j = len(Ps) // 2 + 2, len(Qs) // 2 - 5
    assert abs(mat[i, j] - O.shell_quadrature(PAR, eps, Qs[j], Ps[i], ["1"], p_res=res)[0]) < 1e-10
    vol = np.nansum(mat) * res * res
    assert abs(vol / CD.energy_shell_volume(system, eps) - 1) < 0.03
    Qs2, Ps2, mat2 = ESP.matrix_QP_iint_dqdp_delta(system, f=f1, ϵ=eps, res=res, symmetricQP=True, show_progress=False)
    assert np.allclose(Qs2, Qs) and np.allclose(Ps2, Ps) and np.allclose(mat2, mat, equal_nan=True, atol=1e-12)
    # energy_shell_average of h_cl itself is ϵ; <q^2 + p^2> agrees between full and symmetric grids
    h = CD.hamiltonian(system)
    assert abs(ESP.energy_shell_average(system, ϵ=eps, f=h, res=0.05) - eps) < 1e-9
    a = ESP.energy_shell_average(system, ϵ=eps, f=lambda x: [x[1] ** 2 + x[3] ** 2, x[0] ** 2], res=0.05)
    b = ESP.energy_shell_average(system, ϵ=eps, f=lambda x: [x[1] ** 2 + x[3] ** 2, x[0] ** 2], res=0.05, symmetricQP=True)
    assert np.allclose(a, b, rtol=1e-9)
    assert abs(CD.energy_shell_volume(system, 1.5) - 8 * math.pi ** 2) < 1e-12 and CD.energy_shell_volume(system, -5.0) == 0.0


@pytest.mark.gpu
def test_classical_path_random_sampler_stays_on_the_shell(pkg, O):
    """src/ClassicalDicke.jl:502-531: points along one chaotic trajectory, all at energy ϵ"""
    from dickemodel_jl_b200 import ClassicalDicke as CD
    system = CD.ClassicalDickeSystem(w0=1, w=1, g=1)
    sample = CD.classical_path_random_sampler(system, ϵ=-0.5, dt=3, seed=4)
    pts = np.array([sample() for _ in range(6)])
    assert np.abs(O.hamiltonian(O.SYS_DICKE, PAR, pts) + 0.5).max() < 1e-9
    # (flipping only the positions (Q, q), as the reference does, is a time-reversal: the walk may retrace itself)
    assert len({tuple(np.round(np.abs(p), 6)) for p in pts}) >= 2
