"""Mirror of src/UPOs.jl (periodic orbits of the classical Dicke model) -- a caller of the integrate path: every
integration (with the fundamental matrix, with the p-plane section, dense along an orbit) runs on the GPU through
dtwa_integrate_fm / dtwa_integrate_events / dtwa_integrate; the Newton algebra on 4x4 / 6x5 matrices stays on the
host, as in the reference.

B200-first difference: the monodromy method is BATCHED.  `monodromy_method_constant_period_batch` /
`monodromy_method_constant_energy_batch` refine many guesses at once -- one launch of the variational kernel per
Newton iteration for the whole batch -- and the reference's one-orbit functions are the batch of one.

Not here: `overlap_of_tube_with_homogenous_state`, `scarring_measure` (src/UPOs.jl:888-949) need eigenstates of the
quantum Hamiltonian (DickeBCE), which is outside the hot path.

Integration tolerances below 1e-14 are raised to 1e-14 (the reference passes 1e-16 to a solver that then runs at the
rounding floor; DOP853 in double precision cannot do better than that floor either).
"""
from __future__ import annotations

import bisect
import math

import numpy as np

from . import ClassicalDicke, ClassicalSystems, PhaseSpaces

_TOL_FLOOR = 1e-14
_T_MAX = 10000.0   # the reference's maximum integration time (src/UPOs.jl:306,487)


def _tol(tol):
    return max(float(tol), _TOL_FLOOR)


def flow(system, u):
    """The Dicke vector field F(u) = (dH/dP, dH/dp, -dH/dQ, -dH/dq) at one point or at u[n][4] (host algebra of the
    Newton steps; trajectories are never integrated here)."""
    w0, w, g = system.parameters()
    u = np.asarray(u, dtype=float)
    Q, q, P, p = u[..., 0], u[..., 1], u[..., 2], u[..., 3]
    s = np.sqrt(4.0 - Q * Q - P * P)
    return np.stack([w0 * P - g * q * Q * P / s, w * p, -w0 * Q - g * q * (s - Q * Q / s), -w * q - g * Q * s], axis=-1)


def hamiltonian_gradient(system, u):
    """src/UPOs.jl:319-325: grad H = (-F3, -F4, F1, F2)"""
    F = flow(system, u)
    return np.stack([-F[..., 2], -F[..., 3], F[..., 0], F[..., 1]], axis=-1)


def hamiltonian_hessian(system, u):
    """Hessian of the Dicke Hamiltonian (the reference multiplies the Jacobian of the flow by the symplectic matrix,
    src/UPOs.jl:738-741)"""
    w0, w, g = system.parameters()
    Q, q, P, p = (float(x) for x in u)
    s = math.sqrt(4.0 - Q * Q - P * P)
    s3 = s ** 3
    H = np.zeros((4, 4))
    H[0, 0] = w0 - g * q * Q * (3.0 / s + Q * Q / s3)
    H[0, 1] = H[1, 0] = g * (s - Q * Q / s)
    H[0, 2] = H[2, 0] = -g * q * P * (1.0 / s + Q * Q / s3)
    H[1, 1] = w
    H[1, 2] = H[2, 1] = -g * Q * P / s
    H[2, 2] = w0 - g * q * Q * (1.0 / s + P * P / s3)
    H[3, 3] = w
    return H


class PO:
    """src/UPOs.jl:34-46,91-96: a periodic orbit of `system` through `u` = [Q,q,P,p] with period `T` (if `T` is
    omitted: approximate_period(system, u, bound=1e-2))."""

    def __init__(self, system, u, T=None):
        self.system = system
        self.u = np.asarray(u, dtype=np.float64).copy()
        if self.u.shape != (4,):
            raise ValueError("u should be [Q,q,P,p]")
        self.T = float(approximate_period(system, self.u, bound=1e-2, verbose=False)) if T is None else float(T)

    def __repr__(self):
        return f"PO @ u={self.u.tolist()}, T={self.T}"

    def equals(self, other, *, Ttol=1e-5, tol=1e-6):
        """src/UPOs.jl:53-58: the same orbit regardless of the initial condition on it"""
        if abs(self.T - other.T) >= Ttol:
            return False
        return contains(other, self.u, tol=tol)

    def __eq__(self, other):
        return isinstance(other, PO) and self.equals(other)

    __hash__ = None

    def __contains__(self, u):
        return contains(self, u)


def contains(po, u, *, tol=1e-6):
    """src/UPOs.jl:70-88 (`u in po`): does the orbit come within squared distance `tol` of `u`?  The orbit is sampled
    densely on the GPU (two levels: the whole period, then the neighbourhood of the closest sample) instead of a
    root search on the distance."""
    u = np.asarray(u, dtype=float)
    if np.sum((u - po.u) ** 2) < tol:
        return True
    n = 2048
    ts = np.linspace(0.0, po.T, n + 1)
    s = ClassicalSystems.integrate(po.system, po.u, po.T, tol=1e-8, saveat=ts)
    d2 = np.sum((s.u - u) ** 2, axis=1)
    i = int(np.argmin(d2))
    if d2[i] < tol:
        return True
    if d2[i] > tol + 4.0 * np.max(np.sum(np.diff(s.u, axis=0) ** 2, axis=1)):   # farther than two sample spacings
        return False
    lo, hi = ts[max(i - 1, 0)], ts[min(i + 1, n)]
    fine = ClassicalSystems.integrate(po.system, po.u, hi, tol=1e-8, saveat=np.linspace(lo, hi, 129))
    return bool(np.min(np.sum((fine.u - u) ** 2, axis=1)) < tol)


def integrate(po, *, tol=1e-16, npoints=1024, **kargs):
    """src/UPOs.jl:100-113: `po.u` from t = 0 to t = po.T.  The GPU path reports states at requested times only, so
    the solution carries `npoints` + 1 equally spaced samples (or `saveat=`) where the reference returns solver steps."""
    saveat = kargs.pop("saveat", None)
    if saveat is None and not kargs.get("get_fundamental_matrix") and kargs.get("callback") is None:
        saveat = np.linspace(0.0, po.T, int(npoints) + 1)
    kargs.pop("save_everystep", None)
    return ClassicalSystems.integrate(po.system, po.u, po.T, tol=_tol(tol), saveat=saveat, **kargs)


def action(po, **kargs):
    """src/UPOs.jl:115-130: S = ∮ P dQ + p dq along the orbit (trapezoidal sum over the samples; the reference's
    one-sided sum over solver steps converges to the same integral)"""
    us = integrate(po, **kargs).u
    dQ, dq = np.diff(us[:, 0]), np.diff(us[:, 1])
    return float(np.sum(0.5 * (us[1:, 2] + us[:-1, 2]) * dQ + 0.5 * (us[1:, 3] + us[:-1, 3]) * dq))


def average_over_PO(po, f, *, tol=1e-12, **kargs):
    """src/UPOs.jl:133-181: (1/T) ∫ f(u(t)) dt over one period.  For a PO: trapezoidal rule on equally spaced samples
    (spectrally accurate for a periodic integrand); for a solution object: the reference's sum over its samples."""
    if isinstance(po, PO):
        us = integrate(po, tol=tol, **kargs).u[:-1]          # one period, end point = start point
        vals = [np.asarray(f(x), dtype=float) for x in us]
        tot = sum(vals[1:], vals[0]) / len(vals)
    else:
        T = po.t[-1] - po.t[0]
        tot = None
        for i in range(1, len(po.t)):
            v = np.asarray(f(po.u[i]), dtype=float) * (po.t[i] - po.t[i - 1]) / T
            tot = v if tot is None else tot + v
    return float(tot) if np.ndim(tot) == 0 else tot


def average_over_PO_QP(po, f, **kargs):
    """src/UPOs.jl:183-190"""
    return average_over_PO(po, lambda x: f(x[0], x[2]), **kargs)


def average_over_PO_qp(po, f, **kargs):
    """src/UPOs.jl:193-200"""
    return average_over_PO(po, lambda x: f(x[1], x[3]), **kargs)


def jz_PO_average(po, **kargs):
    return average_over_PO_QP(po, PhaseSpaces.jz, **kargs)


def jy_PO_average(po, **kargs):
    return average_over_PO_QP(po, PhaseSpaces.jy, **kargs)


def jx_PO_average(po, **kargs):
    return average_over_PO_QP(po, PhaseSpaces.jx, **kargs)


def mirror_Qq(po):
    """src/UPOs.jl:234-239"""
    return PO(po.system, po.u * np.array([-1.0, -1.0, 1.0, 1.0]), po.T)


def mirror_Pp(po):
    """src/UPOs.jl:243-248"""
    return PO(po.system, po.u * np.array([1.0, 1.0, -1.0, -1.0]), po.T)


def _p_plane_crossings(system, u0, offset, direction, tol, t_chunks=(64.0, 512.0, 4096.0, _T_MAX)):
    """Crossings of p = offset in time order, generated chunk by chunk (each chunk one launch of the section kernel,
    continued from the end state of the previous one) up to the reference's maximum time: yields (t, u)."""
    t0, u = 0.0, np.asarray(u0, dtype=float)
    for t1 in t_chunks:
        r = ClassicalSystems.poincare_section_batch(system, u[None, :], t1 - t0, coef=[0, 0, 0, 1], offset=offset, direction=direction,
                                                    tol=tol, max_events=4096)
        n = int(min(r["count"][0], 4096))
        for k in range(n):
            yield t0 + float(r["t"][0, k]), r["u"][0, k].copy()
        if r["status"][0] != 0 or r["count"][0] > 4096:
            return
        t0, u = t1, r["u_end"][0]


def approximate_period(system, u0, *, bound=0.1, tol=1e-8, max_p_crossings=math.inf, verbose=True):
    """src/UPOs.jl:252-308: the time at which u0 first comes back to its own p-plane within `bound` of u0 (0 if it
    does not before t = 10000 or before `max_p_crossings` crossings)."""
    u0 = np.asarray(u0, dtype=float)
    missed = 0
    # The reference only listens to crossings with p decreasing (its callback has affect! = nothing, :305), which can
    # never see the return of an orbit that starts with p increasing; here the crossings counted are those in the
    # direction in which u0 itself moves through its p-plane (the same ones as the reference's when dp/dt < 0 at u0).
    direction = -1 if flow(system, u0)[3] < 0 else 1
    for t, u in _p_plane_crossings(system, u0, float(u0[3]), direction, tol):
        if np.linalg.norm(u - u0) < bound:
            return t
        missed += 1
        if missed >= max_p_crossings:
            if verbose:
                print(f"[ Info: Exceeded max_p_crossings = {max_p_crossings}")
            return 0.0
    return 0.0


def find_p_zero(system, u0, *, tol=1e-6, negative=False):
    """src/UPOs.jl:454-496: evolve u0 until it crosses p = 0 upwards (downwards if `negative`); NaN if it never does"""
    u0 = np.asarray(u0, dtype=float)
    for _, u in _p_plane_crossings(system, u0, 0.0, -1 if negative else 1, tol):
        return u
    if abs(u0[3]) < tol:
        return u0
    return math.nan


# ---- the monodromy method, batched -------------------------------------------------------------------------------

def _end_state_and_monodromy(system, U, T, tol, devices=None):
    """(u(T), M(T)) for every row of U (and its own period T[i]): one launch of the variational kernel per distinct T"""
    U = np.asarray(U, dtype=float).reshape(-1, 4)
    T = np.broadcast_to(np.asarray(T, dtype=float), (len(U),))
    u1, M = np.empty_like(U), np.empty((len(U), 4, 4))
    kw = {"devices": devices} if devices else {}
    for Tv in np.unique(T):
        sel = np.nonzero(T == Tv)[0]
        x, fm, status, _ = ClassicalSystems.integrate_fm_batch(system, U[sel], [0.0, float(Tv)], tol=_tol(tol), **kw)
        u1[sel], M[sel] = x[:, -1], fm[:, -1]
        bad = status != 0
        u1[sel[bad]] = np.nan
    return u1, M


def monodromy_method_step_constant_period(system, u0, T, *, tol=1e-12):
    """src/UPOs.jl:311-318 (Eq. (31) of Simonović 1999): (u0 - (M - 1)^-1 (u(T) - u0), |u(T) - u0|)"""
    u0 = np.asarray(u0, dtype=float)
    u1, M = _end_state_and_monodromy(system, u0[None], T, tol)
    return u0 - np.linalg.solve(M[0] - np.eye(4), u1[0] - u0), float(np.linalg.norm(u1[0] - u0))


def _energy_step(system, u0, T, u1, M):
    """src/UPOs.jl:326-338 (App. A.1 of Pilatowsky-Cameo 2021): least-squares solution of
    [[M - 1, F(u1)], [grad H(u1)^T, 0], [xi^T, 0]] (du, dT) = (u0 - u1, 0, 0), xi = e_p"""
    A = np.zeros((6, 5))
    A[:4, :4] = M - np.eye(4)
    A[:4, 4] = flow(system, u1)
    A[4, :4] = hamiltonian_gradient(system, u1)
    A[5, 3] = 1.0
    res = np.linalg.lstsq(A, np.concatenate([u0 - u1, [0.0, 0.0]]), rcond=None)[0]
    return u0 + res[:4], T + res[4]


def monodromy_method_step_constant_energy(system, u0, T, *, tol=1e-12):
    """src/UPOs.jl:326-338: (u, T, |u(T) - u0|) after one step"""
    u0 = np.asarray(u0, dtype=float)
    u1, M = _end_state_and_monodromy(system, u0[None], T, tol)
    un, Tn = _energy_step(system, u0, float(T), u1[0], M[0])
    return un, Tn, float(np.linalg.norm(u1[0] - u0))


def monodromy_method_constant_period_batch(system, U0, T, *, maxiters=100, tol=1e-8, inttol=None, devices=None):
    """Many guesses at once: returns (U[n][4], converged[n], iterations[n]).  Every Newton iteration integrates the
    variational system of all guesses that have not converged in one launch."""
    inttol = tol / 10 if inttol is None else inttol
    U = np.array(U0, dtype=float).reshape(-1, 4)
    T = np.broadcast_to(np.asarray(T, dtype=float), (len(U),)).copy()
    conv, iters = np.zeros(len(U), dtype=bool), np.zeros(len(U), dtype=int)
    alive = np.ones(len(U), dtype=bool)
    for _ in range(maxiters):
        idx = np.nonzero(alive)[0]
        if not len(idx):
            break
        u1, M = _end_state_and_monodromy(system, U[idx], T[idx], inttol, devices)
        for k, i in enumerate(idx):
            iters[i] += 1
            d = u1[k] - U[i]
            if not np.all(np.isfinite(d)):
                alive[i] = False       # left the phase space: this guess does not converge
                continue
            try:
                new = U[i] - np.linalg.solve(M[k] - np.eye(4), d)
            except np.linalg.LinAlgError:
                alive[i] = False
                continue
            if not np.all(np.isfinite(new)) or system.out_of_bounds(new):
                alive[i] = False
                continue
            if np.linalg.norm(d) < tol:               # The reference has already replaced u by the updated point when it
                conv[i], alive[i] = True, False       # tests this distance (:383-386); M - 1 is singular on a periodic
                continue                              # orbit, so that last update can throw a converged point off by far
            U[i] = new                                # more than tol.  Here the point that passed the test is returned.
    return U, conv, iters


def monodromy_method_constant_period(system, u0, T, *, maxiters=100, tol=1e-8, inttol=None):
    """src/UPOs.jl:343-389"""
    U, conv, _ = monodromy_method_constant_period_batch(system, [u0], T, maxiters=maxiters, tol=tol, inttol=inttol)
    if not conv[0]:
        if np.all(np.isfinite(U[0])) and not system.out_of_bounds(U[0]):
            print("┌ Error: Maximum iterations reached")      # the reference logs and returns what it has (:384-386)
        else:
            raise RuntimeError("Convergence error")
    return PO(system, U[0], T)


def monodromy_method_constant_energy_batch(system, U0, T0, *, maxiters=100, tol=1e-8, inttol=None, correct_energy=True, devices=None):
    """Many (u, T) guesses at once, each at the energy of its own guess: (U[n][4], T[n], converged[n], iterations[n])"""
    inttol = tol / 10 if inttol is None else inttol
    U = np.array(U0, dtype=float).reshape(-1, 4)
    T = np.broadcast_to(np.asarray(T0, dtype=float), (len(U),)).copy()
    H = ClassicalDicke.hamiltonian(system)
    E = np.array([H(u) for u in U])
    conv, iters = np.zeros(len(U), dtype=bool), np.zeros(len(U), dtype=int)
    alive = np.ones(len(U), dtype=bool)
    for _ in range(maxiters):
        idx = np.nonzero(alive)[0]
        if not len(idx):
            break
        if correct_energy:     # back onto the energy shell along q (:437-440)
            for i in idx:
                sgn = ClassicalDicke.q_sign(system, U[i])
                try:
                    U[i] = ClassicalDicke.Point(system, Q=U[i][0], P=U[i][2], p=U[i][3], sgn=sgn, ϵ=E[i])
                except ValueError:      # no q puts (Q, P, p) on this energy shell
                    alive[i] = False
            idx = np.nonzero(alive)[0]
            if not len(idx):
                break
        u1, M = _end_state_and_monodromy(system, U[idx], T[idx], inttol, devices)
        for k, i in enumerate(idx):
            iters[i] += 1
            d = u1[k] - U[i]
            if not np.all(np.isfinite(d)):
                alive[i] = False
                continue
            new, Tn = _energy_step(system, U[i], T[i], u1[k], M[k])
            if not (Tn > 0) or not np.all(np.isfinite(new)) or system.out_of_bounds(new):
                alive[i] = False                       # "Convergence error" (:442-444)
                continue
            if np.linalg.norm(d) < tol:               # (as above: return the point that passed the test)
                conv[i], alive[i] = True, False
                continue
            U[i], T[i] = new, Tn
    return U, T, conv, iters


def monodromy_method_constant_energy(system, u0, T, *, maxiters=100, tol=1e-8, inttol=None, correct_energy=True):
    """src/UPOs.jl:391-452"""
    U, Ts, conv, iters = monodromy_method_constant_energy_batch(system, [u0], T, maxiters=maxiters, tol=tol, inttol=inttol,
                                                                correct_energy=correct_energy)
    if not conv[0]:
        raise RuntimeError("Maximum iterations" if iters[0] >= maxiters else "Convergence error")
    return PO(system, U[0], Ts[0])


def monodromy(po, *, tol=1e-12):
    """the fundamental matrix over one period"""
    return _end_state_and_monodromy(po.system, po.u[None], po.T, tol)[1][0]


def lyapunov(po):
    """src/UPOs.jl:498-510 (Eq. (B3) of Pilatowsky-Cameo 2021): max log|eig M(T)| / T"""
    return float(np.max(np.log(np.abs(np.linalg.eigvals(monodromy(po))))) / po.T)


def lyapunov_batch(pos, *, tol=1e-12, devices=None):
    """Lyapunov exponents of many periodic orbits of one system: one launch per distinct period"""
    if not pos:
        return np.empty(0)
    _, M = _end_state_and_monodromy(pos[0].system, np.array([p.u for p in pos]), np.array([p.T for p in pos]), tol, devices)
    return np.array([np.max(np.log(np.abs(np.linalg.eigvals(m)))) / p.T for m, p in zip(M, pos)])


def energy(po):
    """src/UPOs.jl:513-518"""
    return float(ClassicalDicke.hamiltonian(po.system)(po.u))


# ---- families ------------------------------------------------------------------------------------------------------

def follow_PO_family_from_period(po, *, step=0.1, tol=1e-5, initalperturbation=(0, 1, 0, 0), verbose=True):
    """src/UPOs.jl:521-625: returns `get(T, ...) -> PO` of the same family with period T; continuation in the period
    with an adaptive step, results cached (sorted by period)."""
    system = po.system
    us, periods = [po.u.copy()], [float(po.T)]
    step0, tol0 = step, tol

    def get(finalperiod, *, tol=tol0, minstep=None, maxstep=step0, step=None, maxiters=1000):
        minstep = step0 / 1000 if minstep is None else minstep
        step = maxstep if step is None else step
        it = 0
        while True:
            it += 1
            if it > maxiters:
                raise RuntimeError("Maximum iterations")
            index = int(np.argmin(np.abs(finalperiod - np.array(periods))))
            last_p = periods[index]
            left = abs(last_p - finalperiod)
            if left == 0:
                break
            step = min(step, left)
            if verbose:
                print(f"\rCurrent period = {last_p:.3f}, step = 10^{math.log10(step):.1f}, iterations = {it}", end="")
            target = last_p + math.copysign(step, finalperiod - last_p)
            U, conv, _ = monodromy_method_constant_period_batch(system, [us[index]], target, inttol=min(1e-12, tol / 100), tol=tol)
            if not conv[0]:
                if step / 1.2 < minstep:
                    raise RuntimeError(f"The algorithm is not converging. Got to {PO(system, us[index], periods[index])}")
                step /= 1.2
                continue
            step = min(maxstep, step * 1.1)
            k = bisect.bisect_right(periods, target)
            us.insert(k, U[0])
            periods.insert(k, target)
        if verbose:
            print()
        return PO(system, us[index], periods[index])

    return get


def follow_PO_family_from_energy(po, *, step=0.1, tol=1e-5, energytol=1e-3, initalperturbation=(0, 1, 0, 0), verbose=True,
                                 correct_energy=True, dontallowlessenergy=False):
    """src/UPOs.jl:627-803 (App. A.2 of Pilatowsky-Cameo 2021): returns `get(ϵ, ...) -> PO` of the same family with
    energy ϵ (within `energytol`).  A step perturbs the last orbit along `initalperturbation` (first step) or along
    grad H, by the amount a second-order Taylor expansion of H says changes the energy by ±step, guesses the period
    by linear extrapolation and runs the constant-energy monodromy method; the step shrinks by 1.2 on failure and
    grows by 1.1 on success.  Results are cached, sorted by energy."""
    system = po.system
    H = ClassicalDicke.hamiltonian(system)
    E0 = float(H(po.u))
    us, energies, periods = [po.u.copy()], [E0], [float(po.T)]
    step0, tol0, etol0 = step, tol, energytol
    first_dir = np.asarray(initalperturbation, dtype=float)

    def get(finalEnergy, *, tol=tol0, energytol=etol0, maxstep=step0, step=step0, minstep=None, maxiters=1000,
            energywiggletolerance=1e-2):
        minstep = min(step, energytol / 10) if minstep is None else minstep
        if dontallowlessenergy:
            finalEnergy = max(finalEnergy, E0)
        it = 0
        while True:
            it += 1
            if it > maxiters:
                raise RuntimeError("Maximum iterations")
            index = max(bisect.bisect_right(energies, finalEnergy) - 1, 0)
            last_e = energies[index]
            left = abs(last_e - finalEnergy)
            if left < energytol:
                break
            step = min(step, left)
            dE = math.copysign(step, finalEnergy - last_e)
            u = us[index]
            if verbose:
                print(f"\rCurrent energy = {last_e:.3f}, step = 10^{math.log10(step):.1f}, iterations = {it}", end="")
            failed, newu, period = False, u, periods[index]
            try:
                grad = hamiltonian_gradient(system, u)
                du = first_dir if index == 0 else grad
                du = du / np.linalg.norm(du)
                a, b = 0.5 * du @ hamiltonian_hessian(system, u) @ du, float(grad @ du)
                s = (-b + math.sqrt(b * b + 4 * a * dE)) / (2 * a)       # dE = b s + a s^2
                if abs(H(u + s * du) - (last_e + dE)) > abs(H(u - s * du) - (last_e + dE)):
                    s = -s
                E = float(H(u + s * du))
                guess = periods[index]
                if index > 0:
                    guess += (E - energies[index - 1]) * (periods[index] - periods[index - 1]) / (energies[index] - energies[index - 1])
                U, Ts, conv, _ = monodromy_method_constant_energy_batch(system, [u + s * du], guess, inttol=min(1e-12, tol / 100), tol=tol,
                                                                        correct_energy=correct_energy)
                if not conv[0]:
                    raise ArithmeticError("no convergence")
                newu, period = U[0], float(Ts[0])
                if E - H(newu) > energywiggletolerance:       # it fell back: the last cached orbit is suspicious too
                    if index > 0:
                        del energies[index], us[index], periods[index]
                    failed = True
                    index -= 1
            except (ArithmeticError, ValueError, np.linalg.LinAlgError):
                failed = True
            En = float(H(newu))
            if failed or En in energies or abs(finalEnergy - En) > left:
                if step / 1.2 < minstep:
                    i = max(index, 0)
                    raise RuntimeError(f"The algorithm is not converging. Got to {PO(system, us[i], periods[i])} with energy {H(us[i])}")
                step /= 1.2
                continue
            step = min(maxstep, step * 1.1)
            k = bisect.bisect_right(energies, En)
            us.insert(k, newu)
            energies.insert(k, En)
            periods.insert(k, period)
        if verbose:
            print()
        return PO(system, us[index], periods[index])

    return get


def family_A(system, **kargs):
    """src/UPOs.jl:805-821: family A of Pilatowsky-Cameo 2021, born at the ground-state fixed point with the period of
    the positive normal frequency (superradiant regime)"""
    kw = dict(tol=1e-8, correct_energy=False, energytol=1e-3)
    kw.update(kargs)
    return follow_PO_family_from_energy(PO(system, ClassicalDicke.minimum_energy_point(system, "+"),
                                           2 * math.pi / ClassicalDicke.normal_frequency(system, "+")), **kw)


def family_B(system, **kargs):
    """src/UPOs.jl:824-837: family B (negative normal frequency)"""
    kw = dict(tol=1e-8, energytol=1e-3)
    kw.update(kargs)
    return follow_PO_family_from_energy(PO(system, ClassicalDicke.minimum_energy_point(system, "+"),
                                           2 * math.pi / ClassicalDicke.normal_frequency(system, "-")), **kw)


def po_coordinates(coords, po, *, tol=1e-12, npoints=1024):
    """src/UPOs.jl:840-853: [(x[c] for c in coords) for x along po]; `coords` are the reference's 1-based indices"""
    us = integrate(po, tol=tol, npoints=npoints).u
    return [tuple(float(x[c - 1]) for c in coords) for x in us]


def QP(po, **kw):
    return po_coordinates([1, 3], po, **kw)


def qp(po, **kw):
    return po_coordinates([2, 4], po, **kw)


def QPp(po, **kw):
    return po_coordinates([1, 3, 4], po, **kw)


def QPq(po, **kw):
    return po_coordinates([1, 3, 2], po, **kw)


def overlap_of_tube_with_homogenous_state(*a, **k):
    raise NotImplementedError("needs the quantum Dicke Hamiltonian's eigenstates (DickeBCE): outside the GPU hot path (SURVEY.md section 8)")


scarring_measure = overlap_of_tube_with_homogenous_state
