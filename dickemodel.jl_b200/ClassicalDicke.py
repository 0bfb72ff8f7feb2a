"""ClassicalDicke -- the classical Dicke model: system type, Hamiltonian, initial-point helpers.

Mirrors src/ClassicalDicke.jl of the reference: ClassicalDickeSystem (:48-52), hamiltonian (:80-88),
normal_frequency (:139-147), minimum_energy_point (:164-177), minimum_energy (:189-197),
discriminant_of_q_solution (:224-228), q_of_ϵ (:255-273), q_sign (:292-303), minimum_ϵ_for (:319-336),
Point (:407-427).  These are host scalar formulas in the reference too; the equations of motion
(:11-17) and the domain (:18-20) live in the CUDA kernels (csrc/dtwa_device.cuh).
"""
from __future__ import annotations

import math

import numpy as np

from . import ClassicalSystems, PhaseSpaces, symbolic
from . import _lib

_ALIASES = {"ω₀": "w0", "ω0": "w0", "w0": "w0", "omega0": "w0", "omega_0": "w0",
            "ω": "w", "w": "w", "omega": "w", "γ": "g", "g": "g", "gamma": "g"}


class DickeSystem(ClassicalSystems.ClassicalSystem):
    """src/ClassicalDicke.jl:30 abstract type DickeSystem"""
    _kind = _lib.SYS_DICKE

    def varnames(self):
        return ["Q", "q", "P", "p"]  # :21

    def vector_field(self, u, pars=None):
        """src/ClassicalDicke.jl:11-16: (dQ, dq, dP, dp)/dt at u = [Q, q, P, p] (host formula, see ClassicalSystems.step)"""
        w0, w, g = self.parameters() if pars is None else pars
        Q, q, P, p = u
        s = math.sqrt(4.0 - Q * Q - P * P)
        return [w0 * P - g * q * Q * P / s, w * p, -w0 * Q - g * q * (s - Q * Q / s), -w * q - g * Q * s]

    def out_of_bounds(self, u):
        return 4 - u[2] ** 2 - u[0] ** 2 <= 0  # :18-20


class ClassicalDickeSystem(DickeSystem):
    """ClassicalDickeSystem(ω₀=..., ω=..., γ=...) -- src/ClassicalDicke.jl:48-52.  Python identifiers
    cannot contain the subscript zero, so ω₀ may be given as ω0 / w0 / omega0 or as **{"ω₀": 1.0}."""

    def __init__(self, **kw):
        vals = {}
        for k, v in kw.items():
            if k not in _ALIASES:
                raise TypeError(f"unknown parameter {k!r}; expected ω₀, ω, γ")
            vals[_ALIASES[k]] = float(v)
        if set(vals) != {"w0", "w", "g"}:
            raise TypeError("ClassicalDickeSystem needs ω₀, ω and γ")
        self._parameters = [vals["w0"], vals["w"], vals["g"]]

    def parameters(self):
        return self._parameters

    def __repr__(self):
        w0, w, g = self._parameters
        return f"ClassicalDickeSystem(ω₀ = {w0}, ω = {w}, γ = {g})"


def hamiltonian(system: DickeSystem):
    """src/ClassicalDicke.jl:80-88.  The returned H accepts a vector of numbers or of symbolic
    variables (so it can be passed as a TWA observable, docs/src/TWAExamples.md:374)."""
    w0, w, g = system.parameters()

    def H(u):
        Q, q, P, p = u
        r = 1 - (Q ** 2 + P ** 2) / 4
        return (w0 / 2) * (Q ** 2 + P ** 2) + (w / 2) * (q ** 2 + p ** 2) + 2 * g * q * Q * symbolic.sqrt(r) - w0
    return H


def energy_shell_volume(system, ϵ):
    """src/ClassicalDicke.jl:110-124 (the reference integrates Pmax(Q) with QuadGK; here Gauss-Legendre on the
    substitution Q = Qmin + (Qmax-Qmin) sin^2(t), which removes the square-root end-point singularities)"""
    w0, w, g = system.parameters()
    if ϵ <= minimum_energy(system):
        return 0.0
    if ϵ >= w0:
        return 8 * math.pi ** 2 / w
    a, b = minimum_nonnegative_Q_for_ϵ(system, ϵ), maximum_Q_for_ϵ(system, ϵ)
    import numpy as np
    x, wt = np.polynomial.legendre.leggauss(200)
    t = (x + 1) * math.pi / 4
    Q = a + (b - a) * np.sin(t) ** 2
    arg = (w * (2 * ϵ + 2 * w0) - Q ** 2 * ((-4 + Q ** 2) * g ** 2 + w * w0)) / (Q ** 2 * g ** 2 + w * w0)
    Pmax = np.sqrt(np.maximum(arg, 0.0))
    integral = float(np.sum(wt * Pmax * (b - a) * np.sin(2 * t)) * math.pi / 4)
    return integral * 2 * math.pi * 4 / w


def normal_frequency(system, sgn="+"):
    w0, w, g = system.parameters()
    gc = math.sqrt(w0 * w) / 2
    if gc > g:
        raise ValueError("No methods for the normal phase (please add it!)")
    s = 1.0 if sgn in ("+", +1, 1) else -1.0
    return math.sqrt((1 / (2 * gc ** 4)) * (w ** 2 * gc ** 4 + w0 ** 2 * g ** 4
                                            + s * math.sqrt((w0 ** 2 * g ** 4 - w ** 2 * gc ** 4) ** 2 + 4 * gc ** 8 * w ** 2 * w0 ** 2)))


def minimum_energy_point(system, Qsign="+"):
    w0, w, g = system.parameters()
    gc = math.sqrt(w0 * w) / 2
    if gc >= g:
        return Point(Q=0, q=0, P=0, p=0)
    s = 1.0 if Qsign in ("+", +1, 1) else -1.0
    return Point(Q=s * math.sqrt(2 - w * w0 / (2 * g ** 2)), q=-s * math.sqrt(4 * g ** 2 / w ** 2 - w0 ** 2 / (4 * g ** 2)), P=0, p=0)


def minimum_energy(system):
    w0, w, g = system.parameters()
    gc = math.sqrt(w0 * w) / 2
    if g <= gc:
        return -w0
    return -w0 * ((gc / g) ** 2 + (g / gc) ** 2) / 2


def discriminant_of_q_solution(system, *, Q, P, p, ϵ):
    w0, w, g = system.parameters()
    r = 1 - (Q ** 2 + P ** 2) / 4
    return 4 * g ** 2 * Q ** 2 * r - p ** 2 * w ** 2 - w * w0 * (P ** 2 + Q ** 2 - 2) + 2 * w * ϵ


def minimum_ϵ_for(system, *, Q=None, q=None, P=None, p=None):
    if [Q, q, P, p].count(None) != 1:
        raise ValueError("You have to give three of Q,q,P,p")
    w0, w, g = system.parameters()
    if q is None:
        r = 1 - (Q ** 2 + P ** 2) / 4
        return -(4 * g ** 2 * Q ** 2 * r - p ** 2 * w ** 2 - w * w0 * (P ** 2 + Q ** 2 - 2)) / (2 * w)
    if Q is None:
        return (2 * p ** 2 * w + 2 * q ** 2 * w + P ** 2 * w0 - math.sqrt((-4 + P ** 2) ** 2 * (4 * q ** 2 * g ** 2 + w0 ** 2))) / 4
    raise NotImplementedError("Sorry, that combinanation of variables is not implemented")


def _clamp(x, lo, hi):
    return min(max(x, lo), hi)


def maximum_P_for_ϵ(system, ϵ):
    """src/ClassicalDicke.jl:351-360"""
    w0, w, g = system.parameters()
    if ϵ >= w0:
        return 2.0
    if w * w0 > math.sqrt(2) * g * math.sqrt(w * (w0 - ϵ)):
        return math.sqrt(2 * (w0 + ϵ) / w0)
    return math.sqrt(_clamp((2 * math.sqrt(2 * g ** 6 * w * (w0 - ϵ))) / (-g ** 4) + w * w0 / (g ** 2) + 4, 0, 4))


def maximum_Q_for_ϵ(system, ϵ):
    """src/ClassicalDicke.jl:372-378"""
    w0, w, g = system.parameters()
    if ϵ >= w0:
        return 2.0
    return math.sqrt(_clamp((4 * g ** 2 - w * w0) / (2 * g ** 2)
                            + (1 / 2) * math.sqrt(max(0, (16 * g ** 4 + 8 * g ** 2 * ϵ * w + w ** 2 * w0 ** 2) / g ** 4)), 0, 4))


def minimum_nonnegative_Q_for_ϵ(system, ϵ):
    """src/ClassicalDicke.jl:390-396"""
    w0, w, g = system.parameters()
    if ϵ >= -w0:
        return 0.0
    return math.sqrt(_clamp((4 * g ** 2 - w * w0) / (2 * g ** 2)
                            - (1 / 2) * math.sqrt(max(0, (16 * g ** 4 + 8 * g ** 2 * ϵ * w + w ** 2 * w0 ** 2) / g ** 4)), 0, 4))


def q_of_ϵ(system, *, Q, P, p, ϵ, sgn="+", returnNaNonError=True):
    """src/ClassicalDicke.jl:255-273"""
    Δ = discriminant_of_q_solution(system, Q=Q, P=P, p=p, ϵ=ϵ)
    if Δ < 0:
        if returnNaNonError:
            return float("nan")
        ϵmin = minimum_ϵ_for(system, Q=Q, P=P, p=p)
        raise ValueError(f"The minimum energy for the given parameters is {ϵmin}. There is no q such that (Q={Q},q,P={P},p={p}) has ϵ={ϵ}.")
    w0, w, g = system.parameters()
    s = 1.0 if sgn in ("+", +1, 1) else -1.0
    return (-2 * g * Q * math.sqrt(1 - (Q ** 2 + P ** 2) / 4) + s * math.sqrt(Δ)) / w


def q_sign(system, x, ϵ=None):
    Q, q, P, p = x
    if ϵ is None:
        ϵ = hamiltonian(system)(x)
    qp = q_of_ϵ(system, Q=Q, P=P, p=p, ϵ=ϵ, sgn="+")
    qm = q_of_ϵ(system, Q=Q, P=P, p=p, ϵ=ϵ, sgn="-")
    return "+" if abs(q - qp) < abs(q - qm) else "-"


def Point(system=None, *, Q, P, p, q=None, ϵ=None, sgn="+"):
    """Point(Q=,q=,P=,p=) -> [Q,q,P,p]  (src/ClassicalDicke.jl:407), or
    Point(system, Q=,P=,p=,ϵ=[,sgn=]) with q solved from the energy (:419-427)."""
    if system is None:
        if q is None:
            raise TypeError("Point needs q (or a system and ϵ)")
        return np.array([Q, q, P, p], dtype=np.float64)
    if ϵ is None:
        raise TypeError("Point(system, ...) needs ϵ")
    return np.array([Q, q_of_ϵ(system, Q=Q, P=P, p=p, ϵ=ϵ, sgn=sgn, returnNaNonError=False), P, p], dtype=np.float64)


def Pointθϕ(*, θ, ϕ, q, p):
    return Point(Q=PhaseSpaces.Q_of_θϕ(θ, ϕ), q=q, P=PhaseSpaces.P_of_θϕ(θ, ϕ), p=p)


def phase_space_dist_squared(x, y):
    Q1, q1, P1, p1 = x
    Q2, q2, P2, p2 = y
    return (q1 - q2) ** 2 + (p1 - p2) ** 2 + PhaseSpaces.arc_between_QP(Q1, P1, Q2, P2) ** 2


q_of_eps, minimum_eps_for = q_of_ϵ, minimum_ϵ_for


def classical_path_random_sampler(system, *, ϵ, dt=3, seed=None):
    """src/ClassicalDicke.jl:502-531: sample() walks along one chaotic trajectory (integrated on the GPU) in
    steps of dt, flipping (Q, q) -> (-Q, -q) at random.  `seed` makes the host-side coin flips reproducible."""
    import random
    from . import ClassicalSystems
    rng = random.Random(seed)
    w0, w, g = system.parameters()
    s = math.sqrt(16 * g ** 4 + 8 * g ** 2 * ϵ * w + w ** 2 * w0 ** 2) / (2 * g ** 2)
    a = 2 - w * w0 / (2 * g ** 2)
    mx = math.sqrt(a + s)
    mn = 0.0 if a < s else math.sqrt(a - s)
    Q = mn if mn == mx else rng.uniform(mn, mx)
    Q = Q * rng.choice((-1, 1))
    state = {"u": Point(system, Q=Q, P=0, p=0, ϵ=ϵ)}

    def sample():
        u = ClassicalSystems.integrate(system, state["u"], dt, save_everystep=False, show_progress=False).u[-1].copy()
        if rng.choice((True, False)):
            u[0], u[1] = -u[0], -u[1]
        state["u"] = u
        return u.copy()
    for _ in range(10):
        sample()
    return sample

