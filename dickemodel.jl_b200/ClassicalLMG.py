"""ClassicalLMG -- the classical Lipkin-Meshkov-Glick model through the same ClassicalSystems API.
Mirrors src/ClassicalLMG.jl: ClassicalLMGSystem (:32-36), hamiltonian (:63-70),
discriminant_of_P_solution (:85-88), P_of_ϵ (:110-125), minimum_ϵ_for (:140-152), Point (:162-184).
Equations of motion (:8-11) and domain (:12-14) live in the CUDA kernels."""
from __future__ import annotations

import math

import numpy as np

from . import ClassicalSystems, PhaseSpaces
from . import _lib

_ALIASES = {"Ω": "Om", "Om": "Om", "Omega": "Om", "omega": "Om", "ξ": "xi", "xi": "xi"}


class ClassicalLMGSystem(ClassicalSystems.ClassicalSystem):
    _kind = _lib.SYS_LMG

    def __init__(self, **kw):
        vals = {}
        for k, v in kw.items():
            if k not in _ALIASES:
                raise TypeError(f"unknown parameter {k!r}; expected Ω, ξ")
            vals[_ALIASES[k]] = float(v)
        if set(vals) != {"Om", "xi"}:
            raise TypeError("ClassicalLMGSystem needs Ω and ξ")
        self._parameters = [vals["Om"], vals["xi"]]

    def parameters(self):
        return self._parameters

    def varnames(self):
        return ["Q", "P"]  # :15

    def vector_field(self, u, pars=None):
        """src/ClassicalLMG.jl:8-11: (dQ, dP)/dt (host formula, see ClassicalSystems.step)"""
        Om, xi = self.parameters() if pars is None else pars
        Q, P = u
        return [P * (Om - xi * Q ** 2 / 2), Q * (xi * P ** 2 / 2 - (2 * xi + Om)) + xi * Q ** 3]

    def out_of_bounds(self, u):
        return 4 - u[1] ** 2 - u[0] ** 2 <= 0  # :12-14

    def __repr__(self):
        return f"ClassicalLMGSystem(Ω = {self._parameters[0]}, ξ = {self._parameters[1]})"


def hamiltonian(system):
    Om, xi = system.parameters()

    def H(u):
        Q, P = u
        return Om * (Q ** 2 + P ** 2) / 2 + xi * (Q ** 2 - P ** 2 * Q ** 2 / 4 - Q ** 4 / 4) - Om
    return H


def discriminant_of_P_solution(system, Q, ϵ):
    Om, xi = system.parameters()
    return (-4 * ϵ + 4 * Q ** 2 * xi - Q ** 4 * xi - 4 * Om + 2 * Q ** 2 * Om) / (Q ** 2 * xi - 2 * Om)


def minimum_ϵ_for(system, *, Q=None, P=None):
    if [Q, P].count(None) != 1:
        raise ValueError("You have to pass one of Q or P")
    Om, xi = system.parameters()
    if P is None:
        return (2 * Om * Q ** 2 + 4 * Q ** 2 * xi - Q ** 4 * xi) / 4 - Om
    return Om * P ** 2 / 2 - Om


def P_of_ϵ(system, *, Q, ϵ, sgn="+", returnNaNonError=True):
    Δ = discriminant_of_P_solution(system, Q, ϵ)
    if Δ < 0:
        if returnNaNonError:
            return float("nan")
        raise ValueError(f"The minimum energy for the given parameters is {minimum_ϵ_for(system, Q=Q)}. There is no P such that (Q={Q},P) has ϵ={ϵ}.")
    return math.sqrt(Δ) if sgn in ("+", +1, 1) else -math.sqrt(Δ)


def Point(system=None, *, Q, P=None, ϵ=None, sgn="+"):
    if system is None:
        return np.array([Q, P], dtype=np.float64)
    return np.array([Q, P_of_ϵ(system, Q=Q, ϵ=ϵ, sgn=sgn, returnNaNonError=False)], dtype=np.float64)


def Pointθϕ(*, θ, ϕ):
    return Point(Q=PhaseSpaces.Q_of_θϕ(θ, ϕ), P=PhaseSpaces.P_of_θϕ(θ, ϕ))


P_of_eps = P_of_ϵ
