"""PhaseSpaces -- (Q,P) <-> (theta,phi) closed forms used by the samplers.
Mirrors src/PhaseSpaces.jl:16,34,49,63,80-100,112,127 of the reference (host scalar code there too)."""
import math


def Q_of_θϕ(θ, ϕ):
    """src/PhaseSpaces.jl:16"""
    return math.sqrt(2 * (1 - math.cos(θ))) * math.cos(ϕ)


def P_of_θϕ(θ, ϕ):
    """src/PhaseSpaces.jl:34"""
    return -math.sqrt(2 * (1 - math.cos(θ))) * math.sin(ϕ)


def θ_of_QP(Q, P):
    """src/PhaseSpaces.jl:49"""
    return math.acos(1 - (Q ** 2 + P ** 2) / 2)


def ϕ_of_QP(Q, P):
    """src/PhaseSpaces.jl:63 -- Julia's mod(x, 2pi) carries the sign of the divisor"""
    r = math.fmod(math.atan2(-P, Q), 2 * math.pi)
    return r + 2 * math.pi if r < 0 else r


def jz(Q, P):
    return -math.cos(θ_of_QP(Q, P))


def jx(Q, P):
    return math.sin(θ_of_QP(Q, P)) * math.cos(ϕ_of_QP(Q, P))


def jy(Q, P):
    return math.sin(θ_of_QP(Q, P)) * math.sin(ϕ_of_QP(Q, P))


def arc_between_θϕ(θ1, ϕ1, θ2, ϕ2):
    """src/PhaseSpaces.jl:112"""
    c = math.cos(θ1) * math.cos(θ2) + math.sin(θ1) * math.sin(θ2) * math.cos(ϕ1 - ϕ2)
    return math.acos(max(-1.0, min(1.0, c)))


def arc_between_QP(Q1, P1, Q2, P2):
    """src/PhaseSpaces.jl:127"""
    return arc_between_θϕ(θ_of_QP(Q1, P1), ϕ_of_QP(Q1, P1), θ_of_QP(Q2, P2), ϕ_of_QP(Q2, P2))


Q_of_θφ, P_of_θφ, φ_of_QP, arc_between_θφ = Q_of_θϕ, P_of_θϕ, ϕ_of_QP, arc_between_θϕ
# ASCII aliases
Q_of_thetaphi, P_of_thetaphi, theta_of_QP, phi_of_QP = Q_of_θϕ, P_of_θϕ, θ_of_QP, ϕ_of_QP
