"""EnergyShellProjections -- the classical part of src/EnergyShellProjections.jl.

`∫∫dqdpδϵ` (:49-96), `matrix_QP∫∫dqdpδϵ` (:159-199) and `energy_shell_average` (:476-508) for functions f of
the classical point [Q,q,P,p] that can be traced into expressions (like TWA observables).  The quadrature runs
on the GPU (dtwa_shell_quadrature): one warp per (Q,P) cell.  The Husimi projections of quantum states
(proj_husimi_QP_matrix, rényi_occupation; they need DickeBCE) are out of scope.

ASCII aliases: iint_dqdp_delta, matrix_QP_iint_dqdp_delta.
"""
from __future__ import annotations

import math

import numpy as np

from . import ClassicalDicke, _lib
from .TWA import _expr_of

last_info = {}


def _exprs(system, f):
    """f -> (list of expression strings, scalar?)  (f may return a number-like or a list, src/EnergyShellProjections.jl:37-38)"""
    if isinstance(f, (list, tuple)):
        return [_expr_of(system, g) for g in f], False
    if callable(f) and not hasattr(f, "s"):
        from . import symbolic
        out = f(symbolic.symbols(system.varnames()))
        if isinstance(out, (list, tuple)):
            return [_expr_of(system, g) for g in out], False
        return [_expr_of(system, out)], True
    return [_expr_of(system, f)], True


def _root(onlyqroot):
    if onlyqroot is None:
        return 0
    return 1 if onlyqroot in ("+", 1, +1) else -1


def _run(system, ϵ, QP, f, p_res, onlyqroot, devices=None):
    exprs, scalar = _exprs(system, f)
    if devices:   # (Q,P) cells are independent: contiguous blocks of cells on several GPUs at once
        QP = np.ascontiguousarray(QP, dtype=np.float64).reshape(-1, 2)
        out, valid, nodes, ms = _lib.split_over_devices(
            devices, len(QP), lambda c, it: c.shell_quadrature(system._c_system(), ϵ, QP[it], exprs, p_res, _root(onlyqroot)))
    else:
        out, valid, nodes, ms = _lib.default_context().shell_quadrature(system._c_system(), ϵ, QP, exprs, p_res, _root(onlyqroot))
    last_info.clear()
    last_info.update(kernel_ms=ms, nodes=int(nodes.sum()), cells=int(len(valid)), cells_in_shell=int(valid.sum()))
    return out, valid.astype(bool), scalar


def iint_dqdp_delta(system, *, ϵ, Q, P, f, p_res=0.01, nonvalue=float("nan"), onlyqroot=None):
    """src/EnergyShellProjections.jl:49-96"""
    out, valid, scalar = _run(system, ϵ, [[Q, P]], f, p_res, onlyqroot)
    if not valid[0]:
        return nonvalue
    return float(out[0, 0]) if scalar else out[0].copy()


def _generate_QPs(symmetricQP, symmetricP, res, system, ϵ):
    """src/EnergyShellProjections.jl:100-118"""
    minP = -math.ceil(ClassicalDicke.maximum_P_for_ϵ(system, ϵ) / res) * res
    minQ = -math.ceil(ClassicalDicke.maximum_Q_for_ϵ(system, ϵ) / res) * res
    lastQ, lastP = -minQ, -minP
    if symmetricQP or symmetricP:
        lastP = 0
    if symmetricQP:
        lastQ = 0
    if (2 / res) % 1 != 0.0:
        raise ValueError("res has to be an integer fraction of 2")
    nP = int(math.floor((lastP - minP) / res + 1e-9)) + 1
    nQ = int(math.floor((lastQ - minQ) / res + 1e-9)) + 1
    return minQ + res * np.arange(nQ), minP + res * np.arange(nP)


def matrix_QP_iint_dqdp_delta(system, *, f, ϵ, res=0.1, symmetricQP=False, symmetricP=None, parallelize=True, show_progress=True,
                              pbatch_size=None, **kargs):
    """src/EnergyShellProjections.jl:159-199: (Qs, Ps, mat) with mat[i, j] = ∫∫dqdpδϵ(Q=Qs[j], P=Ps[i]) -- the
    reference's orientation (`QPs=[(Q,P) for P in Ps, Q in Qs]`).  All cells in one launch."""
    symmetricP = symmetricQP if symmetricP is None else symmetricP
    Qs, Ps = _generate_QPs(symmetricQP, symmetricP, res, system, ϵ)
    QP = np.column_stack([np.tile(Qs, len(Ps)), np.repeat(Ps, len(Qs))])      # (Q, P) for P in Ps for Q in Qs
    out, valid, scalar = _run(system, ϵ, QP, f, res, kargs.pop("onlyqroot", None), kargs.pop("devices", None))
    nonvalue = kargs.pop("nonvalue", float("nan"))
    if kargs:
        raise TypeError(f"unsupported keyword(s): {sorted(kargs)}")
    out[~valid] = nonvalue if nonvalue is not None else np.nan
    mat = out.reshape(len(Ps), len(Qs), -1)
    if symmetricQP or symmetricP:      # mirror over P (:188-191)
        mat = np.concatenate([mat, mat[-2::-1]], axis=0)
        Ps = np.concatenate([Ps, -Ps[-2::-1]])
    if symmetricQP:                    # mirror over Q (:192-195)
        mat = np.concatenate([mat, mat[:, -2::-1]], axis=1)
        Qs = np.concatenate([Qs, -Qs[-2::-1]])
    return Qs, Ps, (mat[..., 0] if scalar else mat)


def energy_shell_average(system, *, ϵ, f, res=0.01, symmetricQP=False, symmetricP=None, devices=None):
    """src/EnergyShellProjections.jl:476-508"""
    symmetricP = symmetricQP if symmetricP is None else symmetricP
    Qs, Ps = _generate_QPs(symmetricQP, symmetricP, res, system, ϵ)
    QP = np.column_stack([np.repeat(Qs, len(Ps)), np.tile(Ps, len(Qs))])     # the reference's loop order `for Q in Qs, P in Ps`
    out, valid, scalar = _run(system, ϵ, QP, f, res, None, devices)
    wq = np.where((QP[:, 0] != 0) & symmetricQP, 2.0, 1.0)
    wp = np.where((QP[:, 1] != 0) & (symmetricP or symmetricQP), 2.0, 1.0)
    wgt = (wq * wp)[valid]
    t = float(np.sum(2 * math.pi * wgt))
    ft = np.sum(out[valid] * wgt[:, None], axis=0)
    if not valid.any():
        return None
    res_ = ft / t
    return float(res_[0]) if scalar else res_


globals()["∫∫dqdpδϵ"] = iint_dqdp_delta
globals()["matrix_QP∫∫dqdpδϵ"] = matrix_QP_iint_dqdp_delta
