"""DickeBCE -- only the system carrier of src/DickeBCE.jl.

`QuantumDickeSystem` (src/DickeBCE.jl:51-72) is a `DickeSystem` that carries `j` (and `Nmax`) and delegates
`parameters` to its classical system, so that every function of the classical path -- `integrate`, `TWA.average`,
`calculate_distribution`, ... -- accepts it, as in the reference's docs.  The quantum side of the module (efficient
coherent basis, diagonalisation, Husimi/Wigner functions) is a different workload and out of scope.
"""
from __future__ import annotations

from . import ClassicalDicke


class QuantumDickeSystem(ClassicalDicke.DickeSystem):
    def __init__(self, classical_system=None, *, j, Nmax=None, **kw):
        if classical_system is None:                       # QuantumDickeSystem(;ω₀,ω,γ,j,Nmax) (:75-80)
            classical_system = ClassicalDicke.ClassicalDickeSystem(**kw)
        elif kw:
            raise TypeError(f"unexpected keyword(s) {sorted(kw)}")
        if (2 * j) != int(2 * j) or j < 0:
            raise ValueError(f"j must be a positive half-integer. Got j = {j}.")       # :57-59
        if Nmax is not None and Nmax <= 0:
            raise ValueError(f"Nmax must be a positive integer. Got Nmax = {Nmax}.")    # :60-62
        self.j = int(j) if j == int(j) else j
        self.Nmax = Nmax
        self.classical_system = classical_system

    def parameters(self):                                   # :70-72
        return self.classical_system.parameters()

    def __repr__(self):
        return f"QuantumDickeSystem({self.classical_system!r}, j = {self.j}, Nmax = {self.Nmax})"
