"""Tiny symbolic layer standing in for SymEngine on the reference's observable path.

The reference accepts a Symbol, a Julia Expr, a SymEngine.Basic or a Julia function as an
observable and turns all of them into a callable (src/TWA.jl:277-299).  A GPU needs an expression,
so here every accepted form is reduced to an `Expr` whose string is handed to the library's
expression compiler (csrc/dtwa_expr.hpp): strings and Expr objects pass through, callables are
traced by calling them on symbols.  Rendering is fully parenthesised, so the operation order of
the Python expression is the operation order on the device.
"""
from __future__ import annotations

import math
import numbers


class Expr:
    __slots__ = ("s",)
    __array_priority__ = 1000

    def __init__(self, s: str):
        self.s = s

    def __str__(self):
        return self.s

    def __repr__(self):
        return f"Expr({self.s})"

    @staticmethod
    def wrap(x):
        if isinstance(x, Expr):
            return x
        if isinstance(x, str):
            return Expr(f"({x})")
        if isinstance(x, bool):
            raise TypeError("booleans are not observables")
        if isinstance(x, numbers.Integral):
            return Expr(repr(float(int(x))) if abs(int(x)) < 2 ** 53 else repr(float(x)))
        if isinstance(x, numbers.Real):
            v = float(x)
            if math.isnan(v) or math.isinf(v):
                raise ValueError("non-finite constant in an observable expression")
            return Expr(f"({v!r})" if v < 0 else repr(v))
        raise TypeError(f"cannot use {type(x).__name__} in an observable expression")

    def _bin(self, op, other, swap=False):
        o = Expr.wrap(other)
        a, b = (o, self) if swap else (self, o)
        return Expr(f"({a.s}{op}{b.s})")

    def __add__(self, o): return self._bin("+", o)
    def __radd__(self, o): return self._bin("+", o, True)
    def __sub__(self, o): return self._bin("-", o)
    def __rsub__(self, o): return self._bin("-", o, True)
    def __mul__(self, o): return self._bin("*", o)
    def __rmul__(self, o): return self._bin("*", o, True)
    def __truediv__(self, o): return self._bin("/", o)
    def __rtruediv__(self, o): return self._bin("/", o, True)
    def __neg__(self): return Expr(f"(-{self.s})")
    def __pos__(self): return self

    def __pow__(self, o):
        if isinstance(o, numbers.Integral) and not isinstance(o, bool):
            return Expr(f"({self.s}^{int(o)})" if o >= 0 else f"({self.s}^(-{-int(o)}))")
        if isinstance(o, numbers.Real) and float(o) == int(o) and abs(o) <= 1024:
            return self.__pow__(int(o))
        return Expr(f"pow({self.s},{Expr.wrap(o).s})")

    def __rpow__(self, o):
        return Expr(f"pow({Expr.wrap(o).s},{self.s})")


def symbols(names):
    return [Expr(n) for n in names]


def _fn(name, pyf):
    def f(x, *rest):
        if isinstance(x, Expr) or any(isinstance(r, Expr) for r in rest):
            args = ",".join(Expr.wrap(a).s for a in (x,) + rest)
            return Expr(f"{name}({args})")
        return pyf(x, *rest)
    f.__name__ = name
    return f


sqrt = _fn("sqrt", math.sqrt)
sin = _fn("sin", math.sin)
cos = _fn("cos", math.cos)
tan = _fn("tan", math.tan)
exp = _fn("exp", math.exp)
log = _fn("log", math.log)
acos = _fn("acos", math.acos)
asin = _fn("asin", math.asin)
atan = _fn("atan", lambda y, *x: math.atan2(y, *x) if x else math.atan(y))
atan2 = _fn("atan2", math.atan2)
abs_ = _fn("abs", abs)
