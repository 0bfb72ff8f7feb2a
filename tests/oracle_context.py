"""Test infrastructure: a stand-in for _lib.Context whose integrate / integrate_fm / integrate_events call the CPU
oracle.  It lets the `-m "not gpu"` suite exercise HOST logic that sits on top of the GPU entry points (the Newton and
continuation loops of UPOs.py) without a GPU.  Never imported by the package."""
import numpy as np

from oracle import oracle as O


class OracleContext:
    def _pb(self, sys, sol):
        par = [sys.params[i] for i in range(4)]
        return dict(alg=int(sol.alg), dt=float(sol.dt), tol=float(sol.tol), dtmin=float(sol.dtmin) or None,
                    maxiters=int(sol.maxiters) or 10 ** 7, backwards=bool(sol.backwards), fast=True), int(sys.kind), par

    @staticmethod
    def _ctrl(sol):
        if int(sol.controller) == 1:
            return "scipy"
        return (sol.beta1, sol.beta2, sol.qmin or (0.2 if int(sol.alg) == O.ALG_DP5 else 1 / 3), sol.qmax or (10.0 if int(sol.alg) == O.ALG_DP5 else 6.0)) if sol.beta1 > 0 else "ode"

    def integrate(self, sys, sol, u0, ts):
        kw, kind, par = self._pb(sys, sol)
        out, status, nacc, nrej = O.integrate(kind, par, u0, ts, controller=self._ctrl(sol), **kw)
        return out, status, nacc + nrej

    def integrate_fm(self, sys, sol, u0, ts, fm0=None):
        kw, kind, par = self._pb(sys, sol)
        return O.integrate_fm(kind, par, u0, ts, fm0, **kw)

    def integrate_events(self, sys, sol, u0, t_end, coef, offset=0.0, direction=0, max_events=1024):
        kw, kind, par = self._pb(sys, sol)
        r = O.integrate_events(kind, par, u0, t_end, coef, offset, direction, max_events, **kw)
        return (*r, 0.0)
