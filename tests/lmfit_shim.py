"""TEST INFRASTRUCTURE ONLY -- the slice of lmfit that ``PyRayHF.library.minimize_parameters`` uses
(library.py:781-808), for a container without lmfit: ``Parameters.add`` and ``minimize(..., method='brute')``.

lmfit's brute method hands ``scipy.optimize.brute`` one ``slice(min, max, brute_step)`` per varied parameter, the
objective ``sum(residual ** 2)`` and ``finish=None``; the result carries the grid node with the smallest objective.
scipy is installed here, so the search itself is the real ``scipy.optimize.brute``.  Parity status of the inversion
rows: pinned against the reference's own ``minimize_parameters`` / ``residual_VH`` driving THIS shim -- lmfit itself
is not available offline.
"""
import copy
import types

import numpy as np
import scipy.optimize


class Parameter:
    def __init__(self, name, value=None, vary=True, min=-np.inf, max=np.inf, brute_step=None):
        self.name, self.value, self.vary, self.min, self.max, self.brute_step = name, value, vary, min, max, brute_step


class Parameters(dict):
    def add(self, name, value=None, vary=True, min=-np.inf, max=np.inf, brute_step=None):
        self[name] = Parameter(name, value, vary, min, max, brute_step)


def minimize(fcn, params, args=(), method='brute'):
    if method != 'brute':
        raise ValueError("the shim implements method='brute' only")
    varied = [p for p in params.values() if p.vary]
    ranges = tuple(slice(float(p.min), float(p.max), float(p.brute_step)) for p in varied)
    work = copy.deepcopy(params)

    def objective(x):
        for p, v in zip(varied, np.atleast_1d(x)):
            work[p.name].value = float(v)
        r = np.asarray(fcn(work, *args), dtype=float)
        return float((r * r).sum())

    x0, fval, grid, jout = scipy.optimize.brute(objective, ranges, finish=None, full_output=True)
    best = copy.deepcopy(params)
    for p, v in zip(varied, np.atleast_1d(x0)):
        best[p.name].value = float(v)
    return types.SimpleNamespace(params=best, brute_x0=np.atleast_1d(x0), brute_fval=fval, brute_grid=grid,
                                 brute_Jout=jout)


def install(module):
    """Give the (stub) ``lmfit`` module object the two names the reference uses."""
    module.Parameters = Parameters
    module.minimize = minimize
    return module
