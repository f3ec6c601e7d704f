"""Rectangular linear sum assignment, restated from scipy's ``_lsap`` extension
(oracle; test infrastructure).

The reference calls ``scipy.optimize.linear_sum_assignment`` at
``/root/reference/src/tracker/core/linear_assignment.py:62``.  scipy is a
third-party dependency that is not vendored under ``/root/reference``
(requirements.txt pins 1.15.3; the build container and the GPU box hold 1.18.1).
Its published algorithm is the shortest-augmenting-path method of D. F. Crouse,
"On implementing 2D rectangular assignment algorithms", IEEE T-AES 52(4), 2016.
This file restates it with the exact scan order and tie rule the C++ uses, so
that a device implementation can be checked for *identical* assignments, ties
included.  ``solve`` (pure Python) documents the algorithm and serves small
cases; ``solve_c`` calls the plain-C restatement in ``oracle/csrc/lsap_ref.c``.
Both are pinned against scipy itself in ``tests/test_oracle_lsap.py``.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def solve(cost):
    """Return (row_ind, col_ind) exactly as scipy.optimize.linear_sum_assignment does."""
    cost = np.asarray(cost, dtype=np.float64)
    if cost.ndim != 2:
        raise ValueError("expected a matrix")
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.empty(0, np.int64), np.empty(0, np.int64)
    transpose = nc < nr
    if transpose:
        cost = cost.T.copy()
        nr, nc = nc, nr
    inf = float("inf")
    u = [0.0] * nr
    v = [0.0] * nc
    path = [-1] * nc
    col4row = [-1] * nr
    row4col = [-1] * nc
    c = cost.tolist()
    for cur in range(nr):
        min_val = 0.0
        i = cur
        SR = [False] * nr
        SC = [False] * nc
        spc = [inf] * nc
        remaining = [nc - 1 - it for it in range(nc)]  # reversed initial order
        num_remaining = nc
        sink = -1
        while sink == -1:
            index = -1
            lowest = inf
            SR[i] = True
            ci = c[i]
            ui = u[i]
            for it in range(num_remaining):
                j = remaining[it]
                r = min_val + ci[j] - ui - v[j]
                if r < spc[j]:
                    path[j] = i
                    spc[j] = r
                # first strict minimum in scan order, replaced by any later equal
                # value whose column is still unassigned
                if spc[j] < lowest or (spc[j] == lowest and row4col[j] == -1):
                    lowest = spc[j]
                    index = it
            min_val = lowest
            if min_val == inf:
                raise ValueError("cost matrix is infeasible")
            j = remaining[index]
            if row4col[j] == -1:
                sink = j
            else:
                i = row4col[j]
            SC[j] = True
            num_remaining -= 1
            remaining[index] = remaining[num_remaining]
        u[cur] += min_val
        for i2 in range(nr):
            if SR[i2] and i2 != cur:
                u[i2] += min_val - spc[col4row[i2]]
        for j2 in range(nc):
            if SC[j2]:
                v[j2] -= min_val - spc[j2]
        j = sink
        while True:
            i = path[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    if transpose:
        order = np.argsort(np.asarray(col4row), kind="stable")
        return np.asarray(col4row, np.int64)[order], order.astype(np.int64)
    return np.arange(nr, dtype=np.int64), np.asarray(col4row, np.int64)


def _load():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liblsap_ref.so")
        if not os.path.exists(path):
            raise RuntimeError(
                "oracle C restatement not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.lsap_ref_solve.restype = ctypes.c_int
        lib.lsap_ref_solve.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_void_p]
        _LIB = lib
    return _LIB


def solve_c(cost):
    """Same contract as ``solve`` through the plain-C restatement."""
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    nr, nc = cost.shape
    if nr == 0 or nc == 0:
        return np.empty(0, np.int64), np.empty(0, np.int64)
    k = min(nr, nc)
    rows = np.empty(k, np.int64)
    cols = np.empty(k, np.int64)
    rc = _load().lsap_ref_solve(nr, nc, cost.ctypes.data, rows.ctypes.data, cols.ctypes.data)
    if rc != 0:
        raise ValueError("lsap_ref_solve failed with %d" % rc)
    return rows, cols
