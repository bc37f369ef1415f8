"""ctypes wrapper of oracle/libjb_oracle.so (the plain-C restatement, jb_oracle.c).

TEST INFRASTRUCTURE ONLY: the checker at sizes the torch restatement is too slow for, and the CPU
baseline of bench.py.  Takes the same duck-typed model tree as jabble_oracle.tree_call.
"""

import ctypes as C
import os

import numpy as np

from . import jabble_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "libjb_oracle.so")
_i32p = C.POINTER(C.c_int32)


class _Comp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p_val", C.c_int32), ("xp0", C.c_double), ("dx", C.c_double),
                ("n_knots", C.c_int64), ("ctrl_offset", C.c_int64), ("n_ctrl_keys", C.c_int64),
                ("ctrl_key", _i32p), ("shift_offset", C.c_int64), ("n_shift", C.c_int64),
                ("shift_key", _i32p), ("stretch_offset", C.c_int64), ("n_stretch", C.c_int64),
                ("stretch_key", _i32p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} missing: run `make -C oracle`")
        _lib = C.CDLL(_PATH)
        _lib.jbo_eval.restype = C.c_int
        _lib.jbo_max_threads.restype = C.c_int
    return _lib


def _leaves(node, out):
    if O._kind(node) in O._CONTAINERS:
        for m in node.models:
            _leaves(m, out)
    else:
        out.append(node)
    return out


def flatten(model, metablock, n_rows):
    """Tree -> (component structs, params_all, fit index, keep-alive list).  Same grammar as the
    product's plan compiler, written independently of it."""
    leaves = _leaves(model, [])
    offs, o = {}, 0
    for lf in leaves:
        offs[id(lf)] = o
        o += int(np.asarray(lf.p).size)
    n_params = o
    params = np.concatenate([np.asarray(lf.p, dtype=np.float64).ravel() for lf in leaves])
    fit = np.concatenate([np.arange(offs[id(lf)], offs[id(lf)] + np.asarray(lf.p).size)
                          for lf in leaves if lf._fit] or [np.zeros(0, int)]).astype(np.int64)
    tops = model.models if O._kind(model) == "AdditiveModel" else [model]
    comps = (_Comp * len(tops))()
    keep = []

    def key(name):
        k = np.ascontiguousarray(np.asarray(metablock[name]).reshape(-1), dtype=np.int32)
        assert k.size == n_rows
        keep.append(k)
        return k.ctypes.data_as(_i32p)

    for cd, top in zip(comps, tops):
        chain = _leaves(top, [])
        cd.shift_offset = cd.stretch_offset = -1
        for n in chain:
            k = O._kind(n)
            if k == "ShiftingModel":
                cd.shift_offset, cd.n_shift, cd.shift_key = offs[id(n)], np.asarray(n.p).size, key(n.which_key)
            elif k == "StretchingModel":
                cd.stretch_offset, cd.n_stretch, cd.stretch_key = offs[id(n)], np.asarray(n.p).size, key(n.which_key)
            else:
                inner = n.model if k == "NormalizationModel" else n
                grid = np.asarray(inner.xs, dtype=np.float64).ravel()
                cd.kind = 1 if k == "NormalizationModel" else 0
                cd.p_val = int(inner.p_val)
                cd.xp0, cd.dx, cd.n_knots = float(grid[0]), float(grid[1] - grid[0]), grid.size
                cd.ctrl_offset = offs[id(n)]
                if cd.kind == 1:
                    cd.n_ctrl_keys, cd.ctrl_key = int(n.size), key(n.which_key)
    return comps, params, fit, keep, n_params


def evaluate(p_fit, datablock, metablock, model, coefficient=1.0, nthreads=0):
    """(loss, grad w.r.t. the fit vector, fisher_all, params_all layout offsets)."""
    xs = np.ascontiguousarray(datablock["xs"], dtype=np.float64)
    ys = np.ascontiguousarray(datablock["ys"], dtype=np.float64)
    iv = np.ascontiguousarray(datablock["yivar"], dtype=np.float64)
    mk = np.ascontiguousarray(np.asarray(datablock["mask"]).astype(np.uint8))
    E, P = xs.shape
    comps, params, fit, keep, n_params = flatten(model, metablock, E)
    p_fit = np.asarray(p_fit, dtype=np.float64).ravel()
    if p_fit.size:
        params[fit] = p_fit
    grad = np.zeros(n_params)
    fisher = np.zeros(n_params)
    loss = C.c_double()
    f64 = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = lib().jbo_eval(C.c_int64(E), C.c_int64(P), f64(xs), f64(ys), f64(iv),
                        mk.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int(len(comps)), comps, f64(params),
                        C.c_int64(n_params), C.c_double(coefficient), C.c_int(nthreads), C.byref(loss),
                        f64(grad), f64(fisher))
    if rc == -1:
        raise ValueError("C oracle: an unmasked pixel gathers outside the knot grid")
    if rc:
        raise MemoryError("C oracle")
    return loss.value, grad[fit], fisher, grad
